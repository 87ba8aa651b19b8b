"""Measured ingest pipeline: WAV files -> pinned staging -> H2D -> device PCM decode ->
(resample) -> peak-normalised fp32 clips, alone and in front of the labeler (files -> .lab).

    python tools/ingest_bench.py [--files 256] [--seconds 30] [--workers 8] [--workload cfg2]

Prints one JSON object: ingest-only audio-s/s and GB/s of PCM, files -> .lab audio-s/s, and the model-only rate on
device-resident clips for comparison.  bench.py embeds the same record under "ingest"."""
import argparse
import json
import os
import shutil
import struct
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def write_wav16(path, x, sr=16000):
    pcm = (np.clip(x, -1, 1) * 32767.0).astype("<i2").tobytes()
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + len(pcm)) + b"WAVEfmt " + struct.pack("<IHHIIHH", 16, 1, 1, sr, sr * 2, 2, 16))
        f.write(b"data" + struct.pack("<I", len(pcm)) + pcm)


def measure(files=256, seconds=30.0, workers=8, workload="cfg2", dev=None):
    import yaml

    from wfl_asr_b200 import infer, ingest, synth
    from wfl_asr_b200.model import BIOPhonemeTagger
    dev = dev or torch.device("cuda", torch.cuda.current_device())
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    tmp = tempfile.mkdtemp(prefix="wfl_ingest_", dir=base)
    try:
        pool = [synth.synth_wave(9000 + i, seconds) * 0.8 for i in range(4)]
        wav_dir = os.path.join(tmp, "wavs")
        os.makedirs(wav_dir)
        for i in range(files):
            write_wav16(os.path.join(wav_dir, f"utt{i:05d}.wav"), pool[i % 4] * (0.6 + 0.4 * ((i * 7) % 11) / 11.0))
        paths = sorted(os.path.join(wav_dir, f) for f in os.listdir(wav_dir))
        audio_s = files * seconds
        # ---- ingest only: files -> normalised fp32 clips on the device
        sess = type("S", (), {})()
        sess.device, sess.config = dev, {"data": {"sample_rate": 16000}}

        def ingest_pass():
            n = 0
            ing = ingest.FolderIngest(paths, dev, infer.read_audio, workers=workers)
            for path, audio, sr in ing:
                f = infer._prepare_audio(sess, path, audio, sr, None)
                n += sum(f["lens"])
            torch.cuda.synchronize(dev)
            return n, ing.bytes_read

        ingest_pass()
        t = time.perf_counter()
        n_samples, n_bytes = ingest_pass()
        dt_ingest = time.perf_counter() - t
        assert n_samples == int(files * seconds * 16000)
        # ---- files -> .lab through the infer.py drop-in
        cfg = synth.workload_config(workload)
        cfg["output"] = {"save_dir": tmp}
        labels = synth.synth_labels(30)
        model = synth.bench_model(BIOPhonemeTagger, cfg, labels)
        torch.save(model.state_dict(), os.path.join(tmp, "best_model.pt"))
        with open(os.path.join(tmp, "phonemes.txt"), "w") as f:
            f.write("\n".join(labels) + "\n")
        with open(os.path.join(tmp, "langs.txt"), "w") as f:
            f.write("en,0\nja,1\n")
        with open(os.path.join(tmp, "config.yaml"), "w") as f:
            yaml.safe_dump(cfg, f)
        args = (os.path.join(tmp, "config.yaml"), os.path.join(tmp, "best_model.pt"))
        kw = dict(device=str(dev), lang_id=0, confidence_threshold=cfg["postprocess"]["confidence_threshold"],
                  files_per_pass=64, decode_workers=workers, quiet=True, use_cache=False)
        infer.infer_folder(wav_dir, *args, output_dir=os.path.join(tmp, "warm"), **kw)
        torch.cuda.synchronize(dev)
        t = time.perf_counter()
        if os.environ.get("WFL_INGEST_PROFILE"):
            import cProfile, pstats
            pr = cProfile.Profile()
            pr.enable()
        res = infer.infer_folder(wav_dir, *args, output_dir=os.path.join(tmp, "labs"), **kw)
        torch.cuda.synchronize(dev)
        dt_lab = time.perf_counter() - t
        if os.environ.get("WFL_INGEST_PROFILE"):
            pr.disable()
            pstats.Stats(pr, stream=sys.stderr).sort_stats("cumulative").print_stats(45)
        n_lab = len([f for f in os.listdir(os.path.join(tmp, "labs")) if f.endswith(".lab")])
        assert n_lab == files and len(res) == files
        return {"files": files, "clip_seconds": seconds, "audio_seconds": audio_s, "format": "RIFF/WAVE PCM16 mono 16 kHz",
                "storage": "tmpfs" if base else "tmp dir", "decode_workers": workers,
                "ingest_only": {"audio_s_per_s": round(audio_s / dt_ingest, 1), "pcm_GBps": round(n_bytes / dt_ingest / 1e9, 3),
                                "what": "file read into pinned memory -> H2D -> device PCM decode -> fp64 peak normalise -> fp32 clips"},
                "files_to_lab": {"audio_s_per_s": round(audio_s / dt_lab, 1), "workload": workload,
                                 "what": "infer.infer_folder(quiet): ingest + forward + post-processing + .lab files written"}}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--files", type=int, default=256)
    ap.add_argument("--seconds", type=float, default=30.0)
    ap.add_argument("--workers", type=int, default=8)
    ap.add_argument("--workload", default="cfg2")
    a = ap.parse_args()
    print(json.dumps(measure(a.files, a.seconds, a.workers, a.workload)))
