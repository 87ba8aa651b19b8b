"""Drop-in for the reference's ``infer.py`` entry point (REF/infer.py): same functions, arguments, CLI flags
and ``.lab`` output, but the model is built once, clips are batched, and everything between the waveform
and the segment list runs on the GPU (csrc/*.cu via pipeline.Labeler).

Differences that are deliberate and documented in DESIGN.md:
  * the model is constructed/loaded once per process, not once per file (REF/infer.py:204-208 per file);
  * no autograd graph is built (the reference never disables grad);
  * ``--sample/--top-k/--top-p/--temperature`` are accepted and validated like the reference, and like the
    reference they do not affect the output (REF/infer.py:283-297 overwrites the sampled ids).
"""
import os
import struct
import sys

import numpy as np
import torch
import yaml

from . import ingest, ops
from .model import BIOPhonemeTagger
from .pipeline import FRAME_DURATION, Labeler
from .utils import (canonical_to_lang, load_langs, load_phoneme_list, load_phoneme_merge_map, save_lab)

frame_duration = FRAME_DURATION
MAX_SEGMENT_DURATION = 30.0


def load_config(config_path="config.yaml"):
    with open(config_path, "r") as f:
        return yaml.safe_load(f)


def split_audio(audio, sr, max_duration=MAX_SEGMENT_DURATION):
    """REF/infer.py:19-28: consecutive chunks of at most ``max_duration`` seconds, no overlap."""
    step = int(max_duration * sr)
    return [audio[s:min(s + step, len(audio))] for s in range(0, len(audio), step)]


def align_phoneme_list(segments_pred, forced_list):
    """REF/infer.py:30-60: map a forced phoneme sequence onto predicted segments (greedy in-order label match,
    then fill the unmatched forced entries with the unused predictions in order).  Host-side list logic."""
    taken = set()
    assigned = [None] * len(forced_list)
    cursor = 0
    for fi, want in enumerate(forced_list):
        for pi in range(cursor, len(segments_pred)):
            if pi not in taken and segments_pred[pi][2] == want:
                assigned[fi] = pi
                taken.add(pi)
                cursor = pi + 1
                break
    free = 0
    for fi in range(len(forced_list)):
        if assigned[fi] is None:
            while free < len(segments_pred) and free in taken:
                free += 1
            if free < len(segments_pred):
                assigned[fi] = free
                taken.add(free)
                free += 1
    out = []
    for fi, want in enumerate(forced_list):
        pi = assigned[fi]
        if pi is not None and pi < len(segments_pred):
            out.append((segments_pred[pi][0], segments_pred[pi][1], want))
    return out


def suppress_low_confidence(logits, id2label, threshold=0.5):
    """REF/infer.py:86-96 on a [T, L] logits tensor -> list of tag strings (runs wfl_decode_frames)."""
    lg = logits.float().cuda().contiguous()
    o_id = [k for k, v in id2label.items() if v == "O"]
    if not o_id:
        raise KeyError("O")
    ids = torch.empty(lg.shape[0], dtype=torch.int32, device=lg.device)
    ops.decode_frames(lg, lg.shape[1], o_id[0], float(threshold), ids)
    return [id2label[i] for i in ids.cpu().tolist()]


def read_audio(path):
    """Returns (float64 samples [N] mono, sample_rate).  Uses soundfile when installed (as the reference does,
    REF/infer.py:217); otherwise a built-in RIFF/WAVE reader for PCM 8/16/24/32 and IEEE float 32/64."""
    try:
        import soundfile as sf
        audio, sr = sf.read(path)
    except ImportError:
        audio, sr = _read_wav(path)
    audio = np.asarray(audio, dtype=np.float64)
    if audio.ndim == 2:
        audio = audio.mean(axis=1)
    return audio, int(sr)


def _read_wav(path):
    with open(path, "rb") as f:
        data = f.read()
    if data[:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise ValueError(f"{path}: not a RIFF/WAVE file")
    pos, fmt, pcm = 12, None, None
    while pos + 8 <= len(data):
        cid, size = data[pos:pos + 4], struct.unpack("<I", data[pos + 4:pos + 8])[0]
        body = data[pos + 8:pos + 8 + size]
        if cid == b"fmt ":
            fmt = struct.unpack("<HHIIHH", body[:16])
            if fmt[0] == 0xFFFE and len(body) >= 26:  # WAVE_FORMAT_EXTENSIBLE: real tag is in the sub-format GUID
                fmt = (struct.unpack("<H", body[24:26])[0],) + fmt[1:]
        elif cid == b"data":
            pcm = body
        pos += 8 + size + (size & 1)
    if fmt is None or pcm is None:
        raise ValueError(f"{path}: missing fmt/data chunk")
    tag, ch, sr, _, _, bits = fmt
    if tag == 1:
        if bits == 8:
            x = (np.frombuffer(pcm, dtype=np.uint8).astype(np.float64) - 128.0) / 128.0
        elif bits == 16:
            x = np.frombuffer(pcm, dtype="<i2").astype(np.float64) / 32768.0
        elif bits == 24:
            b = np.frombuffer(pcm[:len(pcm) // 3 * 3], dtype=np.uint8).reshape(-1, 3).astype(np.int32)
            v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
            x = (v - ((v & 0x800000) << 1)).astype(np.float64) / 8388608.0
        elif bits == 32:
            x = np.frombuffer(pcm, dtype="<i4").astype(np.float64) / 2147483648.0
        else:
            raise ValueError(f"{path}: unsupported PCM width {bits}")
    elif tag == 3:
        x = np.frombuffer(pcm, dtype="<f4" if bits == 32 else "<f8").astype(np.float64)
    else:
        raise ValueError(f"{path}: unsupported WAVE format tag {tag}")
    if ch > 1:
        x = x[:len(x) // ch * ch].reshape(-1, ch)
    return x, sr


class _Session:
    """Model + labeler built once per (config, checkpoint, device)."""
    _cache = {}

    def __init__(self, config_path, checkpoint_path, device):
        self.config = load_config(config_path)
        save_dir = self.config["output"]["save_dir"]
        self.labels = load_phoneme_list(os.path.join(save_dir, "phonemes.txt"))
        self.lang2id = load_langs(os.path.join(save_dir, "langs.txt"))
        mm_path = os.path.join(save_dir, "phoneme_merge_map.json")
        self.merge_map = load_phoneme_merge_map(mm_path) if os.path.exists(mm_path) else None
        if not str(device).startswith("cuda") or not torch.cuda.is_available():
            raise RuntimeError("wfl_asr_b200 runs on a CUDA device only (no CPU path); got device=%r" % (device,))
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        model = BIOPhonemeTagger(self.config, self.labels)
        state_dict = torch.load(checkpoint_path, map_location="cpu", weights_only=True)
        model.load_state_dict(state_dict)
        self.model = model.to(self.device).eval()
        pp = self.config["postprocess"]
        self.labeler = Labeler(self.model, median_filter=pp["median_filter"], merge_mode=pp["merge_segments"])

    @classmethod
    def get(cls, config_path, checkpoint_path, device):
        key = (os.path.abspath(config_path), os.path.abspath(checkpoint_path), str(device),
               os.path.getmtime(checkpoint_path), os.path.getmtime(config_path))
        if key not in cls._cache:
            cls._cache = {key: cls(config_path, checkpoint_path, device)}
        return cls._cache[key]


def _normalised_clips(audio, sr, dev):
    """REF/infer.py:234-244 + :113-115: whole-file peak normalisation, split into <= 30 s chunks when longer, each
    chunk normalised again by its own peak.  Returns (fp32 [n_clips, width] device tensor, chunk lengths)."""
    n = len(audio)
    flat = audio if torch.is_tensor(audio) else torch.from_numpy(np.ascontiguousarray(audio, dtype=np.float64)).to(dev)
    scratch = torch.empty(64, dtype=torch.float64, device=dev)
    # clip boundaries are built ON the device (arange): torch.tensor(host list, device=...) is a synchronous pageable
    # copy, 0.2 ms per file in a folder run
    whole_file = torch.arange(0, n + 1, n, dtype=torch.int64, device=dev)  # [0, n]
    if n / sr > MAX_SEGMENT_DURATION:
        step = int(MAX_SEGMENT_DURATION * sr)
        lens = [min(step, n - s) for s in range(0, n, step)]
        whole = torch.empty_like(flat)
        ops.peak_normalize(flat, whole_file, 1, None, scratch, out_f64=whole)
        begins = torch.arange(0, (len(lens) + 1) * step, step, dtype=torch.int64, device=dev).clamp_(max=n)  # [0, step, .., n]
        if len(lens) > scratch.numel():
            scratch = torch.empty(len(lens), dtype=torch.float64, device=dev)
        out = torch.empty(len(lens), step, device=dev)
        ops.peak_normalize(whole, begins, len(lens), out, scratch)
        return out, lens, True
    out = torch.empty(1, n, device=dev)
    ops.peak_normalize(flat, whole_file, 1, out, scratch)
    return out, [n], False


def _forward_clips(sess, clip_rows, lens, lang_id, batch_clips=32):
    """Runs the model over clips (``clip_rows[i]``: fp32 device row holding ``lens[i]`` valid samples).  Whisper
    pads/truncates every clip to 30 s itself (zero padding, as its feature extractor does), so clips of any files
    batch together; WavLM is length-sensitive (no attention mask in the reference), so its
    clips run at their exact lengths, equal lengths sharing a batch.  Returns per-clip (logits [T, L], offsets [T, 2])."""
    model, dev = sess.model, sess.device
    lang_ids = [lang_id] if lang_id is not None else list(sess.lang2id.values())
    if lang_id is not None and lang_id > max(sess.lang2id.values()):
        raise ValueError(f"Error: Language ID ({lang_id}) is higher than the latest ID ({max(sess.lang2id.values())}) "
                         f"of this model.\n Languages and Codes available: {sess.lang2id}")
    groups = {}
    if model.encoder_type == "whisper":
        groups[None] = list(range(len(lens)))
    else:
        for i, ln in enumerate(lens):
            groups.setdefault(ln, []).append(i)
    logits_out, offsets_out = [None] * len(lens), [None] * len(lens)
    batched = []
    for ln, members in groups.items():
        for s0 in range(0, len(members), batch_clips):
            idx = members[s0:s0 + batch_clips]
            if ln is None:
                width = max(lens[i] for i in idx)
                wave = torch.zeros(len(idx), width, device=dev)
                for j, i in enumerate(idx):
                    wave[j, :lens[i]] = clip_rows[i][:lens[i]]
            else:
                wave = torch.stack([clip_rows[i][:ln] for i in idx])
            if len(lang_ids) == 1:
                lt = torch.full((len(idx),), lang_ids[0], dtype=torch.long, device=dev)
                acc_l, acc_o = model(wave, lt)  # fresh tensors (model.forward clones the engine's views)
            else:  # REF/infer.py:265-276: mean over languages when --lang-id is unset (encoder runs once here)
                acc_l, acc_o = model.forward_language_mean(wave, lang_ids)
            batched.append((acc_l, acc_o))
            for j, i in enumerate(idx):
                logits_out[i], offsets_out[i] = acc_l[j], acc_o[j]
    if model.encoder_type == "whisper":  # one group in clip order, one frame count: hand the batches back whole
        return (torch.cat([b[0] for b in batched]) if len(batched) > 1 else batched[0][0],
                torch.cat([b[1] for b in batched]) if len(batched) > 1 else batched[0][1])
    return logits_out, offsets_out


def _prepare_file(sess, audio_path):
    """REF/infer.py:191-244 for one file: forced phoneme list, decode, resample, normalise, split into <= 30 s chunks."""
    dev = sess.device
    forced = None
    phoneme_txt = audio_path.replace(".wav", ".txt")
    if os.path.exists(phoneme_txt):
        forced = []
        with open(phoneme_txt, "r", encoding="utf-8") as f:
            for line in f:
                forced.extend(line.strip().split())
        print(f"Loaded forced phoneme list with {len(forced)} phonemes.")
    audio, sr = read_audio(audio_path)
    return _prepare_audio(sess, audio_path, audio, sr, forced)


def _prepare_audio(sess, audio_path, audio, sr, forced):
    """``audio``: decoded samples (numpy, as soundfile returns them) or an fp64 mono device tensor (ingest.FolderIngest)."""
    dev = sess.device
    if len(audio) == 0:
        raise ValueError(f"{audio_path}: empty audio")
    target_sr = sess.config["data"]["sample_rate"]
    if sr != target_sr:  # REF/infer.py:217-220 (torchaudio.functional.resample on the host) -> csrc/resample.cu
        audio = ingest.resample(audio if torch.is_tensor(audio) else ingest.to_device_mono(audio, dev), sr, target_sr)
        sr = target_sr
    clips, lens, chunked = _normalised_clips(audio, sr, dev)
    if chunked:
        print(f"Audio is too long ({len(audio)/sr:.1f}s), splitting...")
    return dict(path=audio_path, clips=clips, lens=lens, chunked=chunked, forced=forced, sr=sr)


def _cache_paths(f, lang_id):
    """REF/infer.py:222-229 (single clip) and :120-126 (30 s chunks): the reference's ``.wfl_cache`` file names, per clip."""
    base_name = os.path.splitext(os.path.basename(f["path"]))[0]
    cache_dir = os.path.join(os.path.dirname(f["path"]), ".wfl_cache")
    lang_suffix = f"_lang{lang_id}" if lang_id is not None else "_avg"
    if f["chunked"]:
        return [(os.path.join(cache_dir, f"{base_name}_seg{idx}{lang_suffix}_logits.pt"),
                 os.path.join(cache_dir, f"{base_name}_seg{idx}{lang_suffix}_offsets.pt")) for idx in range(len(f["lens"]))]
    return [(os.path.join(cache_dir, f"{base_name}{lang_suffix}_logits.pt"),
             os.path.join(cache_dir, f"{base_name}{lang_suffix}_offsets.pt"))]


def _forward_clips_cached(sess, files, rows, lens, lang_id, quiet):
    """The reference's logits cache (REF/infer.py:222-232,246-249,278-280 and :120-131,158-161): logits / offsets of a
    clip are read from ``<audio dir>/.wfl_cache/`` when present, else computed and written there.  Like the reference's,
    the cache is keyed by file name and language only -- not by checkpoint or config -- so it is the caller's job to
    clear it when either changes (``use_cache=False`` / WFL_NO_CACHE=1 bypasses it)."""
    paths = [p for f in files for p in _cache_paths(f, lang_id)]
    hit = [os.path.exists(p[0]) for p in paths]
    if not any(hit):
        logits, offsets = _forward_clips(sess, rows, lens, lang_id)
    else:
        miss = [i for i, h in enumerate(hit) if not h]
        lg_m, of_m = _forward_clips(sess, [rows[i] for i in miss], [lens[i] for i in miss], lang_id) if miss else ([], [])
        logits, offsets = [None] * len(rows), [None] * len(rows)
        for k, i in enumerate(miss):
            logits[i], offsets[i] = lg_m[k], of_m[k]
        for i, h in enumerate(hit):
            if h:
                if not quiet:
                    print(f"Loaded cached logits for {os.path.basename(paths[i][0])}")
                lg = torch.load(paths[i][0], map_location=sess.device, weights_only=False)
                logits[i] = lg.squeeze(0).float()
                offsets[i] = (torch.load(paths[i][1], map_location=sess.device, weights_only=False).float()
                              if os.path.exists(paths[i][1]) else torch.zeros(logits[i].shape[0], 2, device=sess.device))
    for i, h in enumerate(hit):
        if not h:
            os.makedirs(os.path.dirname(paths[i][0]), exist_ok=True)
            torch.save(logits[i].unsqueeze(0).clone(), paths[i][0])  # [1, T, L] like the reference's avg_logits
            torch.save(offsets[i].clone(), paths[i][1])              # [T, 2]
    return logits, offsets


def _label_files(sess, files, lang_id, confidence_threshold, with_text=False, use_cache=False, quiet=False):
    """Forward + post-processing for a list of prepared files in one go (REF/infer.py:246-319 per file): all clips of
    all files share the model batches, one post-processing pass decodes every clip and merges the chunks of each
    file, one D2H copy brings all segment records back.  Returns one segment list per file."""
    dev = sess.device
    lang_name = None
    if lang_id is not None:
        for n, i in sess.lang2id.items():
            if i == lang_id:
                lang_name = n
                break
    rows, lens, shifts, begins = [], [], [], [0]
    for f in files:
        t = 0.0
        for j, ln in enumerate(f["lens"]):  # REF/infer.py:180-182: chunk j starts where the previous ones ended
            rows.append(f["clips"][j])
            lens.append(ln)
            shifts.append(t)
            t += ln / f["sr"]
        begins.append(len(rows))
    if use_cache and not os.environ.get("WFL_NO_CACHE"):
        logits, offsets = _forward_clips_cached(sess, files, rows, lens, lang_id, quiet)
    else:
        logits, offsets = _forward_clips(sess, rows, lens, lang_id)

    labeler = sess.labeler
    labeler.threshold = float(confidence_threshold)
    names = None
    if sess.merge_map and lang_name:  # REF/infer.py:303-307 remap before merging
        names = [canonical_to_lang(p, lang_name, sess.merge_map) for p in labeler.phon]
    labeler.set_output_names(names)
    # every clip has the same T for Whisper; WavLM clips differ -> pad to the longest, decode each on its own length
    n = len(lens)
    if torch.is_tensor(logits):  # Whisper: [n, 1500, L] as the model produced it
        lg, of = logits, offsets
        T = lg.shape[1]
        tl = [T] * n
    else:
        T = max(l.shape[0] for l in logits)
        lg = torch.zeros(n, T, logits[0].shape[-1], device=dev)
        of = torch.zeros(n, T, 2, device=dev)
        tl = []
        for i in range(n):
            lg[i, :logits[i].shape[0]] = logits[i]
            of[i, :offsets[i].shape[0]] = offsets[i]
            tl.append(logits[i].shape[0])
    _, merged, nout, fcb, n_files = labeler.postprocess(
        lg, of, torch.tensor(tl, dtype=torch.int32, device=dev),
        torch.tensor(begins, dtype=torch.int32, device=dev),
        torch.tensor(shifts, dtype=torch.float64, device=dev))  # + 0.0 for single-chunk files: exact in fp64
    if with_text:
        return labeler.fetch_with_htk(merged, nout, fcb, n_files, T)
    return labeler.fetch(merged, nout, fcb, n_files, T)


def _finish_file(f, segments_pred, output_lab_path, quiet=False, lab_text=None):
    """REF/infer.py:312-328: optional forced-phoneme alignment, .lab output.  ``lab_text``: the .lab text of
    ``segments_pred`` when the caller already has it (computed for the whole pass on the device)."""
    forced = f["forced"]
    if forced is None and lab_text is not None and output_lab_path:
        dir_path = os.path.dirname(output_lab_path)
        if dir_path:
            os.makedirs(dir_path, exist_ok=True)
        with open(output_lab_path, "w", encoding="utf-8") as fh:
            fh.write(lab_text)
        if not quiet:
            print(f"Predictions saved to: {output_lab_path}")
        return segments_pred
    if forced is not None:
        aligned = align_phoneme_list(segments_pred, forced)
        if "SP" not in forced and "AP" not in forced:
            before = [s for s in segments_pred if s[2] in ("SP", "AP") and s[1] <= aligned[0][0]]
            after = [s for s in segments_pred if s[2] in ("SP", "AP") and s[0] >= aligned[-1][1]]
            segments_pred = before + aligned + after
        else:
            segments_pred = aligned
    if output_lab_path:
        dir_path = os.path.dirname(output_lab_path)
        if dir_path:
            os.makedirs(dir_path, exist_ok=True)
        save_lab(output_lab_path, segments_pred)
        if not quiet:
            print(f"Predictions saved to: {output_lab_path}")
    return segments_pred


def infer_audio(audio_path, config_path="config.yaml", checkpoint_path="best_model.pt",
                output_lab_path=None, device="cuda", lang_id=None,
                sample=False, top_k=0, top_p=0.0, temperature=1.0,
                confidence_threshold=0.0, use_cache=True):
    """REF/infer.py:186-328.  ``use_cache``: the reference's ``.wfl_cache`` logits cache beside the audio file (on, as in
    the reference; see _forward_clips_cached for its staleness caveat)."""
    sess = _Session.get(config_path, checkpoint_path, device)
    with torch.cuda.device(sess.device):  # every launch below goes to the streams of the session's device (-d cuda:1)
        f = _prepare_file(sess, audio_path)
        segments_pred = _label_files(sess, [f], lang_id, confidence_threshold, use_cache=use_cache)[0]
        return _finish_file(f, segments_pred, output_lab_path)


def infer_folder(folder_path: str, config_path: str = "config.yaml", checkpoint_path: str = "best_model.pt",
                 output_dir: str = "outputs", device: str = "cuda", lang_id: int = None,
                 sample=False, top_k=0, top_p=0.0, temperature=1.0, confidence_threshold=0.0,
                 files_per_pass: int = 128, decode_workers: int = 8, quiet: bool = False, use_cache: bool = True):
    """REF/infer.py:330-357, same files, same .lab outputs and the same per-file printout, but the folder is labeled
    ``files_per_pass`` files at a time: their audio is decoded on a thread pool, and all their clips share the model
    batches and one post-processing pass (``_label_files``) instead of one batch-1 pass per file.  ``quiet`` drops the
    reference's per-file / per-segment printout (bulk use)."""
    import builtins
    print = (lambda *a, **k: None) if quiet else builtins.print  # noqa: A001
    wav_files = [f for f in os.listdir(folder_path) if f.lower().endswith(".wav")]
    os.makedirs(output_dir, exist_ok=True)
    sess = _Session.get(config_path, checkpoint_path, device)
    results = {}
    with torch.cuda.device(sess.device):
        # decode_workers threads read the files into pinned memory; PCM -> float64, resampling and normalisation run on
        # the device (ingest.FolderIngest).  ONE ingest spans the folder: while a pass is being labeled its workers
        # already read the next pass's files (a window of 4 x decode_workers files is in flight)
        decoded = iter(ingest.FolderIngest([str(os.path.join(folder_path, w)) for w in wav_files], sess.device, read_audio,
                                           workers=decode_workers))
        for s0 in range(0, len(wav_files), max(1, files_per_pass)):
            names = wav_files[s0:s0 + max(1, files_per_pass)]
            files = []
            for w in names:
                path, audio, sr = next(decoded)
                print(f"\nInferencing: {w}")
                forced = None
                phoneme_txt = path.replace(".wav", ".txt")
                if os.path.exists(phoneme_txt):
                    forced = []
                    with open(phoneme_txt, "r", encoding="utf-8") as fh:
                        for line in fh:
                            forced.extend(line.strip().split())
                    print(f"Loaded forced phoneme list with {len(forced)} phonemes.")
                files.append(_prepare_audio(sess, path, audio, sr, forced))
            per_file, texts = _label_files(sess, files, lang_id, confidence_threshold, with_text=True, use_cache=use_cache,
                                           quiet=quiet)
            for w, f, segs, text in zip(names, files, per_file, texts):
                output_lab_path = os.path.join(output_dir, w.replace(".wav", ".lab"))
                segments = _finish_file(f, segs, str(output_lab_path), quiet=quiet, lab_text=text)
                results[w] = segments
                if not quiet:  # (the f-strings alone cost 0.3 ms per file)
                    print("Predicted segments:")
                    for start, end, ph in segments:
                        print(f"({round(start, 2)}, {round(end, 2)}, {ph})")
    return results


def cli_command():
    """The click command behind ``python -m wfl_asr_b200.infer`` (same flags as REF/infer.py:362-373)."""
    import click
    from pathlib import Path

    @click.command(help='Infer with WFL')
    @click.argument('path', metavar='PATH')
    @click.option('--checkpoint', '-ckpt', type=str, required=True, help='Path to WFL Checkpoint.')
    @click.option('--config', '-c', type=str, required=True, help='Path to Config file.')
    @click.option('--output', '-o', type=str, required=False, default=".", help='Path to output labels.')
    @click.option('--lang-id', '-l', type=int, required=False, default=None, help='Language ID.')
    @click.option('--sample', '-s', is_flag=True, help='Enable sampling instead of argmax')
    @click.option('--top-k', '-tk', type=int, default=0, help='Top-K sampling (range: 1-20)')
    @click.option('--top-p', '-tp', type=float, default=0.0, help='Top-P sampling (range: 0.1-1)')
    @click.option('--temperature', '-temp', type=float, default=1.0, help='Sampling temperature (range: 0.1-2)')
    @click.option('--device', '-d', type=str, default="auto", help='Device to use: "cuda" or "cuda:0".')
    @click.option('--confidence-threshold', '-ct', type=float, default=None,
                  help='Suppress predictions with low confidence. Set 0 to disable.')
    def main(path, checkpoint, config, output, lang_id, sample, top_k, top_p, temperature, device, confidence_threshold):
        if sample:  # same validation and messages as REF/infer.py:377-392
            if top_k <= 0 and top_p <= 0.0:
                print("Sampling is enabled but neither --top-k nor --top-p is set.")
                sys.exit(1)
            if top_k > 0 and top_p > 0.0:
                print("You can't use both --top-k and --top-p at the same time.")
                sys.exit(1)
            if top_k < 0:
                print("top-k must be ≥ 1.")
                sys.exit(1)
            if top_p < 0.0 or top_p > 1.0:
                print("top-p must be between 0.1 and 1.0.")
                sys.exit(1)
            if temperature <= 0.0:
                print("temperature must be greater than 0.")
                sys.exit(1)
        requested = device.lower()
        if requested == "auto":
            device = "cuda"
        if not torch.cuda.is_available() or not device.startswith("cuda"):
            print("wfl_asr_b200 needs a CUDA device (B200, sm_100a); there is no CPU path.", file=sys.stderr)
            sys.exit(1)
        inf_path = Path(path)
        cfg = load_config(Path(config))
        if confidence_threshold is None:
            confidence_threshold = cfg["postprocess"].get("confidence_threshold", 0.0)
        output_path = inf_path if output == "." else output
        if not inf_path.exists():
            print(f"Unable to locate folder {str(inf_path)}")
            sys.exit(1)
        if lang_id is not None and lang_id <= -1:
            lang_id = None
        kw = dict(config_path=str(config), checkpoint_path=str(checkpoint), device=device, lang_id=lang_id,
                  sample=sample, top_k=top_k, top_p=top_p, temperature=temperature,
                  confidence_threshold=confidence_threshold)
        if inf_path.is_dir():
            infer_folder(folder_path=str(inf_path), output_dir=str(output_path), **kw)
        else:
            segments = infer_audio(audio_path=str(inf_path), output_lab_path=str(output_path), **kw)
            print("Predicted segments:")
            for start, end, ph in segments:
                print(f"({round(start, 2)}, {round(end, 2)}, {ph})")

    return main


def _cli():
    cli_command()()


if __name__ == "__main__":
    _cli()
