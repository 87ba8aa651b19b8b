"""CPU: the C-ABI library builds, loads and exports every symbol include/wfl_b200.h declares; host-side logic
(packing, label tables, wav reader, forced alignment, drop-in model state_dict) with no compute calls."""
import os
import re
import struct
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def test_library_exports_header_symbols():
    import __graft_entry__ as ge
    ge.build()
    from wfl_asr_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "wfl_b200.h")).read()
    declared = set(re.findall(r"^(?:int|void|const char\*)\s+(wfl_\w+)\s*\(", header, re.M))
    assert len(declared) >= 22
    for name in ("wfl_create", "wfl_set_weight", "wfl_finalize", "wfl_forward", "wfl_postprocess", "wfl_destroy"):
        assert name in declared  # the handle-level surface of SURVEY.md section 8b
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/wfl_b200.h but not exported"
    bound = set(_lib.SIGNATURES) | set(_lib.NOARG) | set(_lib.VOID)
    assert declared <= bound, f"header symbols without a ctypes signature: {declared - bound}"
    assert lib.wfl_abi_version() == 1


def test_ctypes_structs_match_the_c_header(tmp_path):
    """The ctypes mirrors in _lib.py must have the C compiler's layout of include/wfl_b200.h (size and the offsets of
    the fields a mismatch would silently corrupt), and the header's constants must equal the Python ones."""
    import ctypes
    import shutil
    import subprocess
    from wfl_asr_b200 import _lib, ops
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    src = tmp_path / "layout.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "wfl_b200.h"\n'
        'int main(void) { printf("%zu %zu %zu %zu %zu %zu %zu %d %d %zu %zu %zu %zu\\n", sizeof(wfl_gemm_desc), offsetof(wfl_gemm_desc, w), '
        'offsetof(wfl_gemm_desc, bias), offsetof(wfl_gemm_desc, out), offsetof(wfl_gemm_desc, tile_n), '
        'offsetof(wfl_gemm_desc, out_col_group_stride), sizeof(wfl_segment), (int)WFL_MAX_SLABS, (int)WFL_WAVLM_STATS_DOUBLES, '
        'sizeof(wfl_config), offsetof(wfl_config, n_labels), offsetof(wfl_config, max_batch), offsetof(wfl_config, wavlm_layer_norm)); return 0; }\n')
    exe = tmp_path / "layout"
    subprocess.check_call([cc, "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    G = _lib.GemmDesc
    want = [ctypes.sizeof(G), G.w.offset, G.bias.offset, G.out.offset, G.tile_n.offset, G.out_col_group_stride.offset,
            ctypes.sizeof(_lib.Segment), _lib.WFL_MAX_SLABS, ops.WAVLM_STATS_DOUBLES,
            ctypes.sizeof(_lib.Config), _lib.Config.n_labels.offset, _lib.Config.max_batch.offset,
            _lib.Config.wavlm_layer_norm.offset]
    assert got == want, (got, want)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_compute_fails_loudly_without_gpu():
    from wfl_asr_b200 import ops
    a = torch.zeros(128, 64, dtype=torch.float16)
    with pytest.raises(ops.WflError):
        ops.linear(a, a, torch.zeros(128, 128, dtype=torch.float16))
    from wfl_asr_b200 import utils
    with pytest.raises(ops.WflError):
        utils.decode_bio_tags(["O", "B-a"])


def test_state_dict_matches_reference_layout():
    """Strict load of weights laid out by the oracle's key/shape table, which tests/test_oracle_forward.py pins to the
    reference module's own state_dict."""
    import make_forward_golden as mfg
    from wfl_asr_b200.model import BIOPhonemeTagger
    for name in mfg.CASES:
        cfg, labels, sd, _, _ = mfg.case_inputs(name)
        model = BIOPhonemeTagger(cfg, labels)
        own = model.state_dict()
        assert set(own) == set(sd), (name, sorted(set(own) ^ set(sd))[:6])
        for k in sd:
            assert own[k].shape == sd[k].shape and own[k].dtype == sd[k].dtype, (name, k)
            if k.startswith("mel_extractor."):  # buffers, not weights: the holder's defaults are torchaudio's values
                assert torch.equal(own[k], sd[k]), (name, k)
        model.load_state_dict(sd, strict=True)
        assert model.label2id["O"] == labels.index("O") and model.encoder_type in ("whisper", "wavlm", "none")


def test_model_rejects_bad_encoder_type():
    from wfl_asr_b200.model import BIOPhonemeTagger
    from wfl_asr_b200 import synth
    cfg = synth.workload_config("cfg2")
    cfg["model"]["encoder_type"] = "hubert"
    with pytest.raises(ValueError):
        BIOPhonemeTagger(cfg, ["O"])


def test_packing_folds():
    from wfl_asr_b200 import packing
    g = torch.Generator().manual_seed(0)
    w, b = torch.randn(8, 6, 5, generator=g), torch.randn(8, generator=g)
    gamma, beta = torch.randn(8, generator=g), torch.randn(8, generator=g)
    mean, var = torch.randn(8, generator=g), torch.rand(8, generator=g) + 0.5
    x = torch.randn(2, 6, 20, generator=g)
    ref = torch.nn.functional.batch_norm(torch.nn.functional.conv1d(x, w, b, padding=2), mean, var, gamma, beta, False, 0.0, 1e-5)
    wf, bf = packing.fold_batchnorm(w, b, gamma, beta, mean, var)
    assert (torch.nn.functional.conv1d(x, wf, bf, padding=2) - ref).abs().max() < 1e-5
    taps = packing.conv_taps(w)
    assert torch.equal(taps.view(8, 5, 6)[:, 3, :], w[:, :, 3])
    wg, bg = packing.interleave_glu(torch.arange(16.0)[:, None].repeat(1, 3), torch.arange(16.0), 8)
    assert bg.tolist() == [0, 1, 2, 3, 8, 9, 10, 11, 4, 5, 6, 7, 12, 13, 14, 15]
    w3 = packing.split_hi_lo(torch.randn(4, 8, generator=g))
    assert w3.shape == (4, 24) and w3.dtype == torch.float16


def test_label_tables_and_alignment():
    from wfl_asr_b200.pipeline import label_tables
    phon, kind, ph = label_tables(["B-a", "B-k", "I-a", "I-k", "O", "weird"])
    assert phon == ["a", "k"] and kind == [1, 1, 2, 2, 0, 3] and ph == [0, 1, 0, 1, -1, -1]
    from wfl_asr_b200.infer import align_phoneme_list, split_audio
    pred = [(0.0, 0.1, "a"), (0.1, 0.2, "x"), (0.2, 0.3, "k"), (0.3, 0.4, "s")]
    assert align_phoneme_list(pred, ["a", "k", "q"]) == [(0.0, 0.1, "a"), (0.2, 0.3, "k"), (0.1, 0.2, "q")]
    assert [len(s) for s in split_audio(np.zeros(16000 * 65), 16000)] == [480000, 480000, 80000]


def test_wav_reader(tmp_path):
    from wfl_asr_b200.infer import _read_wav
    x = (np.sin(np.arange(1000) * 0.05) * 20000).astype("<i2")
    p = tmp_path / "a.wav"
    with open(p, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + x.nbytes) + b"WAVEfmt " + struct.pack("<IHHIIHH", 16, 1, 1, 16000, 32000, 2, 16))
        f.write(b"data" + struct.pack("<I", x.nbytes) + x.tobytes())
    audio, sr = _read_wav(str(p))
    assert sr == 16000 and np.array_equal(audio, x.astype(np.float64) / 32768.0)


def test_cli_flags_and_error_paths(tmp_path):
    """python -m wfl_asr_b200.infer: the reference's flags (REF/infer.py:362-373), its sampling-flag validation
    messages with exit status 1 (REF/infer.py:377-392), and -- on a box without a GPU -- a loud refusal, not a CPU run."""
    from click.testing import CliRunner
    from wfl_asr_b200 import infer
    cmd = infer.cli_command()
    flags = {o for p in cmd.params for o in p.opts}
    assert {"--checkpoint", "-ckpt", "--config", "-c", "--output", "-o", "--lang-id", "-l", "--sample", "-s", "--top-k", "-tk",
            "--top-p", "-tp", "--temperature", "-temp", "--device", "-d", "--confidence-threshold", "-ct"} <= flags
    base = ["clip.wav", "-ckpt", "m.pt", "-c", "config.yaml"]
    run = CliRunner().invoke
    for extra, msg in ((["-s"], "neither --top-k nor --top-p"), (["-s", "-tk", "3", "-tp", "0.5"], "both --top-k and --top-p"),
                       (["-s", "-tp", "1.5"], "top-p must be between"), (["-s", "-tk", "3", "-temp", "0"], "temperature must be")):
        r = run(cmd, base + extra)
        assert r.exit_code == 1 and msg in r.output, (extra, r.output)
    assert run(cmd, ["clip.wav"]).exit_code == 2  # click: missing required --checkpoint / --config
    if not torch.cuda.is_available():
        r = run(cmd, base)
        assert r.exit_code == 1 and "no CPU path" in r.output


def test_bench_reference_arm_contract():
    """bench.py --impl reference (the CPU arm the driver times beside the GPU arm): runs without a GPU, prints ONE JSON
    line with the contract's keys, same metric/unit as the GPU arm, e2e == value, zero transfer bytes."""
    import json
    import subprocess
    import sys as _sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([_sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-sample-clips", "1"], capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "audio_seconds_labeled_per_second" and d["unit"] == "audio-s/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 == d["e2e"]["d2h_bytes_per_step"]
    assert "workload" in d["config"]


def test_model_rejects_shapes_the_reference_cannot_run():
    """d % conformer_heads != 0 fails in the reference's nn.MultiheadAttention (REF/model.py:26); even conv kernel
    sizes change the reference's frame alignment / length (REF/model.py:33,46-49,126-133) and are not built."""
    import copy

    import pytest

    from wfl_asr_b200 import synth
    from wfl_asr_b200.model import BIOPhonemeTagger
    labels = synth.synth_labels(4)
    base = synth.workload_config("cfg2")
    base["model"]["encoder_layers_override"] = 1
    for key, val in (("conformer_heads", 3), ("conformer_kernel_size", 30)):
        cfg = copy.deepcopy(base)
        cfg["model"][key] = val
        with pytest.raises(ValueError):
            BIOPhonemeTagger(cfg, labels)
    cfg = copy.deepcopy(base)
    cfg["model"].update(enable_dilated_conv=True, dilated_conv_kernel=4)
    with pytest.raises(ValueError):
        BIOPhonemeTagger(cfg, labels)
    cfg["model"].update(enable_dilated_conv=False)  # an unused even kernel is not an error
    BIOPhonemeTagger(cfg, labels)


def test_wav_header_parser_and_device_formats():
    """ingest.parse_wav_header: plain PCM, WAVE_FORMAT_EXTENSIBLE, odd-sized chunks before the data chunk, malformed
    files; ingest.device_pcm_format: which encodings wfl_pcm_to_f64 takes (the rest is decoded on the host)."""
    import io
    import struct

    import pytest

    from wfl_asr_b200 import ingest, ops

    def wav(tag, ch, sr, bits, pcm, extra=b"", extensible=False):
        if extensible:
            fmt = struct.pack("<HHIIHH", 0xFFFE, ch, sr, sr * ch * bits // 8, ch * bits // 8, bits) + struct.pack("<HHI", 22, bits, 0)
            fmt += struct.pack("<H", tag) + b"\x00" * 14
        else:
            fmt = struct.pack("<HHIIHH", tag, ch, sr, sr * ch * bits // 8, ch * bits // 8, bits)
        body = b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt + extra + b"data" + struct.pack("<I", len(pcm)) + pcm
        return b"RIFF" + struct.pack("<I", len(body)) + body

    pcm = bytes(range(24))
    tag, ch, sr, bits, off, size = ingest.parse_wav_header(io.BytesIO(wav(1, 2, 44100, 16, pcm)))
    assert (tag, ch, sr, bits, size) == (1, 2, 44100, 16, 24) and off == 44
    odd = b"LIST" + struct.pack("<I", 3) + b"abc\x00"  # odd-sized chunk is followed by a pad byte
    buf = wav(3, 1, 16000, 32, pcm, extra=odd)
    tag, ch, sr, bits, off, size = ingest.parse_wav_header(io.BytesIO(buf))
    assert (tag, ch, sr, bits, size) == (3, 1, 16000, 32, 24) and buf[off:off + size] == pcm
    tag, *_ = ingest.parse_wav_header(io.BytesIO(wav(1, 1, 8000, 24, pcm, extensible=True)))
    assert tag == 1
    for bad in (b"", b"RIFF\x00\x00\x00\x00WAVX", wav(1, 1, 8000, 16, pcm)[:30]):
        with pytest.raises(ValueError):
            ingest.parse_wav_header(io.BytesIO(bad))
    assert ingest.device_pcm_format(1, 16) == ops.PCM_S16 and ingest.device_pcm_format(1, 32) == ops.PCM_S32
    assert ingest.device_pcm_format(3, 32) == ops.PCM_F32
    assert ingest.device_pcm_format(1, 24) is None and ingest.device_pcm_format(1, 8) is None and ingest.device_pcm_format(3, 64) is None
