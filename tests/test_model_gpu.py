"""GPU (-m gpu): the drop-in model / pipeline / infer entry point against the oracle.

Tolerances (BASELINE.json north_star): logits max error <= 1e-2 of the fp32 logit scale, frame-tag agreement
reported and asserted against a margin-aware bar, and -- given the tags and offsets the GPU produced --
segments and .lab text bit-exact with the reference's post-processing."""
import os
import struct
import sys

import numpy as np
import pytest
import torch
import yaml

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_forward_golden as mfg  # noqa: E402
from oracle import postproc_oracle as po  # noqa: E402
from oracle import torch_oracle as to  # noqa: E402
from wfl_asr_b200.model import BIOPhonemeTagger  # noqa: E402
from wfl_asr_b200.pipeline import Labeler  # noqa: E402

DEV = torch.device("cuda:0")
NORTH_STAR_TAG_AGREEMENT = 0.999  # BASELINE.json north_star: frame tag agreement >= 99.9 % (asserted on full-size clips)
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "forward_golden.npz"))
_SEL = [s for s in os.environ.get("WFL_TEST_CASES", "").split(",") if s]
SUPPORTED = [n for n in mfg.CASES if not _SEL or any(s in n for s in _SEL)]


def _build(name):
    cfg, labels, sd, wave, lang = mfg.case_inputs(name)
    model = BIOPhonemeTagger(cfg, labels)
    model.load_state_dict(sd, strict=True)
    return cfg, labels, sd, wave, lang, model.to(DEV).eval()


def _compare(name, logits, offsets, ref_l, ref_o):
    scale = ref_l.abs().max().item()
    err = (logits - ref_l).abs().max().item()
    agree = (logits.argmax(-1) == ref_l.argmax(-1)).float().mean().item()
    top2 = ref_l.topk(2, dim=-1).values
    margin = top2[..., 0] - top2[..., 1]
    safe = margin > 2 * err
    agree_safe = (logits.argmax(-1) == ref_l.argmax(-1))[safe].float().mean().item() if safe.any() else 1.0
    off_err = (offsets - ref_o).abs().max().item()
    print(f"[{name}] logits max err {err:.3e} / scale {scale:.3e} = {err / scale:.3e}; tag agreement {agree:.4%} "
          f"(frames with margin > 2*err: {safe.float().mean().item():.2%}, agreement there {agree_safe:.4%}); "
          f"offsets max err {off_err:.3e}")
    return err / scale, agree, agree_safe, off_err


@pytest.mark.parametrize("name", SUPPORTED)
def test_forward_matches_oracle(name):
    cfg, labels, sd, wave, lang, model = _build(name)
    logits, offsets = model(wave.to(DEV), lang.to(DEV))
    logits, offsets = logits.float().cpu(), offsets.float().cpu()
    ref_l, ref_o = to.forward(wave, sd, cfg, lang)
    # the oracle itself is pinned to the reference by tests/test_oracle_forward.py; cross-check the fixture too
    assert (ref_l[:, ::mfg.STRIDE] - torch.from_numpy(GOLD[name + "/logits"])).abs().max().item() <= 2e-5 * max(1.0, ref_l.abs().max().item())
    rel, agree, agree_safe, off_err = _compare(name, logits, offsets, ref_l, ref_o)
    # north_star tolerance: 1e-2 of the logit scale.  fp16 operands + fp32 accumulation measure 3e-4 .. 1e-3, so the
    # test holds the implementation to 2e-3; tag agreement target 99.9 % (measured 99.90 .. 100 %; the bar below
    # leaves room for one or two near-tie frames at these small test sizes)
    assert rel <= 2e-3, f"logit error {rel} above the fp16 bar 2e-3 (north_star tolerance 1e-2)"
    assert agree_safe == 1.0
    # north_star bar 99.9 %, asserted as such on the full-size clips (test_full_size_*); these fixtures are one or two
    # short clips of random-init logits (about 1 % of their frames have a top-2 margin below twice the logit error), so
    # the raw bar here leaves room for two near-tie frames -- every frame outside that margin must agree (above)
    frames = ref_l.shape[0] * ref_l.shape[1]
    mismatched = int((logits.argmax(-1) != ref_l.argmax(-1)).sum())
    assert mismatched <= max(2, int(frames * (1.0 - NORTH_STAR_TAG_AGREEMENT))), f"{mismatched} of {frames} frames differ"
    assert off_err <= 2e-3


def test_lang_none_skips_projection():
    name = "whisper_base_cfg2"
    cfg, labels, sd, wave, lang, model = _build(name)
    logits, offsets = model(wave.to(DEV), None)
    ref_l, ref_o = to.forward(wave, sd, cfg, None)
    rel, agree, agree_safe, _ = _compare(name + "/lang=None", logits.float().cpu(), offsets.float().cpu(), ref_l, ref_o)
    assert rel <= 2e-3 and agree_safe == 1.0


@pytest.mark.parametrize("name", [n for n in ("whisper_base_full", "wavlm_base_plus", "mel_none_full") if n in SUPPORTED])
def test_language_mean_fast_path(name):
    """REF/infer.py:265-276 (no --lang-id): mean over per-language forwards.  The fast path runs the encoder once and
    must equal the per-language full forwards bit for bit, and the fp32 oracle's mean within the logit tolerance."""
    cfg, labels, sd, wave, lang, model = _build(name)
    ids = list(range(cfg["model"]["num_languages"]))
    B = wave.shape[0]
    lm, om = model.forward_language_mean(wave.to(DEV), ids)
    per = []
    for i in ids:
        lg, of = model(wave.to(DEV), torch.full((B,), i, dtype=torch.long, device=DEV))
        per.append((lg.clone(), of.clone()))
    assert torch.equal(lm, torch.stack([p[0] for p in per]).mean(dim=0))
    assert torch.equal(om, torch.stack([p[1] for p in per]).mean(dim=0))
    refs = [to.forward(wave, sd, cfg, torch.full((B,), i, dtype=torch.long)) for i in ids]
    ref_l = torch.stack([r[0] for r in refs]).mean(dim=0)
    ref_o = torch.stack([r[1] for r in refs]).mean(dim=0)
    rel, agree, agree_safe, off_err = _compare(name + "/lang-mean", lm.float().cpu(), om.float().cpu(), ref_l, ref_o)
    assert rel <= 2e-3 and agree_safe == 1.0 and off_err <= 2e-3


@pytest.mark.parametrize("name,use_lang,mll", [("wavlm_base_plus", True, 80), ("wavlm_base_plus", False, 120),
                                                ("wavlm_large", True, 90), ("whisper_base_full", False, 1400),
                                                ("whisper_base_full", True, 1600), ("mel_none_full", True, 90)])
def test_max_label_len(name, use_lang, mll):
    """REF/model.py:166-174: the batched training/eval caller (REF/train.py:485-495) fixes T to the label length --
    hidden states truncated or zero-padded after the encoder, before lang_proj / BiLSTM / Conformer."""
    if name not in SUPPORTED:
        pytest.skip("case filtered out")
    cfg, labels, sd, wave, lang, model = _build(name)
    lang_d = lang.to(DEV) if use_lang else None
    logits, offsets = model(wave.to(DEV), lang_d, max_label_len=mll)
    ref_l, ref_o = to.forward(wave, sd, cfg, lang if use_lang else None, max_label_len=mll)
    assert tuple(logits.shape) == tuple(ref_l.shape) and logits.shape[1] == mll
    rel, agree, agree_safe, off_err = _compare(f"{name}/mll{mll}", logits.float().cpu(), offsets.float().cpu(), ref_l, ref_o)
    assert rel <= 2e-3 and agree_safe == 1.0 and off_err <= 2e-3


@pytest.mark.parametrize("name,heads", [("wavlm_base_plus", 6), ("whisper_base_cfg2", 4)])
def test_conformer_head_dim_128_runs_padded(name, heads):
    """Head dims the attention kernels are not built for (128: conformer_heads 4 at d 512, the reference's default
    head count; 6 at d 768) run with every head zero-padded to the next built size (256) -- same numbers, same bar."""
    if name not in SUPPORTED:
        pytest.skip("case filtered out")
    import copy
    cfg, labels, _, wave, lang = mfg.case_inputs(name)
    cfg = copy.deepcopy(cfg)
    cfg["model"]["conformer_heads"] = heads
    sd = to.random_state_dict(cfg, len(labels), seed=31)
    model = BIOPhonemeTagger(cfg, labels)
    model.load_state_dict(sd, strict=True)
    model = model.to(DEV).eval()
    logits, offsets = model(wave.to(DEV), lang.to(DEV))
    assert model.engine().conf_hdp == 256
    ref_l, ref_o = to.forward(wave, sd, cfg, lang)
    rel, agree, agree_safe, off_err = _compare(f"{name}/heads{heads}", logits.float().cpu(), offsets.float().cpu(), ref_l, ref_o)
    assert rel <= 2e-3 and agree_safe == 1.0 and off_err <= 2e-3


def test_sub_batched_forward_is_bitwise_identical(monkeypatch):
    """engine.forward may push a large batch through in equal sub-batches (L2-resident working set);
    clips are independent and the kernels batch-invariant, so the result must equal the single pass bit for bit."""
    cfg, labels, sd, wave, lang, model = _build("whisper_base_cfg2")
    wave4 = torch.cat([wave, wave.flip(0) * 0.5, wave * 0.25, wave.flip(0)], dim=0)[:4].contiguous().to(DEV)
    lang4 = torch.tensor([0, 1, 1, 0], device=DEV)
    monkeypatch.setenv("WFL_SUB_BATCH", "0")
    one_l, one_o = (t.clone() for t in model(wave4, lang4))
    for step in ("2", "1"):
        monkeypatch.setenv("WFL_SUB_BATCH", step)
        l, o = model(wave4, lang4)
        assert torch.equal(l, one_l) and torch.equal(o, one_o), f"sub-batches of {step} differ from the single pass"
    monkeypatch.setenv("WFL_SUB_BATCH", "2")
    l, o = model(wave4, None)
    monkeypatch.setenv("WFL_SUB_BATCH", "0")
    l1, o1 = model(wave4, None)
    assert torch.equal(l, l1) and torch.equal(o, o1)


def test_cpu_input_fails_loudly():
    cfg, labels, sd, wave, lang, model = _build("whisper_base_cfg2")
    with pytest.raises(RuntimeError):
        model(wave, lang)
    with pytest.raises(RuntimeError):
        BIOPhonemeTagger(cfg, labels)(wave, lang)  # model never moved to the GPU


@pytest.mark.parametrize("median_k,mode,thr", [(1, "right", 0.0), (5, "right", 0.5), (2, "previous", 0.3), (3, "left", 0.02),
                                               (7, "none", 0.0)])
def test_pipeline_lab_bit_exact_given_gpu_logits(median_k, mode, thr):
    """Feed the GPU's own logits/offsets to the reference post-processing restatement: ids, segments (fp64) and
    .lab text must be identical."""
    name = "whisper_base_cfg2"
    cfg, labels, sd, wave, lang, model = _build(name)
    wave2 = torch.cat([wave, wave.flip(1) * 0.7], dim=0).to(DEV)
    lang2 = torch.tensor([0, 1], device=DEV)
    logits, offsets = model(wave2, lang2)
    lab = Labeler(model, median_filter=median_k, merge_mode=mode, confidence_threshold=thr)
    ids, merged, nout, fcb, n_files = lab.postprocess(logits, offsets)
    got = lab.fetch(merged, nout, fcb, n_files, logits.shape[1])
    lg, of = logits.float().cpu().numpy(), offsets.float().cpu().numpy()
    mismatched_frames = 0
    for b in range(2):
        ref_ids = po.suppress_low_confidence_ids(lg[b], labels.index("O"), thr)
        if median_k > 1:
            ref_ids = po.median_filter_ids(ref_ids, median_k)
        gpu_ids = ids[b].cpu().numpy()
        mismatched_frames += int((gpu_ids != ref_ids).sum())
        # "given identical tag sequences": run the oracle decode on the GPU's ids
        segs = po.decode_bio_tags([labels[i] for i in gpu_ids], 0.02, of[b])
        segs = po.merge_adjacent_segments(segs, mode)
        assert got[b] == segs
        from wfl_asr_b200.utils import htk_lines
        assert htk_lines(got[b]) == "".join(po.lab_lines(segs))
    assert mismatched_frames <= 2  # softmax rounding may flip a frame that sits exactly on the threshold


class _LabelStub(torch.nn.Module):
    """What pipeline.Labeler needs from a model when only post-processing runs: the label tables and a device."""

    def __init__(self, labels):
        super().__init__()
        self.label_list = list(labels)
        self.label2id = {t: i for i, t in enumerate(labels)}
        self.id2label = dict(enumerate(labels))
        self.anchor = torch.nn.Parameter(torch.zeros(1))


def test_canonical_to_lang_remap_and_merge_golden():
    """REF/infer.py:303-310: decoded segments are renamed through phoneme_merge_map.json for the requested language and
    THEN merged, so two different model phonemes that map to one output name merge.  Here the remap is the ph_class
    table of wfl_merge_segments (pipeline.Labeler.set_output_names); segments and .lab text must equal what the
    reference's decode_bio_tags -> canonical_to_lang -> merge_adjacent_segments -> save_lab produced
    (tests/golden/align_golden.json.gz, reference-generated)."""
    import gzip
    import json
    from wfl_asr_b200 import utils
    with gzip.open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "align_golden.json.gz"), "rt") as f:
        gold = json.load(f)
    labels, mm = gold["labels"], gold["merge_map"]
    stub = _LabelStub(labels).to(DEV)
    merged_two = 0
    for rec in gold["remap"]:
        T = len(rec["tags"])
        ids = torch.tensor([labels.index(t) for t in rec["tags"]])
        logits = torch.nn.functional.one_hot(ids, len(labels)).float()[None].mul(12.0).to(DEV)
        offsets = torch.tensor(rec["offsets"], dtype=torch.float32)[None].to(DEV)
        for mode in ("right", "left", "previous", "none"):
            lab = Labeler(stub, median_filter=1, merge_mode=mode, confidence_threshold=0.0)
            lab.set_output_names([utils.canonical_to_lang(p, rec["lang"], mm) for p in lab.phon])
            _, merged, nout, fcb, n_files = lab.postprocess(logits, offsets)
            got = lab.fetch(merged, nout, fcb, n_files, T)[0]
            want = [tuple(r) for r in rec["merged"][mode]]
            assert got == want, (rec["lang"], mode, T)
            assert utils.htk_lines(got) == rec["lab"][mode]
        plain = Labeler(stub, median_filter=1, merge_mode="right", confidence_threshold=0.0)
        _, merged, nout, fcb, n_files = plain.postprocess(logits, offsets)
        merged_two += len(plain.fetch(merged, nout, fcb, n_files, T)[0]) > len(rec["merged"]["right"])
    assert merged_two >= 5  # the remap really merged segments that stay apart without it


def test_utils_dropins_match_oracle(golden):
    from wfl_asr_b200 import utils
    for rec in golden["decode"][:12]:
        off = np.asarray(rec["offsets"], dtype=np.float32) if rec["offsets"] is not None else None
        segs = utils.decode_bio_tags(rec["tags"], 0.02, torch.from_numpy(off) if off is not None else None)
        assert segs == [(s, e, p) for s, e, p in rec["segments"]]
        for mode in ("right", "left", "previous", "none"):
            assert utils.merge_adjacent_segments(list(segs), mode) == [(s, e, p) for s, e, p in rec["merged"][mode]]
            assert utils.htk_lines(utils.merge_adjacent_segments(list(segs), mode)) == rec["lab"][mode]
    with pytest.raises(ValueError):
        utils.merge_adjacent_segments([(0.0, 1.0, "a")], "sideways")


def test_label_stream_matches_synchronous_label():
    """The pipelined host-buffer API (copy stream + double buffering) returns exactly what the synchronous call does."""
    cfg, labels, sd, wave, lang, model = _build("whisper_base_cfg2")
    lab = Labeler(model, median_filter=3, merge_mode="right", confidence_threshold=0.0)
    batches = [torch.cat([wave, wave.flip(1) * s], 0).pin_memory() for s in (0.9, 0.5, 0.7)]
    lang2 = torch.tensor([0, 1], device=DEV)
    want = [lab.label(b.to(DEV), lang2) for b in batches]
    got = list(lab.label_stream(iter(batches), lang2))
    assert got == want and sum(len(s) for batch in got for s in batch) > 0


def test_graph_replay_matches_direct_launches():
    """label_host replays the whole pass from a CUDA graph: same segments as direct launches, also when the input
    values and the language ids change between replays of one captured graph."""
    cfg, labels, sd, wave, lang, model = _build("whisper_base_full")
    direct = Labeler(model, median_filter=3, merge_mode="right", confidence_threshold=0.2, use_graphs=False)
    graphed = Labeler(model, median_filter=3, merge_mode="right", confidence_threshold=0.2, use_graphs=True)
    batch0 = torch.cat([wave, wave.flip(1) * 0.8], 0).pin_memory()
    batch1 = torch.cat([wave.flip(1) * 0.6, wave * 0.9], 0).pin_memory()
    n = batch0.shape[0]
    for batch, lg in ((batch0, [0, 1]), (batch1, [1, 0]), (batch0, [1, 1])):
        lt = torch.tensor((lg * n)[:n], device=DEV)
        assert graphed.label_host(batch, lt) == direct.label_host(batch, lt)
    assert len(graphed._graphs) == 1 and not direct._graphs


def test_graph_replay_survives_workspace_eviction():
    """The engine / labeler workspace caches keep one shape each.  Alternating two shapes through label_host must not
    let a replayed graph touch the evicted (freed) workspaces of its shape: each captured pass owns references to
    what it captured.  Shapes A, B, A, B ... with graphs on must equal direct launches, also with allocator churn
    between the calls (so freed blocks would be handed out again if the graph did not hold them)."""
    cfg, labels, sd, wave, lang, model = _build("wavlm_base_plus")
    direct = Labeler(model, median_filter=3, merge_mode="right", confidence_threshold=0.1, use_graphs=False)
    graphed = Labeler(model, median_filter=3, merge_mode="right", confidence_threshold=0.1, use_graphs=True)
    a = wave.clone().pin_memory()                       # [2, 32000]
    b = torch.cat([wave, wave.flip(0)], 0)[:3, :24000].contiguous().pin_memory()  # other batch size AND length
    la, lb = torch.tensor([0, 1], device=DEV), torch.tensor([1, 0, 1], device=DEV)
    want_a, want_b = direct.label_host(a, la), direct.label_host(b, lb)
    junk = []
    for r in range(3):
        assert graphed.label_host(a, la) == want_a, f"round {r}: shape A differs after shape B evicted its workspaces"
        junk.append(torch.randn(1 << 22, device=DEV))  # churn: reuse whatever the caches freed
        assert graphed.label_host(b, lb) == want_b, f"round {r}: shape B differs"
        junk.append(torch.randn(1 << 22, device=DEV))
    assert len(graphed._graphs) == 2


def test_forward_returns_fresh_tensors():
    """REF/infer.py:268-275 keeps one logits tensor per language and averages afterwards: ``forward`` must return
    tensors that a later forward does not overwrite (the engine's workspace views stay internal)."""
    cfg, labels, sd, wave, lang, model = _build("whisper_base_cfg2")
    x = wave.to(DEV)
    B = x.shape[0]
    l0, o0 = model(x, torch.zeros(B, dtype=torch.long, device=DEV))
    keep_l, keep_o = l0.clone(), o0.clone()
    l1, o1 = model(x, torch.ones(B, dtype=torch.long, device=DEV))
    assert l0.is_contiguous() and l0.data_ptr() != l1.data_ptr()
    assert torch.equal(l0, keep_l) and torch.equal(o0, keep_o)
    assert not torch.equal(l0, l1)  # the two languages really differ
    mean = torch.stack([l0, l1]).mean(dim=0)
    lm, _ = model.forward_language_mean(x, [0, 1])
    assert torch.equal(mean, lm)


def _write_wav(path, x, sr=16000):
    pcm = (np.clip(x, -1, 1) * 32767.0).astype("<i2").tobytes()
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + len(pcm)) + b"WAVEfmt " + struct.pack("<IHHIIHH", 16, 1, 1, sr, sr * 2, 2, 16))
        f.write(b"data" + struct.pack("<I", len(pcm)) + pcm)


@pytest.mark.parametrize("seconds,file_sr", [(4.0, 16000), (47.3, 16000), (5.0, 44100)])
def test_infer_audio_end_to_end(tmp_path, seconds, file_sr):
    """infer.py entry point: config.yaml + phonemes.txt + langs.txt + checkpoint + wav -> .lab, against the oracle
    chain (peak normalise, <=30 s chunks, forward, threshold, median, decode, shift, merge, save_lab)."""
    from wfl_asr_b200 import infer
    name = "whisper_base_cfg2"
    cfg, labels, sd, _, _ = mfg.case_inputs(name)
    cfg["output"] = {"save_dir": str(tmp_path)}
    cfg["postprocess"] = {"median_filter": 3, "merge_segments": "right", "confidence_threshold": 0.1}
    (tmp_path / "phonemes.txt").write_text("\n".join(labels) + "\n")
    (tmp_path / "langs.txt").write_text("en,0\nja,1\n")
    with open(tmp_path / "config.yaml", "w") as f:
        yaml.safe_dump(cfg, f)
    torch.save(sd, tmp_path / "best_model.pt")
    x = to.synth_wave(77, seconds, sr=file_sr) * 0.8
    wav = tmp_path / "clip.wav"
    _write_wav(str(wav), x, sr=file_sr)
    out_lab = tmp_path / "out" / "clip.lab"
    segs = infer.infer_audio(str(wav), str(tmp_path / "config.yaml"), str(tmp_path / "best_model.pt"), str(out_lab),
                             device="cuda:0", lang_id=1, confidence_threshold=0.1, use_cache=False)
    text = out_lab.read_text()
    assert text == "".join(po.lab_lines(segs))
    # oracle chain on the same file
    audio, sr = infer.read_audio(str(wav))
    if sr != 16000:  # REF/infer.py:217-220
        from oracle import resample_oracle as ro
        audio, sr = ro.resample(audio, sr, 16000), 16000
    audio = po.peak_normalize(audio)
    chunks = [audio] if len(audio) / sr <= 30.0 else [audio[s:s + 480000] for s in range(0, len(audio), 480000)]
    all_segs, t, n_frames, n_agree = [], 0.0, 0, 0
    model = infer._Session.get(str(tmp_path / "config.yaml"), str(tmp_path / "best_model.pt"), "cuda:0").model
    for ch in chunks:
        if len(chunks) > 1:
            ch = po.peak_normalize(ch)
        w = torch.tensor(ch, dtype=torch.float32)[None]
        ref_l, ref_o = to.forward(w, sd, cfg, torch.tensor([1]))
        g_l, g_o = model(w.to(DEV), torch.tensor([1], device=DEV))
        g_l, g_o = g_l[0].float().cpu().numpy(), g_o[0].float().cpu().numpy()
        ids = po.suppress_low_confidence_ids(g_l, labels.index("O"), 0.1)
        ref_ids = po.suppress_low_confidence_ids(ref_l[0].numpy(), labels.index("O"), 0.1)
        n_frames += len(ids)
        n_agree += int((ids == ref_ids).sum())
        ids = po.median_filter_ids(ids, 3)
        s = po.decode_bio_tags([labels[i] for i in ids], 0.02, g_o)
        all_segs += po.shift_segments(s, t) if len(chunks) > 1 else s
        t += len(ch) / sr
    expect = po.merge_adjacent_segments(all_segs, "right")
    print(f"[infer {seconds}s] {len(segs)} segments; frame-tag agreement with the fp32 oracle {n_agree / n_frames:.4%}")
    assert segs == expect
    assert n_agree / n_frames >= 0.995


def test_wfl_cache_logits_cache(tmp_path):
    """REF/infer.py:222-232,246-249,278-280 (+ :120-131 for 30 s chunks): logits / offsets are cached beside the audio
    under .wfl_cache/ with the reference's file names and tensor shapes, and a second call labels from the cache."""
    from wfl_asr_b200 import infer
    cfg, labels, sd, _, _ = mfg.case_inputs("whisper_base_cfg2")
    cfg["output"] = {"save_dir": str(tmp_path)}
    cfg["postprocess"] = {"median_filter": 3, "merge_segments": "right", "confidence_threshold": 0.1}
    (tmp_path / "phonemes.txt").write_text("\n".join(labels) + "\n")
    (tmp_path / "langs.txt").write_text("en,0\nja,1\n")
    with open(tmp_path / "config.yaml", "w") as f:
        yaml.safe_dump(cfg, f)
    torch.save(sd, tmp_path / "best_model.pt")
    wavs = tmp_path / "w"
    wavs.mkdir()
    _write_wav(str(wavs / "short.wav"), to.synth_wave(5, 2.5) * 0.8)
    _write_wav(str(wavs / "long.wav"), to.synth_wave(6, 31.0) * 0.8)
    args = (str(tmp_path / "config.yaml"), str(tmp_path / "best_model.pt"))
    first = {n: infer.infer_audio(str(wavs / n), *args, None, device="cuda:0", lang_id=1, confidence_threshold=0.1)
             for n in ("short.wav", "long.wav")}
    cache = wavs / ".wfl_cache"
    names = sorted(p.name for p in cache.iterdir())
    assert names == ["long_seg0_lang1_logits.pt", "long_seg0_lang1_offsets.pt", "long_seg1_lang1_logits.pt",
                     "long_seg1_lang1_offsets.pt", "short_lang1_logits.pt", "short_lang1_offsets.pt"]
    lg = torch.load(cache / "short_lang1_logits.pt", weights_only=False)
    of = torch.load(cache / "short_lang1_offsets.pt", weights_only=False)
    assert tuple(lg.shape) == (1, 1500, len(labels)) and tuple(of.shape) == (1500, 2)  # the reference's shapes
    # second call: same segments, and it really reads the cache (poison one entry and see the labels change)
    again = infer.infer_audio(str(wavs / "short.wav"), *args, None, device="cuda:0", lang_id=1, confidence_threshold=0.1)
    assert again == first["short.wav"]
    torch.save(torch.zeros_like(lg), cache / "short_lang1_logits.pt")
    poisoned = infer.infer_audio(str(wavs / "short.wav"), *args, None, device="cuda:0", lang_id=1, confidence_threshold=0.1)
    assert poisoned != first["short.wav"]
    fresh = infer.infer_audio(str(wavs / "short.wav"), *args, None, device="cuda:0", lang_id=1, confidence_threshold=0.1,
                              use_cache=False)
    assert fresh == first["short.wav"]
    # the language-mean path caches under the "_avg" suffix
    infer.infer_audio(str(wavs / "short.wav"), *args, None, device="cuda:0", lang_id=None, confidence_threshold=0.1)
    assert (cache / "short_avg_logits.pt").exists()


@pytest.mark.parametrize("name", ["whisper_base_cfg2", "wavlm_base_plus", "mel_none_full"])
def test_infer_folder_batched_equals_per_file(tmp_path, name, capsys):
    """infer_folder labels the folder in shared batches; every .lab must equal what infer_audio writes for the same
    file alone (long file -> 30 s chunks, 44.1 kHz file -> resampled, forced phoneme list -> aligned), for the
    single-language and the language-mean path."""
    from wfl_asr_b200 import infer
    cfg, labels, sd, _, _ = mfg.case_inputs(name)
    cfg["output"] = {"save_dir": str(tmp_path)}
    cfg["postprocess"] = {"median_filter": 3, "merge_segments": "right", "confidence_threshold": 0.1}
    (tmp_path / "phonemes.txt").write_text("\n".join(labels) + "\n")
    (tmp_path / "langs.txt").write_text("en,0\nja,1\n")
    with open(tmp_path / "config.yaml", "w") as f:
        yaml.safe_dump(cfg, f)
    torch.save(sd, tmp_path / "best_model.pt")
    folder = tmp_path / "wavs"
    folder.mkdir()
    whisper = cfg["model"]["encoder_type"] == "whisper"
    specs = [("a.wav", 3.1, 16000), ("b.wav", 33.0 if whisper else 2.2, 16000), ("c.wav", 2.2, 44100), ("d.wav", 3.1, 16000)]
    for fn, secs, sr in specs:
        _write_wav(str(folder / fn), to.synth_wave(hash(fn) % 97, secs, sr=sr) * 0.7, sr=sr)
    (folder / "d.txt").write_text("p3 p7 p1\n")
    args = (str(tmp_path / "config.yaml"), str(tmp_path / "best_model.pt"))
    for lang in (1, None):
        out_dir = tmp_path / f"out_{lang}"
        got = infer.infer_folder(str(folder), *args, output_dir=str(out_dir), device="cuda:0", lang_id=lang,
                                 confidence_threshold=0.1, files_per_pass=3, use_cache=False)
        assert set(got) == {s[0] for s in specs}
        for fn, _, _ in specs:
            single = tmp_path / f"single_{lang}_{fn}.lab"
            segs = infer.infer_audio(str(folder / fn), *args, str(single), device="cuda:0", lang_id=lang,
                                     confidence_threshold=0.1, use_cache=False)
            assert got[fn] == segs, f"{fn} (lang {lang}) differs between folder and single-file labeling"
            assert (out_dir / fn.replace(".wav", ".lab")).read_text() == single.read_text()
    capsys.readouterr()


@pytest.mark.parametrize("name,bucket", [("wavlm_base_plus", 8000), ("wavlm_base_plus", 1), ("whisper_base_cfg2", 8000),
                                         ("mel_none_full", 8000)])
def test_bulk_label_corpus_ragged(name, bucket):
    """bulk.label_corpus (BASELINE configs[3]: utterances of different lengths, length-bucketed): per bucket batch the
    logits match the fp32 oracle on the same zero-padded batch (the reference's batched caller pads without masks,
    REF/train.py:22-36), every utterance is decoded on its own frame count, and the segments equal the reference
    post-processing of the GPU's own tags and offsets exactly.  bucket=1 -> exact-length groups (per-file semantics)."""
    from wfl_asr_b200 import bulk, shard
    cfg, labels, sd, _, _, model = _build(name)
    secs = [0.61, 1.37, 0.62, 2.0, 1.36, 0.9, 1.37]
    waves = [to.synth_wave(300 + i, s).astype(np.float32) for i, s in enumerate(secs)]
    lens = [len(w) for w in waves]
    langs = [i % 2 for i in range(len(waves))]
    got = bulk.label_corpus(model, waves, langs, median_filter=3, merge_mode="right", confidence_threshold=0.1,
                            max_clips=3, bucket_samples=bucket)
    assert len(got) == len(waves)
    etype = cfg["model"]["encoder_type"]
    bsz = 480000 if etype == "whisper" else bucket
    seen = 0
    for padded, group in shard.plan_batches(lens, 1, etype, 3, 32 * 480000, bsz)[0]:
        host = torch.zeros(len(group), padded)
        for j, i in enumerate(group):
            host[j, :lens[i]] = torch.from_numpy(waves[i])
        lt = torch.tensor([langs[i] for i in group])
        g_l, g_o = model(host.to(DEV), lt.to(DEV))
        g_l, g_o = g_l.float().cpu(), g_o.float().cpu()
        ref_l, ref_o = to.forward(host, sd, cfg, lt)
        rel, agree, agree_safe, off_err = _compare(f"{name}/bulk{padded}", g_l, g_o, ref_l, ref_o)
        assert rel <= 2e-3 and agree_safe == 1.0
        for j, i in enumerate(group):
            n_fr = min(g_l.shape[1], shard.frames_for(lens[i], etype))
            ids = po.suppress_low_confidence_ids(g_l[j, :n_fr].numpy(), labels.index("O"), 0.1)
            ids = po.median_filter_ids(ids, 3)
            want = po.merge_adjacent_segments(po.decode_bio_tags([labels[k] for k in ids], 0.02, g_o[j, :n_fr].numpy()), "right")
            assert got[i] == want, f"utterance {i} ({lens[i]} samples) differs"
            seen += 1
    assert seen == len(waves)


def test_full_size_cfg2_properties():
    """BASELINE configs[1] at full size (whisper-base, 6 encoder layers + 4 Conformer, batch 32 x 30 s): size-independent
    properties -- run-to-run determinism (bitwise), batch invariance (clip i inside the batch == clip i alone, bitwise),
    an fp32-oracle check on one clip of the batch, and .lab idempotence of the merge."""
    from wfl_asr_b200 import synth
    cfg = synth.workload_config("cfg2")
    labels = synth.synth_labels(30)
    model = synth.bench_model(BIOPhonemeTagger, cfg, labels)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model = model.to(DEV).eval()
    B = 32
    clips = [synth.synth_wave(900 + i, 30.0 if i % 3 else 7.5 + 0.5 * i) for i in range(B)]  # mixed clip lengths
    wave = torch.from_numpy(np.stack([np.pad(w, (0, 480000 - len(w))) for w in clips]).astype(np.float32))
    lang = torch.tensor([i % 2 for i in range(B)], device=DEV)
    l1, o1 = model(wave.to(DEV), lang)
    l1, o1 = l1.clone(), o1.clone()
    l2, o2 = model(wave.to(DEV), lang)
    assert torch.equal(l1, l2) and torch.equal(o1, o2), "forward is not deterministic"
    for i in (0, 13, 31):
        li, oi = model(wave[i:i + 1].to(DEV), lang[i:i + 1])
        assert torch.equal(li[0], l1[i]) and torch.equal(oi[0], o1[i]), f"clip {i} differs between batch 32 and batch 1"
    chk = [1, 4, 7, 10, 13, 16, 19, 22]  # eight full 30 s clips: 12 000 frames against the fp32 oracle
    ref_l, ref_o = to.forward(wave[chk], sd, cfg, lang[chk].cpu())
    rel, agree, agree_safe, off_err = _compare(f"cfg2 full size, clips {chk}", l1[chk].float().cpu(), o1[chk].float().cpu(), ref_l, ref_o)
    assert rel <= 2e-3 and agree_safe == 1.0 and agree >= NORTH_STAR_TAG_AGREEMENT
    lab = Labeler(model, median_filter=5, merge_mode="right", confidence_threshold=0.5)
    ids, merged, nout, fcb, n_files = lab.postprocess(l1, o1)
    segs = lab.fetch(merged, nout, fcb, n_files, 1500)
    assert sum(len(s) for s in segs) > 0
    for b in (0, 13):
        want = po.merge_adjacent_segments(po.decode_bio_tags([labels[k] for k in ids[b].cpu().numpy()], 0.02,
                                                             o1[b].float().cpu().numpy()), "right")
        assert segs[b] == want
        assert po.merge_adjacent_segments(list(segs[b]), "right") == segs[b]  # merging is idempotent


def test_full_size_cfg3_properties():
    """BASELINE configs[2] at full size (whisper-small + BiLSTM(2) + 4 Conformer + dilated stack, batch 64 x 30 s):
    determinism, batch invariance (the batch-64 pass runs the BiLSTM with 16 clips per cluster and the GEMMs as CTA
    pairs, the batch-1 pass with 8 per cluster and single CTAs -- same bits), and an fp32-oracle check on one clip."""
    from wfl_asr_b200 import synth
    cfg = synth.workload_config("cfg3")
    labels = synth.synth_labels(30)
    model = synth.bench_model(BIOPhonemeTagger, cfg, labels)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model = model.to(DEV).eval()
    B = 64
    base = [synth.synth_wave(700 + i, 30.0) for i in range(4)]
    wave = torch.from_numpy(np.stack([base[i % 4] * (0.5 + 0.5 * ((i * 7) % 11) / 11.0) for i in range(B)]).astype(np.float32))
    lang = torch.tensor([i % 2 for i in range(B)], device=DEV)
    l1, o1 = model(wave.to(DEV), lang)
    l1, o1 = l1.clone(), o1.clone()
    l2, o2 = model(wave.to(DEV), lang)
    assert torch.equal(l1, l2) and torch.equal(o1, o2), "forward is not deterministic"
    for i in (0, 41, 63):
        li, oi = model(wave[i:i + 1].to(DEV), lang[i:i + 1])
        assert torch.equal(li[0], l1[i]) and torch.equal(oi[0], o1[i]), f"clip {i} differs between batch 64 and batch 1"
    chk = [0, 21, 41, 62]  # 6 000 frames against the fp32 oracle
    ref_l, ref_o = to.forward(wave[chk], sd, cfg, lang[chk].cpu())
    rel, agree, agree_safe, off_err = _compare(f"cfg3 full size, clips {chk}", l1[chk].float().cpu(), o1[chk].float().cpu(), ref_l, ref_o)
    assert rel <= 2e-3 and agree_safe == 1.0 and agree >= NORTH_STAR_TAG_AGREEMENT


def _full_size(workload, B, seconds, probe, check, seed0):
    """Full-depth model of a BASELINE config on B clips: run-to-run determinism, batch invariance of the probed clips
    (bitwise), and the fp32 oracle on the ``check`` clips at the north_star bars: logits within 1e-2 (held to 2e-3),
    frame-tag agreement >= 99.9 % over all checked frames."""
    from wfl_asr_b200 import synth
    cfg = synth.workload_config(workload)
    labels = synth.synth_labels(30)
    model = synth.bench_model(BIOPhonemeTagger, cfg, labels)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model = model.to(DEV).eval()
    base = [synth.synth_wave(seed0 + i, seconds) for i in range(min(B, 8))]
    wave = torch.from_numpy(np.stack([base[i % len(base)] * (0.5 + 0.5 * ((i * 7) % 11) / 11.0) for i in range(B)]).astype(np.float32))
    lang = torch.tensor([i % 2 for i in range(B)], device=DEV)
    l1, o1 = model(wave.to(DEV), lang)
    l2, o2 = model(wave.to(DEV), lang)
    assert torch.equal(l1, l2) and torch.equal(o1, o2), "forward is not deterministic"
    for i in probe:
        li, oi = model(wave[i:i + 1].to(DEV), lang[i:i + 1])
        assert torch.equal(li[0], l1[i]) and torch.equal(oi[0], o1[i]), f"clip {i} differs between batch {B} and batch 1"
    check = list(check)
    ref_l, ref_o = to.forward(wave[check], sd, cfg, lang[check].cpu())
    rel, agree, agree_safe, off_err = _compare(f"{workload} full size, clips {check}", l1[check].float().cpu(),
                                               o1[check].float().cpu(), ref_l, ref_o)
    assert rel <= 2e-3 and agree_safe == 1.0 and off_err <= 2e-3
    assert agree >= NORTH_STAR_TAG_AGREEMENT, f"{workload}: tag agreement {agree:.4%} over {ref_l.shape[0] * ref_l.shape[1]} frames"
    return model, labels, l1, o1


def test_full_size_cfg1_properties():
    """BASELINE configs[0] at full size: WavLM-base-plus (12 layers, GroupNorm front-end, post-LN) + 2 Conformer blocks
    (heads 2 -> head dim 384), 10 s clips -> 499 frames; a batch of 8 (3 992 frames against the oracle) that also
    checks batch invariance against the batch-1 pass the config names."""
    model, labels, l1, o1 = _full_size("cfg1", 8, 10.0, probe=(0, 7), check=range(8), seed0=500)
    assert l1.shape[1] == 499
    lab = Labeler(model, median_filter=5, merge_mode="right", confidence_threshold=0.5)
    ids, merged, nout, fcb, n_files = lab.postprocess(l1, o1)
    segs = lab.fetch(merged, nout, fcb, n_files, l1.shape[1])
    for b in range(3):
        want = po.merge_adjacent_segments(po.decode_bio_tags([labels[k] for k in ids[b].cpu().numpy()], 0.02,
                                                             o1[b].float().cpu().numpy()), "right")
        assert segs[b] == want


def test_full_size_cfg4_properties():
    """BASELINE configs[3] at full size: WavLM-large (24 pre-LN layers, LayerNorm front-end, input normalisation) + 6
    Conformer blocks with heads 2 -> head dim 512 (attention_big_kernel<512>), d = 1024; 6 clips x 16 s (the corpus'
    mean utterance length) -> 799 frames, 4 of them against the oracle."""
    model, labels, l1, o1 = _full_size("cfg4", 6, 16.0, probe=(0, 5), check=(0, 2, 3, 5), seed0=520)
    assert l1.shape[1] == 799 and model.engine().conf_hdp == 512


def test_full_size_cfg5_properties():
    """BASELINE configs[4] at full size: Whisper-large-v3 (32 layers, d 1280, 20 heads, 128 mels) + BiLSTM(2) with H 640
    (lstm_kernel<640,16,8>) + 8 Conformer blocks with head dim 640 (attention_big_kernel<640>) + dilated stack."""
    model, labels, l1, o1 = _full_size("cfg5", 4, 30.0, probe=(1,), check=(1, 2), seed0=540)
    assert l1.shape[1] == 1500 and model.engine().conf_hdp == 640 and model.engine().lstm_hp == 640


def test_encoder_run_to_run_determinism_soak():
    """Six passes of a whisper-small encoder-only model (d 768: the persistent GEMMs' last round is more than half
    full, the shape on which a barrier-phase race in the attention kernel once made the last clips of a batch differ
    run to run) must be bitwise identical."""
    from wfl_asr_b200 import synth
    cfg = synth.workload_config("cfg3")
    cfg["model"].update(enable_bilstm=False, enable_dilated_conv=False, num_conformer_layers=0)
    model = synth.bench_model(BIOPhonemeTagger, cfg, synth.synth_labels(30)).to(DEV).eval()
    B = 32
    base = [synth.synth_wave(700 + i, 30.0) for i in range(4)]
    wave = torch.from_numpy(np.stack([base[i % 4] * (0.5 + 0.5 * ((i * 7) % 11) / 11.0) for i in range(B)]).astype(np.float32)).to(DEV)
    lang = torch.tensor([i % 2 for i in range(B)], device=DEV)
    first = None
    for r in range(6):
        l, o = model(wave, lang)
        if first is None:
            first = (l.clone(), o.clone())
        else:
            bad = (l != first[0]).flatten(1).any(dim=1).nonzero().flatten().tolist()
            assert not bad, f"pass {r} differs from pass 0 in clips {bad}"
            assert torch.equal(o, first[1])


def test_correct_label_process_file_end_to_end(tmp_path):
    """correct_label.process_file (REF/correct_label.py:157-184): wav + .lab -> boundaries detected on the GPU -> snapped
    .lab, against the oracle chain (restated detector + the reference-pinned snapping), including the pre-made
    ``_boundary.txt`` path and its clean-up."""
    from oracle import correct_label_oracle as co
    from wfl_asr_b200 import correct_label as cl
    y = (to.synth_wave(91, 6.0) * 0.8)
    y[30000:31000] *= 0.02
    wav = tmp_path / "utt.wav"
    _write_wav(str(wav), y)
    pcm = (np.clip(y, -1, 1) * 32767.0).astype("<i2").astype(np.float32) / 32768.0  # what the file holds
    lab_in = "0 9000000 a\n9000000 19500000 b\n19500000 41000000 c\n41000000 60000000 d\n"
    (tmp_path / "utt.lab").write_text(lab_in)
    cl.process_file(str(wav))
    want_pred, *_ = co.detect_boundaries(pcm, 16000)
    segs = [(float(a) / 1e7, float(b) / 1e7, c) for a, b, c in (ln.split() for ln in lab_in.splitlines())]
    want = "".join(f"{int(s * 1e7)} {int(e * 1e7)} {l}\n" for s, e, l in co.snap_segments(segs, want_pred))
    got = (tmp_path / "utt.lab").read_text()
    assert got == want and got != lab_in and not (tmp_path / "utt_boundary.txt").exists()
    # a pre-made boundary file is used instead of the detector, then removed
    (tmp_path / "utt.lab").write_text(lab_in)
    (tmp_path / "utt_boundary.txt").write_text("0.910000\n1.940000\n")
    cl.process_file(str(wav))
    # each predicted boundary is used once (REF/correct_label.py:48,66,79): a's end takes 0.91, so b's start stays
    assert (tmp_path / "utt.lab").read_text() == "0 9100000 a\n9000000 19400000 b\n19500000 41000000 c\n41000000 60000000 d\n"
    assert not (tmp_path / "utt_boundary.txt").exists()
