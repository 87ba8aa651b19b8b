"""CPU: host side of the forward path (weight packing, buffer layouts, launch sequence of wfl_asr_b200.engine) run
over a torch model of the C-ABI's semantics (tests/ops_sim.py) and compared with the fp32 oracle.  This is a check
of the PACKING, not of the kernels -- the kernels are checked on the GPU box by the ``-m gpu`` tests.  It covers the
zero-padded layouts in particular (hidden size 80 of encoder_type "none": K blocks, attention heads 40 -> 64, BiLSTM
units 40 -> 192), which no BASELINE config exercises."""
import copy
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_forward_golden as mfg  # noqa: E402
import ops_sim  # noqa: E402
from oracle import torch_oracle as to  # noqa: E402
from wfl_asr_b200.engine import Engine  # noqa: E402

CPU = torch.device("cpu")


def _run(monkeypatch, cfg, sd, labels, wave, lang, **kw):
    ops_sim.install(monkeypatch)
    eng = Engine(sd, cfg, len(labels), CPU)
    logits, offsets = eng.forward(wave, lang, **kw)
    return eng, logits.clone(), offsets.clone()


def _check(logits, offsets, ref_l, ref_o, tol=2e-3, min_agree=0.99):
    scale = ref_l.abs().max().item()
    rel = (logits - ref_l).abs().max().item() / scale
    agree = (logits.argmax(-1) == ref_l.argmax(-1)).float().mean().item()
    off = (offsets - ref_o).abs().max().item()
    print(f"rel {rel:.3e} agree {agree:.4f} offsets {off:.3e}")
    assert rel <= tol and agree >= min_agree and off <= tol


@pytest.mark.parametrize("variant", ["full", "no_lang", "no_dilated", "conformer_only", "heads4_ffx4_mels64"])
def test_mel_none_engine_packing(monkeypatch, variant):
    cfg, labels, sd, wave, lang = mfg.case_inputs("mel_none_full")
    if variant == "no_lang":
        lang = None
    elif variant != "full":
        cfg = copy.deepcopy(cfg)
        if variant == "no_dilated":
            cfg["model"].update(enable_dilated_conv=False)
        elif variant == "conformer_only":
            cfg["model"].update(enable_bilstm=False, enable_dilated_conv=False)
        else:  # another hidden size (64: no K padding, head dim 16 -> 64, units 32 -> 192) and the default expansion
            cfg["data"]["n_mels"] = 64
            cfg["model"].update(conformer_heads=4, conformer_ff_expansion=4, bilstm_num_layer=1)
        sd = to.random_state_dict(cfg, len(labels), seed=21)
    eng, logits, offsets = _run(monkeypatch, cfg, sd, labels, wave, lang)
    ref_l, ref_o = to.forward(wave, sd, cfg, lang)
    assert logits.shape == ref_l.shape and offsets.shape == ref_o.shape
    if variant == "conformer_only":
        # Without the BiLSTM the Conformer's MHA runs on raw mel-power-scale activations (REF/model.py:42 has no norm in
        # front of it): at random init the attention logits have rms 240 / max 2100, and a softmax over logits of that
        # size is ill-conditioned for ANY 16-bit q/k (one fp16 ulp of q.k is ~1).  Layout check only.
        _check(logits, offsets, ref_l, ref_o, tol=5e-2, min_agree=0.9)
    else:
        _check(logits, offsets, ref_l, ref_o)


def test_mel_none_language_mean_and_max_label_len(monkeypatch):
    cfg, labels, sd, wave, lang = mfg.case_inputs("mel_none_full")
    ops_sim.install(monkeypatch)
    eng = Engine(sd, cfg, len(labels), CPU)
    outs = eng.forward_languages(wave, [0, 1])
    for lid, (lg, of) in enumerate(outs):
        lt = torch.full((wave.shape[0],), lid, dtype=torch.long)
        one_l, one_o = eng.forward(wave, lt)
        assert torch.equal(lg, one_l) and torch.equal(of, one_o)
    for mll in (90, 120):  # REF/model.py:166-174: truncate / zero-pad the hidden states to the label length
        lg, of = eng.forward(wave, lang, max_label_len=mll)
        ref_l, ref_o = to.forward(wave, sd, cfg, lang, max_label_len=mll)
        _check(lg.clone(), of.clone(), ref_l, ref_o)


def test_whisper_engine_through_the_simulator(monkeypatch):
    """Validates tests/ops_sim.py itself on a path the GPU tests pin with the real kernels (and keeps a CPU regression
    check on the Whisper packing: conv taps, BatchNorm fold, GLU interleave, BiLSTM row order, split-precision tail)."""
    cfg, labels, sd, wave, lang = mfg.case_inputs("whisper_base_full")
    wave, lang = wave[:1], lang[:1]
    eng, logits, offsets = _run(monkeypatch, cfg, sd, labels, wave, lang)
    ref_l, ref_o = to.forward(wave, sd, cfg, lang)
    _check(logits, offsets, ref_l, ref_o)


@pytest.mark.parametrize("name", ["wavlm_base_plus", "wavlm_large"])
def test_wavlm_engine_through_the_simulator(monkeypatch, name):
    """WavLM packing on the CPU: conv stack as paired-row stride-2 GEMMs, GroupNorm / LayerNorm variants, weight-norm
    fold, the positional conv as one grouped contraction (d/16 = 48 or 64 channels per group), gated relative-position
    bias tables, post-LN (base-plus) and pre-LN (large) layer orders."""
    cfg, labels, sd, wave, lang = mfg.case_inputs(name)
    wave, lang = wave[:1], lang[:1]
    eng, logits, offsets = _run(monkeypatch, cfg, sd, labels, wave, lang)
    ref_l, ref_o = to.forward(wave, sd, cfg, lang)
    _check(logits, offsets, ref_l, ref_o)


@pytest.mark.parametrize("name,use_lang,mll", [("wavlm_base_plus", True, 80), ("wavlm_base_plus", False, 120),
                                                ("wavlm_large", True, 90), ("whisper_base_full", False, 1400)])
def test_max_label_len_paths(monkeypatch, name, use_lang, mll):
    """REF/model.py:166-174 (the batched training/eval caller fixes T to the label length): hidden states truncated or
    zero-padded AFTER the encoder's final LayerNorm, for encoders with and without a final norm, with and without the
    language projection, into the BiLSTM or straight into the Conformer."""
    cfg, labels, sd, wave, lang = mfg.case_inputs(name)
    wave, lang = wave[:1], (lang[:1] if use_lang else None)
    ops_sim.install(monkeypatch)
    eng = Engine(sd, cfg, len(labels), CPU)
    lg, of = eng.forward(wave, lang, max_label_len=mll)
    ref_l, ref_o = to.forward(wave, sd, cfg, lang, max_label_len=mll)
    assert lg.shape == ref_l.shape and lg.shape[1] == mll
    _check(lg.clone(), of.clone(), ref_l, ref_o)


def test_conformer_head_dim_padded_to_a_built_size(monkeypatch):
    """conformer_heads 6 at d = 768 gives head_dim 128, which the attention kernels are not built for: the engine pads
    every head's q/k/v rows and out_proj columns to 256 (scale stays 1/sqrt(128))."""
    cfg, labels, sd, wave, lang = mfg.case_inputs("wavlm_base_plus")
    cfg = copy.deepcopy(cfg)
    cfg["model"]["conformer_heads"] = 6
    wave, lang = wave[:1], lang[:1]
    eng, logits, offsets = _run(monkeypatch, cfg, sd, labels, wave, lang)
    assert eng.conf_hdp == 256 and eng.conf_aw == 6 * 256
    ref_l, ref_o = to.forward(wave, sd, cfg, lang)
    _check(logits, offsets, ref_l, ref_o)


def test_sub_batched_forward_layout(monkeypatch):
    """WFL_SUB_BATCH: the batch in equal parts writing straight into the full-size outputs (bitwise equality with the
    single pass is a property of the real kernels and is asserted on the GPU; here: same values, right slices)."""
    cfg, labels, sd, wave, lang = mfg.case_inputs("mel_none_full")
    ops_sim.install(monkeypatch)
    eng = Engine(sd, cfg, len(labels), CPU)
    monkeypatch.setenv("WFL_SUB_BATCH", "0")
    one_l, one_o = (t.clone() for t in eng.forward(wave, lang))
    monkeypatch.setenv("WFL_SUB_BATCH", "1")
    l, o = eng.forward(wave, lang)
    # torch's CPU matmul is not batch-invariant and an fp16 rounding flip propagates: layout-level tolerance
    assert l.shape == one_l.shape and (l - one_l).abs().max().item() <= 1e-3 * one_l.abs().max().item()
    assert (o - one_o).abs().max().item() <= 1e-4
