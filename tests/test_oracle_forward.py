"""CPU: pins oracle/torch_oracle.py (the fp32 forward restatement that travels to the GPU box) to
(a) committed outputs of the unmodified reference model (tests/golden/forward_golden.npz) and
(b) when /root/reference is present, the live reference module on the same weights."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_forward_golden as mfg  # noqa: E402
from oracle import ref_loader, torch_oracle as to  # noqa: E402

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "forward_golden.npz"))


@pytest.mark.parametrize("name", list(mfg.CASES))
def test_oracle_matches_reference_fixture(name):
    cfg, labels, sd, wave, lang = mfg.case_inputs(name)
    logits, offsets = to.forward(wave, sd, cfg, lang)
    ref_l = torch.from_numpy(GOLD[name + "/logits"])
    ref_o = torch.from_numpy(GOLD[name + "/offsets"])
    scale = ref_l.abs().max().item()
    assert (logits[:, ::mfg.STRIDE] - ref_l).abs().max().item() <= 2e-5 * max(scale, 1.0)
    assert (offsets[:, ::mfg.STRIDE] - ref_o).abs().max().item() <= 2e-6
    agree = (logits.argmax(-1).numpy() == GOLD[name + "/argmax"]).mean()
    assert agree >= 0.999


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference only exists in the authoring container")
def test_oracle_matches_live_reference():
    name = "whisper_base_cfg2"
    cfg, labels, sd, wave, lang = mfg.case_inputs(name)
    ref = ref_loader.build_reference_model(cfg, labels, layer_override=cfg["model"]["encoder_layers_override"],
                                           randomize_bn=False)
    ref.load_state_dict(sd, strict=True)  # the oracle's key/shape table equals the reference's
    with torch.no_grad():
        rl, ro = ref(wave, None)  # lang_id=None skips lang_proj (REF/model.py:176)
    ol, oo = to.forward(wave, sd, cfg, None)
    assert (rl - ol).abs().max().item() <= 2e-5 * max(rl.abs().max().item(), 1.0)
    assert (ro - oo).abs().max().item() <= 2e-6


def test_mel_filters_match_transformers():
    tf_audio = pytest.importorskip("transformers.audio_utils")
    for n_mels in (80, 128):
        ref = tf_audio.mel_filter_bank(num_frequency_bins=201, num_mel_filters=n_mels, min_frequency=0.0,
                                       max_frequency=8000.0, sampling_rate=16000, norm="slaney", mel_scale="slaney")
        assert np.abs(to.slaney_mel_filters(n_mels) - ref.astype(np.float32)).max() <= 1e-7


def test_htk_filters_and_mel_power_match_torchaudio():
    """encoder_type "none" (REF/model.py:85-90): the oracle's restated filter bank / MelSpectrogram against torchaudio."""
    ta = pytest.importorskip("torchaudio")
    from wfl_asr_b200.frontend import htk_mel_filters
    for n_mels in (64, 80, 128):
        ref = ta.functional.melscale_fbanks(201, 0.0, 8000.0, n_mels, 16000, norm=None, mel_scale="htk")
        assert torch.equal(to.htk_mel_filters(n_mels), ref) and torch.equal(htk_mel_filters(n_mels), ref)
    ms = ta.transforms.MelSpectrogram(sample_rate=16000, n_fft=400, hop_length=320, n_mels=80)
    wave = torch.stack([torch.from_numpy(to.synth_wave(5 + i, 1.03)).float() for i in range(2)])
    ref = ms(wave).transpose(1, 2)
    got = to.mel_power(wave, ms.spectrogram.window, ms.mel_scale.fb, 320)
    assert got.shape == ref.shape == (2, 1 + wave.shape[1] // 320, 80)
    assert (got - ref).abs().max().item() <= 1e-6 * ref.abs().max().item()


def test_wavlm_frame_count():
    assert to.wavlm_num_frames(160000) == 499
    assert to.wavlm_num_frames(480000) == 1499
