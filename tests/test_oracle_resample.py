"""CPU: the resampling oracle (numpy restatement of torchaudio.functional.resample) against outputs of torchaudio
itself (tests/golden/resample_golden.npz), and the product's host-side filter bank against the oracle's kernel."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_resample_golden as mrg  # noqa: E402
from oracle import resample_oracle as ro  # noqa: E402

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "resample_golden.npz"))


@pytest.mark.parametrize("sr,target,n", mrg.CASES)
def test_oracle_matches_torchaudio_golden(sr, target, n):
    y = ro.resample(mrg.signal(sr, n, sr + n), sr, target)
    g = GOLD[f"{sr}_{target}_{n}"]
    assert y.shape == g.shape
    assert np.abs(y - g).max() <= 1e-13  # fp64, summation order differs from torch's conv1d


def test_live_torchaudio_when_available():
    torchaudio = pytest.importorskip("torchaudio")
    import torch
    x = mrg.signal(24000, 5000, 7)
    want = torchaudio.functional.resample(torch.tensor(x), orig_freq=24000, new_freq=16000).numpy()
    assert np.abs(ro.resample(x, 24000, 16000) - want).max() <= 1e-13


@pytest.mark.parametrize("sr,target", [(44100, 16000), (48000, 16000), (8000, 16000), (22050, 16000)])
def test_product_filter_bank_equals_oracle_kernel(sr, target):
    from wfl_asr_b200.ingest import sinc_resample_bank
    bank, width, orig, new = sinc_resample_bank(sr, target)
    k, w2, o2, n2 = ro.sinc_kernel(sr, target)
    assert (width, orig, new) == (w2, o2, n2)
    assert np.abs(bank.numpy().T - k).max() <= 1e-15


def test_same_rate_is_identity():
    x = mrg.signal(16000, 100, 1)
    assert np.array_equal(ro.resample(x, 16000, 16000), x)
