"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the resampling the reference applies to non-16 kHz files.

Only ``tests/`` may import this module; the product path (``wfl_asr_b200.ingest`` + ``csrc/resample.cu``) never does.

Restates REF/infer.py:217-220 -> ``torchaudio.functional.resample(torch.tensor(audio), orig_freq=sr, new_freq=16000)``
on the float64 waveform (torchaudio is a third-party dependency of the reference, pinned 2.6.0 in
REF/requirements.txt:14, installed 2.11.0 here): TORCHAUDIO/functional/functional.py ``_get_sinc_resample_kernel``
(Hann-windowed sinc, lowpass_filter_width 6, rolloff 0.99) and ``_apply_sinc_resample_kernel`` (zero pad by
(width, width + orig), strided correlation, truncate to ceil(new * n / orig)).
Parity pin: ``tests/golden/resample_golden.npz`` holds outputs of torchaudio itself (``make_resample_golden.py``).
"""
import math

import numpy as np


def sinc_kernel(orig_freq, new_freq, lowpass_filter_width=6, rolloff=0.99):
    """-> (kernel [new, 2*width + orig] float64, width, orig, new) with orig/new reduced by their gcd."""
    g = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // g, int(new_freq) // g
    base_freq = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base_freq)
    idx = np.arange(-width, width + orig, dtype=np.float64)[None, :] / orig
    t = np.arange(0, -new, -1, dtype=np.float64)[:, None] / new + idx
    t = np.clip(t * base_freq, -lowpass_filter_width, lowpass_filter_width)
    window = np.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t = t * math.pi
    with np.errstate(invalid="ignore", divide="ignore"):
        k = np.where(t == 0, 1.0, np.sin(t) / t)
    return k * window * (base_freq / orig), width, orig, new


def resample(x, orig_freq, new_freq):
    x = np.asarray(x, dtype=np.float64)
    if orig_freq == new_freq:
        return x
    k, width, orig, new = sinc_kernel(orig_freq, new_freq)
    n = len(x)
    padded = np.concatenate([np.zeros(width), x, np.zeros(width + orig)])
    frames = (len(padded) - k.shape[1]) // orig + 1
    win = np.lib.stride_tricks.sliding_window_view(padded, k.shape[1])[::orig][:frames]  # [frames, taps]
    out = (win @ k.T).reshape(-1)  # frame-major, phase-minor
    return out[:int(math.ceil(new * n / orig))]
