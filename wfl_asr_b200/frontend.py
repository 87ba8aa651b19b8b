"""Host-side constants of the Whisper log-mel front-end (computed once in fp64, uploaded as fp32).

Arithmetic they feed: csrc/logmel.cu (replaces TF/models/whisper/feature_extraction_whisper.py:135-164,
filters per TF/audio_utils.py mel_filter_bank(norm="slaney", mel_scale="slaney"))."""
import functools

import numpy as np
import torch

N_FFT, HOP, N_BINS, BASIS_COLS = 400, 160, 201, 448


def dft_basis(window=None):
    """[400][448] fp32: column 2k = w[n] cos(2 pi k n / 400), 2k+1 = -w[n] sin(...), k <= 200; rest zero.
    ``window``: 400 samples (default: the periodic Hann window Whisper uses)."""
    n = np.arange(N_FFT, dtype=np.float64)
    if window is None:
        window = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / N_FFT)  # torch.hann_window(400) (periodic)
    window = np.asarray(window, dtype=np.float64)
    k = np.arange(N_BINS, dtype=np.float64)
    ang = 2.0 * np.pi * np.outer(n, k) / N_FFT
    basis = np.zeros((N_FFT, BASIS_COLS), dtype=np.float64)
    basis[:, 0:2 * N_BINS:2] = window[:, None] * np.cos(ang)
    basis[:, 1:2 * N_BINS:2] = -window[:, None] * np.sin(ang)
    return basis.astype(np.float32)


def mel_filters(n_mels, sr=16000, fmin=0.0, fmax=8000.0):
    """Slaney-scale, Slaney-normalised triangular filters, [201][n_mels] fp32."""
    def hz_to_mel(f):
        f = np.asarray(f, dtype=np.float64)
        return np.where(f >= 1000.0, 15.0 + np.log(np.maximum(f, 1e-10) / 1000.0) * (27.0 / np.log(6.4)), 3.0 * f / 200.0)

    def mel_to_hz(m):
        m = np.asarray(m, dtype=np.float64)
        return np.where(m >= 15.0, 1000.0 * np.exp((np.log(6.4) / 27.0) * (m - 15.0)), 200.0 * m / 3.0)

    freqs = np.linspace(0, sr // 2, N_BINS)
    edges = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))
    width = np.diff(edges)
    slopes = edges[None, :] - freqs[:, None]
    fb = np.maximum(0.0, np.minimum(-slopes[:, :-2] / width[:-1], slopes[:, 2:] / width[1:]))
    fb *= (2.0 / (edges[2:] - edges[:-2]))[None, :]
    return fb.astype(np.float32)


def dft_basis_split(window=None):
    """f16 [448][3*448] = [W_hi | W_mid | W_hi]: the DFT basis transposed (row = output column, K = sample index,
    zero padded 400 -> 448) and split hi + mid for the split-precision tensor-core contraction in csrc/logmel.cu."""
    wt = torch.zeros(BASIS_COLS, BASIS_COLS)
    wt[:, :N_FFT] = torch.from_numpy(dft_basis(window)).t()
    hi = wt.to(torch.float16)
    mid = (wt - hi.float()).to(torch.float16)
    return torch.cat([hi, mid, hi], dim=1).contiguous()


PLANE_SAMPLES = 3003 * HOP  # one padded-waveform plane of the strided view (csrc/logmel.cu kPlane)


@functools.lru_cache(maxsize=None)
def _constants_cpu(n_mels):
    return dft_basis_split(), torch.from_numpy(mel_filters(n_mels)).contiguous()


_DEVICE_CACHE = {}


def whisper_frontend_constants(n_mels, device):
    key = (n_mels, str(device))
    if key not in _DEVICE_CACHE:
        b, f = _constants_cpu(n_mels)
        _DEVICE_CACHE[key] = (b.to(device), f.to(device))
    return _DEVICE_CACHE[key]


def htk_mel_filters(n_mels, n_freqs=N_BINS, sr=16000, f_min=0.0, f_max=None):
    """Initial value of torchaudio MelSpectrogram's ``mel_scale.fb`` buffer (melscale_fbanks, norm=None, mel_scale="htk",
    fp32 arithmetic) as REF/model.py:85-90 constructs it; a checkpoint's own buffer replaces it on load."""
    import math
    f_max = float(sr // 2) if f_max is None else f_max
    freqs = torch.linspace(0, sr // 2, n_freqs)
    lo, hi = (2595.0 * math.log10(1.0 + f / 700.0) for f in (f_min, f_max))
    pts = 700.0 * (10 ** (torch.linspace(lo, hi, n_mels + 2) / 2595.0) - 1.0)
    width = pts[1:] - pts[:-1]
    slopes = pts.unsqueeze(0) - freqs.unsqueeze(1)
    falling = (-1.0 * slopes[:, :-2]) / width[:-1]
    rising = slopes[:, 2:] / width[1:]
    return torch.max(torch.zeros(1), torch.min(falling, rising))
