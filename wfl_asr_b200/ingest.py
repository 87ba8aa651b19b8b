"""Real-audio ingest: decoded waveform -> 16 kHz -> peak-normalised fp32 clips, on the device.

The reference does this per file on the host (REF/infer.py:217-244): ``soundfile.read`` (float64), mono mix-down,
``torchaudio.functional.resample`` on the float64 tensor, ``audio / (max|audio| + 1e-8)``, 30 s chunking with a second
per-chunk normalisation.  Here the file bytes are decoded on host worker threads (the only part that has to stay on the
CPU), and everything after the H2D copy runs in ``csrc/resample.cu`` (fp64 polyphase sinc bank, same filter as
torchaudio) and ``csrc/rowops.cu`` (fp64 peak normalisation).
"""
import math
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from . import ops

LOWPASS_FILTER_WIDTH = 6   # torchaudio.functional.resample defaults (the reference passes none)
ROLLOFF = 0.99
_BANKS = {}


def sinc_resample_bank(orig_freq, new_freq):
    """TORCHAUDIO/functional/functional.py ``_get_sinc_resample_kernel`` (sinc_interp_hann) in fp64, returned tap-major:
    (bank [2*width + orig, new] float64 CPU tensor, width, orig, new) with orig/new divided by their gcd."""
    if int(orig_freq) != orig_freq or int(new_freq) != new_freq or orig_freq <= 0 or new_freq <= 0:
        raise ValueError("sample rates must be positive integers")
    g = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // g, int(new_freq) // g
    base_freq = min(orig, new) * ROLLOFF
    width = math.ceil(LOWPASS_FILTER_WIDTH * orig / base_freq)
    idx = torch.arange(-width, width + orig, dtype=torch.float64)[None, :] / orig
    t = torch.arange(0, -new, -1, dtype=torch.float64)[:, None] / new + idx  # [new, taps]
    t = t * base_freq
    t = t.clamp_(-LOWPASS_FILTER_WIDTH, LOWPASS_FILTER_WIDTH)
    window = torch.cos(t * math.pi / LOWPASS_FILTER_WIDTH / 2) ** 2
    t = t * math.pi
    scale = base_freq / orig
    kernels = torch.where(t == 0, torch.tensor(1.0, dtype=torch.float64), t.sin() / t)
    kernels = kernels * (window * scale)
    return kernels.t().contiguous(), width, orig, new


def resample(audio_f64, orig_freq, new_freq):
    """fp64 [n] device tensor at ``orig_freq`` -> fp64 [ceil(new * n / orig)] device tensor at ``new_freq``."""
    if orig_freq == new_freq:
        return audio_f64
    if not audio_f64.is_cuda:
        raise RuntimeError("wfl_asr_b200 has no CPU path: resample() needs a CUDA tensor")
    key = (int(orig_freq), int(new_freq), audio_f64.device)
    ent = _BANKS.get(key)
    if ent is None:
        bank, width, orig, new = sinc_resample_bank(orig_freq, new_freq)
        if bank.numel() > (1 << 27):
            raise ValueError(f"resampling {orig_freq} -> {new_freq} Hz needs a {bank.shape[0]} x {bank.shape[1]} filter bank "
                             "(the rates share no useful common factor)")
        if len(_BANKS) >= 8:
            _BANKS.pop(next(iter(_BANKS)))
        ent = _BANKS[key] = (bank.to(audio_f64.device), width, orig, new)
    bank, width, orig, new = ent
    n = audio_f64.numel()
    n_out = -(-new * n // orig)
    x = audio_f64.contiguous()
    out = torch.empty(n_out, dtype=torch.float64, device=x.device)
    ops.resample_sinc(x, orig, new, width, bank, out)
    return out


def to_device_mono(audio, device):
    """numpy [n] or [n, channels] (any float/int dtype already scaled like soundfile) -> fp64 [n] device tensor."""
    a = np.asarray(audio, dtype=np.float64)
    if a.ndim == 2:
        a = a.mean(axis=1)  # REF/infer.py:218-219
    return torch.from_numpy(np.ascontiguousarray(a)).to(device, non_blocking=True)


def load_files(paths, target_sr, device, reader, workers=8):
    """Decode ``paths`` on a thread pool (file IO + PCM -> float64 releases the GIL in numpy) and yield, in order,
    (path, fp64 device waveform at ``target_sr``).  H2D copies of file i+1.. overlap the resampling of file i."""
    with ThreadPoolExecutor(max_workers=max(1, workers)) as pool:
        futures = [pool.submit(reader, p) for p in paths]
        for p, fut in zip(paths, futures):
            audio, sr = fut.result()
            yield p, resample(to_device_mono(audio, device), sr, target_sr)
