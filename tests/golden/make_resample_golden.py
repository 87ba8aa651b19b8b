"""Generates tests/golden/resample_golden.npz from torchaudio itself (authoring container only):
torchaudio.functional.resample on float64 inputs, exactly as REF/infer.py:217-220 calls it."""
import os

import numpy as np
import torch

CASES = [(44100, 16000, 11025), (48000, 16000, 12001), (22050, 16000, 7777), (8000, 16000, 4000), (32000, 16000, 9)]


def signal(sr, n, seed):
    g = np.random.default_rng(seed)
    t = np.arange(n) / sr
    return 0.5 * np.sin(2 * np.pi * 440.0 * t) + 0.3 * np.sin(2 * np.pi * 3100.0 * t) + 0.2 * g.standard_normal(n)


def main():
    import torchaudio
    out = {}
    for sr, target, n in CASES:
        x = signal(sr, n, sr + n)
        y = torchaudio.functional.resample(torch.tensor(x), orig_freq=sr, new_freq=target).numpy()
        out[f"{sr}_{target}_{n}"] = y
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "resample_golden.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
