"""Architecture tables for the encoders the reference instantiates by name
(REF/model.py:57-91: ``whisper_model`` / ``wavlm_model`` / encoder_type "none").  Values are the public model-card
hyper-parameters; they cannot be fetched offline, so ``config["model"]`` may
override any of them through ``encoder_arch_override`` (a dict) and the test-only
``encoder_layers_override``."""
import math

import torch

WHISPER = {
    "tiny": dict(d=384, layers=4, heads=6, ffn=1536, mels=80),
    "base": dict(d=512, layers=6, heads=8, ffn=2048, mels=80),
    "small": dict(d=768, layers=12, heads=12, ffn=3072, mels=80),
    "medium": dict(d=1024, layers=24, heads=16, ffn=4096, mels=80),
    "large": dict(d=1280, layers=32, heads=20, ffn=5120, mels=80),
    "large-v2": dict(d=1280, layers=32, heads=20, ffn=5120, mels=80),
    "large-v3": dict(d=1280, layers=32, heads=20, ffn=5120, mels=128),
}
WAVLM = {
    "base": dict(d=768, layers=12, heads=12, ffn=3072, norm="group", stable_ln=False, do_normalize=False),
    "base-plus": dict(d=768, layers=12, heads=12, ffn=3072, norm="group", stable_ln=False, do_normalize=False),
    "large": dict(d=1024, layers=24, heads=16, ffn=4096, norm="layer", stable_ln=True, do_normalize=True),
}
WAVLM_CONV = dict(dim=512, kernels=(10, 3, 3, 3, 3, 2, 2), strides=(5, 2, 2, 2, 2, 2, 2), pos_k=128, pos_groups=16,
                  num_buckets=320, max_distance=800)


def encoder_arch(config):
    m = config["model"]
    et = m["encoder_type"].lower()
    if et == "whisper":
        name = m["whisper_model"].split("whisper-")[-1]
        table = WHISPER
    elif et == "wavlm":
        name = m["wavlm_model"].split("wavlm-")[-1]
        table = WAVLM
    elif et in ("none", "null"):
        # REF/model.py:82-91: no encoder; torchaudio MelSpectrogram(n_fft 400, hop = frame_duration * sample_rate) power
        # features ARE the hidden states, so the hidden size is n_mels
        data = config["data"]
        n_mels = int(data.get("n_mels", 80))
        return dict(type="none", d=n_mels, mels=n_mels, layers=0, sample_rate=int(data["sample_rate"]),
                    hop=int(data.get("frame_duration", 0.02) * data["sample_rate"]))
    else:
        raise ValueError("Unsupported encoder type. Use 'whisper', 'wavlm', or 'none'.")  # REF/model.py:94
    if name not in table:
        raise ValueError(f"unknown {et} architecture '{name}'; known: {sorted(table)}")
    a = dict(table[name])
    a.update(m.get("encoder_arch_override", {}) or {})
    if "encoder_layers_override" in m:
        a["layers"] = m["encoder_layers_override"]
    a["type"] = et
    return a


def whisper_sinusoids(length, channels, max_timescale=10000.0):
    """Initial value of Whisper's fixed positional table (TF/models/whisper/modeling_whisper.py:55-64)."""
    inc = math.log(max_timescale) / (channels // 2 - 1)
    inv = torch.exp(-inc * torch.arange(channels // 2))
    t = torch.arange(length).view(-1, 1) * inv.view(1, -1)
    return torch.cat([t.sin(), t.cos()], dim=1)


def wavlm_num_frames(n):
    for k, s in zip(WAVLM_CONV["kernels"], WAVLM_CONV["strides"]):
        n = (n - k) // s + 1
    return n
