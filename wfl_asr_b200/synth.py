"""Synthetic workload of SURVEY.md section 8(d): deterministic 16 kHz utterances, label list and config used by
bench.py, smoke() and the examples.  (The test oracle has its own generator; this one ships with the product.)"""
import copy

import numpy as np
import torch

BASE_CONFIG = {
    "data": {"sample_rate": 16000, "frame_duration": 0.02, "n_mels": 80},
    "model": {"encoder_type": "whisper", "whisper_model": "openai/whisper-base", "wavlm_model": "microsoft/wavlm-base-plus",
              "freeze_encoder": False, "enable_bilstm": True, "bilstm_num_layer": 2, "enable_dilated_conv": True,
              "dilated_conv_depth": 2, "dilated_conv_kernel": 3, "num_conformer_layers": 2, "conformer_heads": 2,
              "conformer_ff_expansion": 2, "conformer_kernel_size": 31, "conformer_dropout": 0.15,
              "lang_emb_dim": 64, "num_languages": 2},
    "output": {"save_dir": "."},
    "postprocess": {"median_filter": 5, "merge_segments": "right", "confidence_threshold": 0.5},
}

# BASELINE.json configs[0..4]
WORKLOADS = {
    "cfg1": dict(model=dict(encoder_type="wavlm", wavlm_model="microsoft/wavlm-base-plus", enable_bilstm=False,
                            enable_dilated_conv=False, num_conformer_layers=2), batch=1, seconds=10.0),
    "cfg2": dict(model=dict(encoder_type="whisper", whisper_model="openai/whisper-base", enable_bilstm=False,
                            enable_dilated_conv=False, num_conformer_layers=4), batch=32, seconds=30.0),
    "cfg3": dict(model=dict(encoder_type="whisper", whisper_model="openai/whisper-small", enable_bilstm=True,
                            bilstm_num_layer=2, enable_dilated_conv=True, num_conformer_layers=4), batch=64, seconds=30.0),
    "cfg4": dict(model=dict(encoder_type="wavlm", wavlm_model="microsoft/wavlm-large", enable_bilstm=False,
                            enable_dilated_conv=False, num_conformer_layers=6), batch=16, seconds=16.0),
    "cfg5": dict(model=dict(encoder_type="whisper", whisper_model="openai/whisper-large-v3", enable_bilstm=True,
                            bilstm_num_layer=2, enable_dilated_conv=True, num_conformer_layers=8), batch=16, seconds=30.0),
}


def workload_config(name):
    cfg = copy.deepcopy(BASE_CONFIG)
    cfg["model"].update(WORKLOADS[name]["model"])
    return cfg


def synth_labels(n_phonemes=30):
    ph = [f"p{i}" for i in range(n_phonemes)]
    return sorted([f"B-{p}" for p in ph] + [f"I-{p}" for p in ph] + ["O"])  # sorted like REF/preprocess.py:163-166


def synth_wave(index, seconds, sr=16000):
    """0.6 x one-pole low-passed Gaussian noise + 0.4 x three sinusoids (100-4000 Hz) under a 2-8 Hz envelope,
    peak-normalised (x / (max|x| + 1e-8)); float64 [N]."""
    g = torch.Generator().manual_seed(1234 + index)
    n = int(round(seconds * sr))
    noise = torch.randn(n, generator=g, dtype=torch.float64).numpy()
    low = np.empty(n)
    alpha, acc = 0.85, 0.0
    # one-pole low-pass y[i] = alpha*y[i-1] + (1-alpha)*x[i], vectorised in blocks via cumulative products
    from scipy.signal import lfilter
    low = lfilter([1 - alpha], [1, -alpha], noise)
    low = low / (np.abs(low).max() + 1e-12)
    t = np.arange(n) / sr
    fr = (100 + 3900 * torch.rand(3, generator=g, dtype=torch.float64)).numpy()
    env_f = float(2 + 6 * torch.rand(1, generator=g, dtype=torch.float64))
    tones = sum(np.sin(2 * np.pi * f * t) for f in fr) / 3.0
    x = 0.6 * low + 0.4 * tones * (0.5 + 0.5 * np.sin(2 * np.pi * env_f * t))
    return x / (np.max(np.abs(x)) + 1e-8)


def randomize_batchnorm(model, seed=1):
    """Random eval-mode BatchNorm statistics so the BN fold is exercised (default init is the identity)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for mod in model.modules():
            if isinstance(mod, torch.nn.BatchNorm1d):
                mod.running_mean.copy_(torch.randn(mod.running_mean.shape, generator=g) * 0.1)
                mod.running_var.copy_(torch.rand(mod.running_var.shape, generator=g) * 0.5 + 0.75)
                mod.weight.copy_(1.0 + 0.1 * torch.randn(mod.weight.shape, generator=g))
                mod.bias.copy_(0.1 * torch.randn(mod.bias.shape, generator=g))
    return model


def bench_model(cls, cfg, labels, seed=0, classifier_gain=60.0):
    """Random-init model of the named architecture (torch.manual_seed(seed)) with randomised BatchNorm statistics.
    The classifier is scaled by ``classifier_gain`` so that random-init logits are peaked enough for frames to clear
    the 0.5 confidence threshold -- otherwise every frame decodes to "O" and the median/BIO/merge kernels would be
    benchmarked on empty output.  Same weights for the CUDA arm and the CPU reference arm."""
    torch.manual_seed(seed)
    model = randomize_batchnorm(cls(cfg, labels), seed + 1)
    with torch.no_grad():
        model.classifier.weight.mul_(classifier_gain)
        model.classifier.bias.mul_(classifier_gain)
    return model


class RaggedCorpus:
    """BASELINE configs[3] corpus: ``n`` utterances with lengths U[2, 30] s drawn with
    ``numpy.random.default_rng(4242)``.  Utterance i is a window of one of ``pool`` 30 s base clips (synth_wave) scaled
    by a per-utterance gain; it is materialised when indexed, so a rank only ever builds its own shard
    (bulk.label_corpus indexes nothing else when ``lengths`` is passed)."""

    def __init__(self, n, seed=4242, pool=8, sr=16000, min_s=2.0, max_s=30.0):
        rng = np.random.default_rng(seed)
        self.seconds = rng.uniform(min_s, max_s, size=n)
        self.lengths = [int(s * sr) for s in self.seconds]
        self.starts = [int(rng.integers(0, int(max_s * sr) - ln + 1)) for ln in self.lengths]
        self.gains = 0.5 + 0.5 * rng.random(n)
        self.sr = sr
        self._pool_size = pool
        self._pool = None

    def _base(self, k):
        if self._pool is None:
            self._pool = [synth_wave(7000 + j, 30.0, self.sr).astype(np.float32) for j in range(self._pool_size)]
        return self._pool[k]

    def __len__(self):
        return len(self.lengths)

    def __getitem__(self, i):
        s, n = self.starts[i], self.lengths[i]
        return self._base(i % self._pool_size)[s:s + n] * np.float32(self.gains[i])

    @property
    def audio_seconds(self):
        return float(sum(self.lengths)) / self.sr
