"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Loads the *unmodified* reference (``/root/reference/model.py``) with the Hugging Face
``from_pretrained`` constructors replaced by offline random-init constructors of the named
architecture (SURVEY.md section 7 step 1 / section 9 "offline oracle recipe").  Only usable in the
authoring container, where ``/root/reference`` exists; the GPU box uses the fixtures this produces
(``tests/golden/``) and the self-contained torch restatement in ``oracle/torch_oracle.py``.

Reference call sites being replaced: REF/model.py:69-70 (Whisper), REF/model.py:74-80 (WavLM).
"""
import importlib
import os
import sys
import types

import torch

REF_DIR = os.environ.get("WFL_REFERENCE_DIR", "/root/reference")

# Public model-card hyper-parameters (cannot be fetched offline).
WHISPER_ARCH = {
    "tiny": dict(d_model=384, encoder_layers=4, encoder_attention_heads=6, encoder_ffn_dim=1536, num_mel_bins=80),
    "base": dict(d_model=512, encoder_layers=6, encoder_attention_heads=8, encoder_ffn_dim=2048, num_mel_bins=80),
    "small": dict(d_model=768, encoder_layers=12, encoder_attention_heads=12, encoder_ffn_dim=3072, num_mel_bins=80),
    "medium": dict(d_model=1024, encoder_layers=24, encoder_attention_heads=16, encoder_ffn_dim=4096, num_mel_bins=80),
    "large-v3": dict(d_model=1280, encoder_layers=32, encoder_attention_heads=20, encoder_ffn_dim=5120, num_mel_bins=128),
}
WAVLM_ARCH = {
    "base-plus": dict(hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072,
                      feat_extract_norm="group", do_stable_layer_norm=False, _do_normalize=False),
    "large": dict(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096,
                  feat_extract_norm="layer", do_stable_layer_norm=True, _do_normalize=True),
}


def whisper_arch(name):
    key = name.split("whisper-")[-1]
    return dict(WHISPER_ARCH[key])


def wavlm_arch(name):
    key = name.split("wavlm-")[-1]
    return dict(WAVLM_ARCH[key])


def available():
    return os.path.isfile(os.path.join(REF_DIR, "model.py"))


def _patch_transformers(layer_override=None):
    import transformers
    from transformers import (Wav2Vec2FeatureExtractor, WavLMConfig, WavLMModel, WhisperConfig,
                              WhisperFeatureExtractor, WhisperModel)

    def whisper_fe(cls, name, *a, **k):
        return WhisperFeatureExtractor(feature_size=whisper_arch(name)["num_mel_bins"])

    def whisper_model(cls, name, *a, **k):
        arch = whisper_arch(name)
        if layer_override is not None:
            arch["encoder_layers"] = layer_override
        cfg = WhisperConfig(decoder_layers=1, decoder_attention_heads=arch["encoder_attention_heads"],
                            decoder_ffn_dim=64, vocab_size=64, max_source_positions=1500,
                            pad_token_id=0, bos_token_id=1, eos_token_id=2, decoder_start_token_id=1,
                            **arch)
        return WhisperModel(cfg)

    def wavlm_cfg(cls, name, *a, **k):
        arch = wavlm_arch(name)
        arch.pop("_do_normalize")
        if layer_override is not None:
            arch["num_hidden_layers"] = layer_override
        return WavLMConfig(**arch)

    def wavlm_model(cls, name, config=None, *a, **k):
        return WavLMModel(config)

    def w2v_fe(cls, name, *a, **k):
        return Wav2Vec2FeatureExtractor(do_normalize=wavlm_arch(name)["_do_normalize"])

    WhisperFeatureExtractor.from_pretrained = classmethod(whisper_fe)
    WhisperModel.from_pretrained = classmethod(whisper_model)
    WavLMConfig.from_pretrained = classmethod(wavlm_cfg)
    WavLMModel.from_pretrained = classmethod(wavlm_model)
    Wav2Vec2FeatureExtractor.from_pretrained = classmethod(w2v_fe)
    return transformers


def load_reference_modules(layer_override=None):
    """Returns (ref_model_module, ref_utils_module, ref_infer_module) imported from REF_DIR.

    ``soundfile`` and ``matplotlib`` are absent here; REF/utils.py:4-5 and REF/infer.py:5 import
    them at module scope but the hot-path functions never touch them, so inert stubs are installed.
    """
    if not available():
        raise RuntimeError(f"reference not found at {REF_DIR}")
    os.environ.setdefault("HF_HUB_OFFLINE", "1")
    _patch_transformers(layer_override)
    for name in ("soundfile", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if "matplotlib" in sys.modules and not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    mods = []
    for name in ("model", "utils", "infer"):
        spec = importlib.util.spec_from_file_location(f"wfl_reference_{name}", os.path.join(REF_DIR, f"{name}.py"))
        mod = importlib.util.module_from_spec(spec)
        # REF/infer.py does ``from model import ...`` / ``from utils import ...``
        saved = {k: sys.modules.get(k) for k in ("model", "utils")}
        if name == "infer":
            sys.modules["model"], sys.modules["utils"] = mods[0], mods[1]
        try:
            spec.loader.exec_module(mod)
        finally:
            if name == "infer":
                for k, v in saved.items():
                    if v is None:
                        sys.modules.pop(k, None)
                    else:
                        sys.modules[k] = v
        mods.append(mod)
    return tuple(mods)


def build_reference_model(config, labels, seed=0, layer_override=None, randomize_bn=True):
    """Random-init reference ``BIOPhonemeTagger`` (REF/model.py:55) in eval mode, fp32, CPU."""
    ref_model, _, _ = load_reference_modules(layer_override)
    torch.manual_seed(seed)
    m = ref_model.BIOPhonemeTagger(config, labels)
    if randomize_bn:
        g = torch.Generator().manual_seed(seed + 1)
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm1d):
                mod.running_mean.copy_(torch.randn(mod.running_mean.shape, generator=g) * 0.1)
                mod.running_var.copy_(torch.rand(mod.running_var.shape, generator=g) * 0.5 + 0.75)
                with torch.no_grad():
                    mod.weight.copy_(1.0 + 0.1 * torch.randn(mod.weight.shape, generator=g))
                    mod.bias.copy_(0.1 * torch.randn(mod.bias.shape, generator=g))
    return m.eval()
