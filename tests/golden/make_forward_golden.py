"""Generates tests/golden/forward_golden.npz by running the UNMODIFIED reference model
(REF/model.py BIOPhonemeTagger, loaded by oracle/ref_loader.py) on deterministic weights from
``oracle.torch_oracle.random_state_dict`` (strict ``load_state_dict``) and deterministic synthetic
audio.  Only reference OUTPUTS are committed (frame-strided logits/offsets + all argmax ids); the
weights are regenerated from the seed on whatever box runs the tests.

Run:  python tests/golden/make_forward_golden.py [case ...]    (needs /root/reference; with case names only those
      entries of the existing file are replaced)
"""
import copy
import os
import sys

import numpy as np
import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader, torch_oracle as to  # noqa: E402

STRIDE = 25

BASE = {
    "data": {"sample_rate": 16000, "frame_duration": 0.02, "n_mels": 80},
    "model": {"encoder_type": "whisper", "whisper_model": "openai/whisper-base", "wavlm_model": "microsoft/wavlm-base-plus",
              "freeze_encoder": False, "enable_bilstm": True, "bilstm_num_layer": 2, "enable_dilated_conv": True,
              "dilated_conv_depth": 2, "dilated_conv_kernel": 3, "num_conformer_layers": 2, "conformer_heads": 2,
              "conformer_ff_expansion": 2, "conformer_kernel_size": 31, "conformer_dropout": 0.15,
              "lang_emb_dim": 64, "num_languages": 2},
    "postprocess": {"median_filter": 5, "merge_segments": "right", "confidence_threshold": 0.5},
}

CASES = {
    # name: (model overrides, encoder layers, batch, seconds, seed)
    "whisper_base_full": (dict(), 2, 2, 3.0, 11),
    "whisper_base_cfg2": (dict(enable_bilstm=False, enable_dilated_conv=False, num_conformer_layers=4), 2, 1, 2.0, 12),
    "wavlm_base_plus": (dict(encoder_type="wavlm", enable_bilstm=False, enable_dilated_conv=False), 2, 2, 2.0, 13),
    # conformer_heads=4 -> head_dim 256 at d=1024, with the dilated stack
    "wavlm_large": (dict(encoder_type="wavlm", wavlm_model="microsoft/wavlm-large", num_conformer_layers=1,
                         conformer_heads=4, enable_bilstm=False), 2, 2, 1.5, 14),
    # encoder_type "none" (REF/model.py:82-91): MelSpectrogram power features, hidden size 80, 101 frames per 2 s
    "mel_none_full": (dict(encoder_type="none"), 0, 2, 2.0, 15),
    # --- the BASELINE.json config SHAPES (encoder depth cut to 2 layers so the fixtures and the CPU oracle stay small) ---
    # configs[0]: WavLM-base-plus + 2 Conformer (config.yaml heads 2 -> head dim 384), batch 1 x 10 s -> T = 499
    "cfg1_wavlm_base_plus_10s": (dict(encoder_type="wavlm", enable_bilstm=False, enable_dilated_conv=False,
                                      num_conformer_layers=2), 2, 1, 10.0, 16),
    # configs[3]: WavLM-large + Conformer with heads 2 -> head dim 512 (attention_big_kernel<512>), d = 1024 packing
    "cfg4_wavlm_large_hd512": (dict(encoder_type="wavlm", wavlm_model="microsoft/wavlm-large", enable_bilstm=False,
                                    enable_dilated_conv=False, num_conformer_layers=2), 2, 2, 2.5, 17),
    # configs[4]: Whisper-large-v3 (d 1280, 20 heads, 128 mels) + BiLSTM(2) H 640 + Conformer heads 2 -> head dim 640
    # (attention_big_kernel<640>, lstm_kernel<640,16,8>) + dilated stack
    "cfg5_whisper_large_v3_full": (dict(whisper_model="openai/whisper-large-v3", num_conformer_layers=2), 2, 1, 3.0, 18),
}


def case_config(name):
    upd, layers, B, secs, seed = CASES[name]
    cfg = copy.deepcopy(BASE)
    cfg["model"].update(upd)
    cfg["model"]["encoder_layers_override"] = layers
    return cfg, B, secs, seed


def case_inputs(name):
    cfg, B, secs, seed = case_config(name)
    labels = to.synth_labels(30)
    sd = to.random_state_dict(cfg, len(labels), seed=seed)
    wave = torch.stack([torch.from_numpy(to.synth_wave(100 + i, secs)).float() for i in range(B)])
    lang = torch.tensor([i % 2 for i in range(B)])
    return cfg, labels, sd, wave, lang


def main():
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "forward_golden.npz")
    only = sys.argv[1:]
    out = dict(np.load(dst)) if only else {}
    for name in (only or CASES):
        cfg, labels, sd, wave, lang = case_inputs(name)
        ref = ref_loader.build_reference_model(cfg, labels, layer_override=cfg["model"]["encoder_layers_override"],
                                               randomize_bn=False)
        ref.load_state_dict(sd, strict=True)
        with torch.no_grad():
            logits, offsets = ref(wave, lang)
        out[name + "/logits"] = logits[:, ::STRIDE].numpy()
        out[name + "/offsets"] = offsets[:, ::STRIDE].numpy()
        out[name + "/argmax"] = logits.argmax(-1).numpy().astype(np.int16)
        print(name, tuple(logits.shape), float(logits.abs().max()))
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst) // 1024, "KiB")


if __name__ == "__main__":
    main()
