"""Python view of the handle-level C ABI (include/wfl_b200.h: wfl_create / wfl_set_weight / wfl_finalize / wfl_forward /
wfl_postprocess / wfl_destroy).  Packing, workspace and the launch schedule live in csrc/handle.cu; this file only
translates a config dict + state_dict into the calls a C, Go or Rust caller would make, and wraps device buffers.

    nm = NativeModel(config, labels, state_dict, device)          # create, set_weight x N, finalize, set_labels
    logits, offsets = nm.forward(wave, lang_id)                   # one wfl_forward (CUDA-graph replay from the 2nd call on)
    segs, nseg = nm.postprocess(logits, offsets, median_filter=5, merge_mode="right", confidence_threshold=0.5)

Scope of the handle: encoder_type "whisper" (csrc/handle.cu); other encoders raise WflError at construction."""
import ctypes

import numpy as np
import torch

from . import _lib, arch as _arch
from ._lib import MERGE_MODES, Config, WflError

ENCODER_CODES = {"whisper": 0, "wavlm": 1, "none": 2, "null": 2}
QUERY_LOGITS_STRIDE, QUERY_FRAMES, QUERY_WORKSPACE_BYTES, QUERY_GRAPHS = 0, 1, 2, 3
SEG_DTYPE = np.dtype([("start", "<f8"), ("end", "<f8"), ("ph", "<i4"), ("pad", "<i4")])


def native_config(config, n_labels, precision_high=True, max_batch=0):
    """config.yaml dict (REF/config.yaml layout) -> wfl_config."""
    m = config["model"]
    a = _arch.encoder_arch(config)
    c = Config()
    c.encoder_type = ENCODER_CODES[m["encoder_type"].lower()]
    c.d, c.layers = a["d"], a["layers"]
    c.heads, c.ffn, c.mels = a.get("heads", 0), a.get("ffn", 0), a.get("mels", 0)
    c.enable_bilstm = int(m.get("enable_bilstm", True))
    c.bilstm_layers = m.get("bilstm_num_layer", 1)
    c.n_conformer = m.get("num_conformer_layers", 2)
    c.conformer_heads = m.get("conformer_heads", 4)
    c.conformer_ff_expansion = m.get("conformer_ff_expansion", 4)
    c.conformer_kernel = m.get("conformer_kernel_size", 31)
    c.enable_dilated = int(m.get("enable_dilated_conv", True))
    c.dilated_depth = m.get("dilated_conv_depth", 2)
    c.dilated_kernel = m.get("dilated_conv_kernel", 3)
    c.n_labels = n_labels
    c.n_languages = m["num_languages"]
    c.lang_emb_dim = m.get("lang_emb_dim", 64)
    c.precision_high = int(precision_high)
    c.max_batch = int(max_batch)
    c.wavlm_layer_norm = int(a.get("norm") == "layer")
    return c


class NativeModel:
    def __init__(self, config, labels, state_dict, device, precision_high=True, max_batch=0):
        self.lib = _lib.load()
        self.dev = torch.device(device)
        self.labels = list(labels)
        self.L = len(self.labels)
        self.handle = ctypes.c_void_p()
        with torch.cuda.device(self.dev):
            cfg = native_config(config, self.L, precision_high, max_batch)
            _lib.check(self.lib.wfl_create(ctypes.byref(cfg), ctypes.byref(self.handle)), "wfl_create")
            try:
                for key, t in state_dict.items():
                    a = np.ascontiguousarray(t.detach().cpu().float().numpy())
                    shape = (ctypes.c_int64 * max(a.ndim, 1))(*a.shape)
                    _lib.check(self.lib.wfl_set_weight(self.handle, key.encode(), a.ctypes.data_as(ctypes.c_void_p), shape,
                                                       a.ndim), "wfl_set_weight")
                _lib.check(self.lib.wfl_finalize(self.handle), "wfl_finalize")
                arr = (ctypes.c_char_p * self.L)(*[s.encode() for s in self.labels])
                _lib.check(self.lib.wfl_set_labels(self.handle, arr, self.L), "wfl_set_labels")
            except Exception:
                self.close()
                raise
        self.Lp = self.query(QUERY_LOGITS_STRIDE)

    def query(self, what, arg=0):
        v = ctypes.c_int64()
        _lib.check(self.lib.wfl_query(self.handle, what, arg, ctypes.byref(v)), "wfl_query")
        return int(v.value)

    def packed(self, name, dtype=torch.uint8, nbytes=None):
        """Copy of one packed weight by its kernel-side name (diagnostic: wfl_packed_buffer); ``nbytes`` gives the
        extent for the "ws.<name>" workspace buffers, whose size the handle does not report."""
        ptr, n = ctypes.c_void_p(), ctypes.c_int64()
        _lib.check(self.lib.wfl_packed_buffer(self.handle, name.encode(), ctypes.byref(ptr), ctypes.byref(n)), "wfl_packed_buffer")
        if nbytes is not None:
            n.value = int(nbytes)
        if not n.value:
            return torch.empty(0, dtype=dtype, device=self.dev)

        class _View:  # the handle's memory, seen through the CUDA array interface for the duration of the copy
            __cuda_array_interface__ = {"shape": (n.value,), "typestr": "|u1", "data": (ptr.value, False), "version": 2}

        with torch.cuda.device(self.dev):
            return torch.as_tensor(_View(), device=self.dev).clone().view(dtype)

    @torch.no_grad()
    def forward(self, wave, lang_id=None, out=None):
        """wave fp32 [B, N] on the device -> (logits [B, T, L] view of a [B, T, Lp] buffer, offsets [B, T, 2])."""
        if not wave.is_cuda or wave.dtype != torch.float32 or wave.stride(1) != 1:
            raise WflError("NativeModel.forward needs a contiguous-row fp32 CUDA tensor (no CPU path)")
        B, N = wave.shape
        T = self.query(QUERY_FRAMES, N)
        if T < 1:  # the reference's conv stack raises on such a clip too (TF/models/wavlm/modeling_wavlm.py:703-751)
            raise WflError(f"clip of {N} samples is shorter than the encoder's receptive field")
        if out is None:
            out = (torch.empty(B, T, self.Lp, device=self.dev), torch.empty(B, T, 2, device=self.dev))
        logits, offsets = out
        lang = None if lang_id is None else lang_id.to(self.dev, torch.int64).contiguous()
        with torch.cuda.device(self.dev):
            stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            rc = self.lib.wfl_forward(self.handle, ctypes.c_void_p(wave.data_ptr()), wave.stride(0),
                                      None if lang is None else ctypes.c_void_p(lang.data_ptr()), B, N,
                                      ctypes.c_void_p(logits.data_ptr()), ctypes.c_void_p(offsets.data_ptr()), stream)
        _lib.check(rc, "wfl_forward")
        self._keep = lang  # the ids must outlive the asynchronous pass
        return logits[:, :, :self.L], offsets

    @torch.no_grad()
    def postprocess(self, logits, offsets, median_filter=1, merge_mode="right", confidence_threshold=0.0, frames=None):
        """-> (segment records uint8 [B, T, 24] on the device, counts int32 [B]).  ``logits``: as forward returned them."""
        B, T = logits.shape[0], logits.shape[1]
        if logits.stride(1) != self.Lp:
            raise WflError("postprocess expects the logits buffer wfl_forward wrote (row stride = logits stride)")
        segs = torch.empty(B, T, 24, dtype=torch.uint8, device=self.dev)
        nseg = torch.empty(B, dtype=torch.int32, device=self.dev)
        with torch.cuda.device(self.dev):
            stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            rc = self.lib.wfl_postprocess(self.handle, ctypes.c_void_p(logits.data_ptr()), ctypes.c_void_p(offsets.data_ptr()),
                                          None if frames is None else ctypes.c_void_p(frames.data_ptr()), B, T,
                                          float(confidence_threshold), int(median_filter), MERGE_MODES[merge_mode],
                                          ctypes.c_void_p(segs.data_ptr()), ctypes.c_void_p(nseg.data_ptr()), stream)
        _lib.check(rc, "wfl_postprocess")
        return segs, nseg

    def close(self):
        if self.handle:
            self.lib.wfl_destroy(self.handle)
            self.handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
