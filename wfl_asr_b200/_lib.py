"""ctypes binding of libwfl_b200.so (include/wfl_b200.h).  No fallback: if the library cannot be
loaded or a call fails, a ``WflError`` is raised."""
import ctypes
import os

from . import build as _build

_C = ctypes
WFL_MAX_SLABS = 128

ACT_NONE, ACT_GELU, ACT_RELU = 0, 1, 2
OUT_STORE_F16, OUT_STORE_F32, OUT_ADD_F32, OUT_GLU_F16 = 0, 1, 2, 3
TAG_O, TAG_B, TAG_I, TAG_OTHER = 0, 1, 2, 3
MERGE_MODES = {"none": 0, "right": 1, "left": 2, "previous": 3}


class WflError(RuntimeError):
    pass


class GemmDesc(_C.Structure):
    _fields_ = [
        ("a", _C.c_void_p),
        ("a_rows", _C.c_int64), ("a_cols", _C.c_int64), ("a_row_stride", _C.c_int64), ("a_batch_stride", _C.c_int64),
        ("batches", _C.c_int32),
        ("w", _C.c_void_p),
        ("n", _C.c_int32), ("slab_k", _C.c_int32), ("num_slabs", _C.c_int32),
        ("slab_row_shift", _C.c_int32 * WFL_MAX_SLABS),
        ("slab_a_col", _C.c_int32 * WFL_MAX_SLABS),
        ("bias", _C.c_void_p),
        ("bias_batch_stride", _C.c_int64),
        ("act", _C.c_int32), ("out_mode", _C.c_int32),
        ("alpha", _C.c_float),
        ("out", _C.c_void_p),
        ("m_rows", _C.c_int64), ("out_row_stride", _C.c_int64), ("out_batch_stride", _C.c_int64),
        ("tile_n", _C.c_int32),
        ("groups", _C.c_int32),
        ("a_col_group_stride", _C.c_int64), ("out_col_group_stride", _C.c_int64),
    ]


class Config(_C.Structure):
    """include/wfl_b200.h wfl_config."""
    _fields_ = [(n, _C.c_int32) for n in (
        "encoder_type", "d", "layers", "heads", "ffn", "mels", "enable_bilstm", "bilstm_layers", "n_conformer",
        "conformer_heads", "conformer_ff_expansion", "conformer_kernel", "enable_dilated", "dilated_depth",
        "dilated_kernel", "n_labels", "n_languages", "lang_emb_dim", "precision_high", "max_batch", "wavlm_layer_norm")]


class Segment(_C.Structure):
    _fields_ = [("start", _C.c_double), ("end", _C.c_double), ("ph", _C.c_int32), ("pad_", _C.c_int32)]


# name -> (argtypes); every function returns int.  Kept in one table so the CPU test can check that
# the library exports exactly what include/wfl_b200.h declares.
_P, _I32, _I64, _F, _D = _C.c_void_p, _C.c_int32, _C.c_int64, _C.c_float, _C.c_double
SIGNATURES = {
    "wfl_device_info": [_P, _P, _P],
    "wfl_gemm": [_C.POINTER(GemmDesc), _P],
    "wfl_attention": [_P, _I64, _I64, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _F, _P, _P, _P, _I64, _I64, _P],
    "wfl_layernorm": [_P, _I64, _I32, _P, _P, _P, _P, _F, _P, _P, _I32, _P],
    "wfl_split_f16": [_P, _I64, _I32, _P, _P],
    "wfl_broadcast_rows": [_P, _I64, _I32, _I32, _P, _P],
    "wfl_gather_cols": [_P, _I64, _I32, _I32, _I32, _P, _P],
    "wfl_mel_power": [_P, _I64, _I32, _I32, _I32, _P, _P, _I32, _P, _I32, _P, _P, _P, _P],
    "wfl_rowdot_sigmoid": [_P, _I64, _I32, _P, _P, _I32, _P, _P],
    "wfl_peak_normalize": [_P, _P, _I32, _P, _I64, _P, _P, _P],
    "wfl_resample_sinc": [_P, _I64, _I32, _I32, _I32, _P, _P, _I64, _P],
    "wfl_pcm_to_f64": [_P, _I32, _I32, _I64, _P, _P],
    "wfl_whisper_logmel": [_P, _I64, _I32, _I32, _P, _P, _I32, _P, _I32, _P, _P, _P, _P, _P],
    "wfl_decode_frames": [_P, _I64, _I32, _I64, _I32, _F, _P, _P],
    "wfl_median_filter": [_P, _P, _P, _I32, _I64, _I32, _P],
    "wfl_bio_decode": [_P, _P, _P, _I32, _I64, _P, _P, _I32, _D, _P, _P, _P, _P],
    "wfl_merge_segments": [_P, _P, _I64, _P, _I32, _P, _I32, _P, _P, _P],
    "wfl_htk_times": [_P, _I64, _P, _P, _P],
    "wfl_lstm_layer": [_P, _P, _I32, _I32, _I32, _P, _P, _P],
    "wfl_wavlm_conv0": [_P, _I64, _I32, _I32, _P, _P, _P, _I32, _P, _I64, _P, _P],
    "wfl_wavlm_gate": [_P, _I64, _I32, _I32, _I32, _I32, _P, _P, _P, _P, _P],
    "wfl_stft_mag": [_P, _I64, _I32, _I32, _I32, _P, _P],
    "wfl_spectral_flux": [_P, _I32, _I32, _P, _P],
    "wfl_mfcc_delta_mag": [_P, _I32, _I32, _P, _P, _I32, _P, _I32, _P, _P, _P, _P, _P],
    # handle level
    "wfl_create": [_C.POINTER(Config), _C.POINTER(_P)],
    "wfl_set_weight": [_P, _C.c_char_p, _P, _C.POINTER(_I64), _I32],
    "wfl_finalize": [_P],
    "wfl_set_labels": [_P, _C.POINTER(_C.c_char_p), _I32],
    "wfl_query": [_P, _I32, _I64, _C.POINTER(_I64)],
    "wfl_packed_buffer": [_P, _C.c_char_p, _C.POINTER(_P), _C.POINTER(_I64)],
    "wfl_forward": [_P, _P, _I64, _P, _I32, _I32, _P, _P, _P],
    "wfl_postprocess": [_P, _P, _P, _P, _I32, _I32, _F, _I32, _I32, _P, _P, _P],
}
NOARG = {"wfl_abi_version": _C.c_int, "wfl_last_error": _C.c_char_p}
VOID = {"wfl_destroy": [_P]}  # functions returning void

_lib = None


def lib_path():
    return _build.LIB_PATH


def load():
    """Loads (building first if the in-tree .so is missing) and returns the ctypes library."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("WFL_LIB") or _build.LIB_PATH  # WFL_LIB: an experiment build of the same sources (tools/)
    if not os.path.exists(path):
        _build.build()
    try:
        lib = _C.CDLL(path)
    except OSError as e:  # fail loudly: there is no other implementation to fall back to
        raise WflError(f"cannot load {path}: {e}") from e
    for name, restype in NOARG.items():
        fn = getattr(lib, name)
        fn.argtypes = []
        fn.restype = restype
    for name, argtypes in VOID.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = None
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name, None)
        if fn is None:
            continue  # optional entry points are checked by callers via hasattr
        fn.argtypes = argtypes
        fn.restype = _C.c_int
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().wfl_last_error()
        raise WflError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")
