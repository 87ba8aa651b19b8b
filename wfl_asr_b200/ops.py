"""Thin torch-tensor wrappers over the C ABI (include/wfl_b200.h).  Torch is used for device memory and
the current stream only; every op below is one or more launches of the hand-written sm_100a kernels.
All wrappers require CUDA tensors and raise ``WflError`` otherwise -- there is no CPU path."""
import ctypes

import torch

from . import _lib
from ._lib import (ACT_GELU, ACT_NONE, ACT_RELU, MERGE_MODES, OUT_ADD_F32, OUT_GLU_F16, OUT_STORE_F16,  # noqa: F401
                   OUT_STORE_F32, GemmDesc, Segment, WflError)

LAUNCHES = 0  # kernels launched through this module (bench.py reports it as gpu_launches)
TIMING = None  # bench.py sets this to a list: timed launches append (kind, tag, algorithmic work, start, end events);
#                work is FLOPs for kind "gemm" / "attention", bytes for the memory-bound kinds, steps for "lstm"
TIMING_MIN_SLABS = 1  # GEMMs: only launches with at least this many K-slabs are bracketed by events (event records
#                       between kernels defeat programmatic dependent launch, so the timed region brackets the dominant
#                       kernel only); the other kinds are bracketed only when TIMING_KINDS names them
TIMING_KINDS = ()


def _t0(kind):
    if TIMING is None or kind not in TIMING_KINDS:
        return None
    e0 = torch.cuda.Event(enable_timing=True)
    e0.record()
    return e0


def _t1(e0, kind, tag, work):
    if e0 is not None:
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        TIMING.append((kind, tag, float(work), e0, e1))


def _count(n=1):
    global LAUNCHES
    LAUNCHES += n


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    if t is None:
        return None
    if not t.is_cuda:
        raise WflError("wfl ops need CUDA tensors (no CPU fallback exists)")
    return ctypes.c_void_p(t.data_ptr())


def gemm(a, w, out, *, n, slab_k, shifts=(0,), cols=(0,), a_rows, a_cols, a_row_stride, a_batch_stride=0, batches=1,
         m_rows=None, out_row_stride=None, out_batch_stride=0, bias=None, bias_batch_stride=0, act=ACT_NONE,
         out_mode=OUT_STORE_F16, alpha=1.0, tile_n=0, groups=1, a_col_group_stride=0, out_col_group_stride=0):
    """acc[b,t,n] = sum_s sum_k A[b, t+shift_s, col_s+k] * W[n, s*slab_k+k]; see wfl_gemm in the header.
    groups > 1: Conv1d(groups=G) in one launch (w [G*n, K], bias [G*n], column strides per group)."""
    d = GemmDesc()
    d.a = a.data_ptr()
    d.a_rows, d.a_cols, d.a_row_stride, d.a_batch_stride, d.batches = a_rows, a_cols, a_row_stride, a_batch_stride, batches
    d.w = w.data_ptr()
    d.n, d.slab_k, d.num_slabs = n, slab_k, len(shifts)
    for i, (s, c) in enumerate(zip(shifts, cols)):
        d.slab_row_shift[i] = s
        d.slab_a_col[i] = c
    d.bias = bias.data_ptr() if bias is not None else None
    d.bias_batch_stride = bias_batch_stride
    d.act, d.out_mode, d.alpha = act, out_mode, alpha
    d.out = out.data_ptr()
    d.m_rows = a_rows if m_rows is None else m_rows
    out_cols = n // 2 if out_mode == OUT_GLU_F16 else n
    d.out_row_stride = out_cols if out_row_stride is None else out_row_stride
    d.out_batch_stride = out_batch_stride
    d.tile_n = tile_n
    d.groups, d.a_col_group_stride, d.out_col_group_stride = groups, a_col_group_stride, out_col_group_stride
    if not (a.is_cuda and w.is_cuda and out.is_cuda):
        raise WflError("wfl_gemm needs CUDA tensors (no CPU fallback exists)")
    timed = TIMING is not None and len(shifts) >= TIMING_MIN_SLABS
    if timed:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    _lib.check(_lib.load().wfl_gemm(ctypes.byref(d), _stream()), "wfl_gemm")
    if timed:
        e1.record()
        tag = f"M{batches * d.m_rows}xN{n}xK{len(shifts) * slab_k}/slabs{len(shifts)}/mode{out_mode}" + (f"/g{groups}" if groups > 1 else "")
        TIMING.append(("gemm", tag, 2.0 * groups * batches * d.m_rows * n * len(shifts) * slab_k, e0, e1))
    _count()


def linear(a2d, w, out2d, **kw):
    """a2d [M, K] f16 (contiguous rows), w [N, K] f16 -> out2d [M, N or N/2]."""
    M, K = a2d.shape
    gemm(a2d, w, out2d, n=w.shape[0], slab_k=K, a_rows=M, a_cols=K, a_row_stride=a2d.stride(0), m_rows=M,
         out_row_stride=out2d.stride(0), **kw)


def attention(qkv, out, *, B, T, H, hd, scale, q_col, k_col, v_col, rel_bias=None, gate=None):
    """qkv f16 [B, T, W]; out f16 [B, T, H*hd]."""
    e0 = _t0("attention")
    rc = _lib.load().wfl_attention(_ptr(qkv), qkv.stride(1), qkv.stride(0), q_col, k_col, v_col, B, T, H, hd, scale,
                                   _ptr(rel_bias), _ptr(gate), _ptr(out), out.stride(1), out.stride(0), _stream())
    _lib.check(rc, "wfl_attention")
    _t1(e0, "attention", f"attention hd{hd} H{H} T{T} B{B}" + ("+relbias" if rel_bias is not None else ""), 4.0 * B * H * T * T * hd)
    _count()


def layernorm(x, gamma, beta, *, out_f32=None, out_f16=None, gamma2=None, beta2=None, eps=1e-5, act_f16=ACT_NONE,
              rows=None):
    rows = x.numel() // x.shape[-1] if rows is None else rows
    e0 = _t0("layernorm")
    rc = _lib.load().wfl_layernorm(_ptr(x), rows, x.shape[-1], _ptr(gamma), _ptr(beta), _ptr(gamma2), _ptr(beta2), eps,
                                   _ptr(out_f32), _ptr(out_f16), act_f16, _stream())
    _lib.check(rc, "wfl_layernorm")
    d = x.shape[-1]  # algorithmic bytes: fp32 row in, f16 and/or fp32 row out
    _t1(e0, "layernorm", f"layernorm d{d}" + ("+f32out" if out_f32 is not None else "") + ("+2nd" if gamma2 is not None else ""),
        rows * d * (4 + (2 if out_f16 is not None else 0) + (4 if out_f32 is not None else 0)))
    _count()


WAVLM_STATS_DOUBLES = 131072  # include/wfl_b200.h WFL_WAVLM_STATS_DOUBLES (doubles of scratch per clip)


def wavlm_conv0(wave, n_samples, w, gamma, beta, norm_mode, out, out_batch_stride, scratch):
    """wave fp32 [B, >=n_samples] -> out f16 [B, out_batch_stride/512 rows, 512] (conv k10 s5 + norm + GELU)."""
    rc = _lib.load().wfl_wavlm_conv0(_ptr(wave), wave.stride(0), n_samples, wave.shape[0], _ptr(w), _ptr(gamma),
                                     _ptr(beta), norm_mode, _ptr(out), out_batch_stride, _ptr(scratch), _stream())
    _lib.check(rc, "wfl_wavlm_conv0")
    _count(3)  # statistics partials, ordered finish, conv + norm + GELU


def wavlm_gate(x_f16, row_stride, B, T, H, hd, gw, gb, gconst, gate):
    rc = _lib.load().wfl_wavlm_gate(_ptr(x_f16), row_stride, B, T, H, hd, _ptr(gw), _ptr(gb), _ptr(gconst), _ptr(gate),
                                    _stream())
    _lib.check(rc, "wfl_wavlm_gate")
    _count()


def split_f16(x, out):
    rows = x.numel() // x.shape[-1]
    e0 = _t0("split_f16")
    _lib.check(_lib.load().wfl_split_f16(_ptr(x), rows, x.shape[-1], _ptr(out), _stream()), "wfl_split_f16")
    _t1(e0, "split_f16", f"split_f16 d{x.shape[-1]}", rows * x.shape[-1] * 8)  # fp32 in, [hi | lo] f16 out
    _count()


def broadcast_rows(src, dst, batches):
    rows, d = src.shape
    _lib.check(_lib.load().wfl_broadcast_rows(_ptr(src), rows, d, batches, _ptr(dst), _stream()), "wfl_broadcast_rows")
    _count()


def rowdot_sigmoid(x_f16, w, b, out):
    rows = x_f16.numel() // x_f16.shape[-1]
    rc = _lib.load().wfl_rowdot_sigmoid(_ptr(x_f16), rows, x_f16.shape[-1], _ptr(w), _ptr(b), w.shape[0], _ptr(out),
                                        _stream())
    _lib.check(rc, "wfl_rowdot_sigmoid")
    _count()


def peak_normalize(samples_f64, clip_begin, n_clips, out, scratch_max, out_f64=None):
    rc = _lib.load().wfl_peak_normalize(_ptr(samples_f64), _ptr(clip_begin), n_clips, _ptr(out),
                                        out.stride(0) if out is not None else 0, _ptr(out_f64), _ptr(scratch_max),
                                        _stream())
    _lib.check(rc, "wfl_peak_normalize")
    _count(2)


def resample_sinc(x_f64, orig, new, width, bank, out_f64):
    """x fp64 [n_in] (device) -> out fp64 [n_out]; orig/new already divided by their gcd; bank fp64 [2*width+orig, new]."""
    rc = _lib.load().wfl_resample_sinc(_ptr(x_f64), x_f64.numel(), orig, new, width, _ptr(bank), _ptr(out_f64),
                                       out_f64.numel(), _stream())
    _lib.check(rc, "wfl_resample_sinc")
    _count(1)


PCM_S16, PCM_S32, PCM_F32 = 16, 32, 3  # include/wfl_b200.h WFL_PCM_*


def pcm_to_f64(pcm_bytes, fmt, channels, n_frames, out_f64):
    """pcm_bytes uint8 device tensor holding ``n_frames`` interleaved frames -> out fp64 [n_frames] (mono)."""
    rc = _lib.load().wfl_pcm_to_f64(_ptr(pcm_bytes), fmt, channels, n_frames, _ptr(out_f64), _stream())
    _lib.check(rc, "wfl_pcm_to_f64")
    _count(1)


def stft_mag(y, n_fft, hop, power, out):
    """y fp32 [n] -> out fp32 [1 + n // hop, n_fft // 2 + 1] (|STFT| or |STFT|^2, centred, zero padded, periodic Hann)."""
    rc = _lib.load().wfl_stft_mag(_ptr(y), y.numel(), n_fft, hop, power, _ptr(out), _stream())
    _lib.check(rc, "wfl_stft_mag")
    _count(1)


def spectral_flux(S, flux):
    rc = _lib.load().wfl_spectral_flux(_ptr(S), S.shape[0], S.shape[1], _ptr(flux), _stream())
    _lib.check(rc, "wfl_spectral_flux")
    _count(1)


def mfcc_delta_mag(P, mel_fb, mel_span, dct, delta_mag):
    frames, bins = P.shape
    n_mels, n_mfcc = mel_fb.shape[0], dct.shape[0]
    db = torch.empty(frames, n_mels, device=P.device)
    mf = torch.empty(frames, n_mfcc, device=P.device)
    mx = torch.empty(1, dtype=torch.int32, device=P.device)
    rc = _lib.load().wfl_mfcc_delta_mag(_ptr(P), frames, bins, _ptr(mel_fb), _ptr(mel_span), n_mels, _ptr(dct), n_mfcc,
                                        _ptr(db), _ptr(mf), _ptr(mx), _ptr(delta_mag), _stream())
    _lib.check(rc, "wfl_mfcc_delta_mag")
    _count(3)
    return mf


def logmel_scratch(B, n_mels, device):
    """Scratch buffers of wfl_whisper_logmel: (planes f16, dft fp32 [B,3000,448], logspec fp32, clip max + filter spans)."""
    from .frontend import PLANE_SAMPLES
    return (torch.empty(2 * PLANE_SAMPLES * B + 4096, dtype=torch.float16, device=device),
            torch.empty(B, 3000, 448, device=device), torch.empty(B, 3000, n_mels, device=device),
            torch.empty(B + 258, device=device))


def whisper_logmel(wave, n_samples, basis_split, filters, n_mels, out, scratch):
    B = wave.shape[0]
    planes, dft, logspec, smax = scratch
    rc = _lib.load().wfl_whisper_logmel(_ptr(wave), wave.stride(0), n_samples, B, _ptr(basis_split), _ptr(filters), n_mels,
                                        _ptr(out), out.shape[-1], _ptr(planes), _ptr(dft), _ptr(logspec), _ptr(smax),
                                        _stream())
    _lib.check(rc, "wfl_whisper_logmel")
    _count(5)  # prep, DFT GEMM, filter spans, power + mel + log, normalise


def mel_power_frames(n_samples, hop):
    return 1 + n_samples // hop


def mel_power_scratch(B, n_samples, hop, device):
    """Scratch buffers of wfl_mel_power: (planes f16, dft fp32 [B, frames, 448], filter spans)."""
    frames = mel_power_frames(n_samples, hop)
    plane = (frames - 1 + (400 + hop - 1) // hop) * hop
    return (torch.empty(2 * plane * B + 4096, dtype=torch.float16, device=device),
            torch.empty(B, frames, 448, device=device), torch.empty(256, device=device))


def mel_power(wave, n_samples, hop, basis_split, filters, n_mels, out, scratch):
    """wave fp32 [B, >= n_samples] -> out fp32 [B, 1 + n_samples // hop, >= n_mels] (MelSpectrogram power, transposed)."""
    planes, dft, span = scratch
    rc = _lib.load().wfl_mel_power(_ptr(wave), wave.stride(0), n_samples, wave.shape[0], hop, _ptr(basis_split),
                                   _ptr(filters), n_mels, _ptr(out), out.stride(-2), _ptr(planes), _ptr(dft), _ptr(span),
                                   _stream())
    _lib.check(rc, "wfl_mel_power")
    _count(4)  # prep, DFT GEMM, filter spans, power + mel


def gather_cols(src, dst, groups, w_in, w_out):
    """dst[r, g*w_out + j] = src[r, g*w_in + j] (fp32)."""
    rows = src.numel() // (groups * w_in)
    _lib.check(_lib.load().wfl_gather_cols(_ptr(src), rows, groups, w_in, w_out, _ptr(dst), _stream()), "wfl_gather_cols")
    _count()


def decode_frames(logits2d, L, o_id, threshold, ids):
    rows = logits2d.shape[0]
    e0 = _t0("decode_frames")
    rc = _lib.load().wfl_decode_frames(_ptr(logits2d), rows, L, logits2d.stride(0), o_id, threshold, _ptr(ids), _stream())
    _lib.check(rc, "wfl_decode_frames")
    _t1(e0, "decode_frames", f"decode_frames L{L}", rows * (4 * L + 4))  # 4L + 4 B/frame
    _count()


def median_filter(ids_in, ids_out, lengths, k):
    n_clips, stride = ids_in.shape
    e0 = _t0("median_filter")
    rc = _lib.load().wfl_median_filter(_ptr(ids_in), _ptr(ids_out), _ptr(lengths), n_clips, stride, k, _stream())
    _lib.check(rc, "wfl_median_filter")
    _t1(e0, "median_filter", f"median_filter k{k}", n_clips * stride * 8)  # int32 in + out
    _count()


def bio_decode(ids, offsets, lengths, label_kind, label_ph, frame_duration, time_shift, segs, nseg):
    n_clips, stride = ids.shape
    e0 = _t0("bio_decode")
    rc = _lib.load().wfl_bio_decode(_ptr(ids), _ptr(offsets), _ptr(lengths), n_clips, stride, _ptr(label_kind),
                                    _ptr(label_ph), label_kind.numel(), frame_duration, _ptr(time_shift), _ptr(segs),
                                    _ptr(nseg), _stream())
    _lib.check(rc, "wfl_bio_decode")
    # ids (4 B/frame) + the two offsets (8 B/frame) read; 24 B per segment written (count unknown on the host: left out)
    _t1(e0, "bio_decode", "bio_decode", n_clips * stride * (4 + (8 if offsets is not None else 0)))
    _count()


def merge_segments(segs, nseg, clip_stride, file_clip_begin, n_files, ph_class, mode, out, nout):
    if mode not in MERGE_MODES:
        raise ValueError(f"Unsupported merge mode: {mode}")  # REF/utils.py:185
    e0 = _t0("merge_segments")
    rc = _lib.load().wfl_merge_segments(_ptr(segs), _ptr(nseg), clip_stride, _ptr(file_clip_begin), n_files,
                                        _ptr(ph_class), MERGE_MODES[mode], _ptr(out), _ptr(nout), _stream())
    _lib.check(rc, "wfl_merge_segments")
    _t1(e0, "merge_segments", f"merge_segments {mode}", 0)  # 2 x 24 B per segment; segment count lives on the device
    _count()


def htk_times(segs, n, start_out, end_out):
    _lib.check(_lib.load().wfl_htk_times(_ptr(segs), n, _ptr(start_out), _ptr(end_out), _stream()), "wfl_htk_times")
    _count()


def lstm_layer(gx, whh, B, T, H, y_f16=None, y_f32=None):
    """gx fp32 [B, T, 8H] (columns [dir][unit][gate]); whh f16 [2, 4H, H] -> y [B, T, 2H]."""
    e0 = _t0("lstm")
    rc = _lib.load().wfl_lstm_layer(_ptr(gx), _ptr(whh), B, T, H, _ptr(y_f16), _ptr(y_f32), _stream())
    _lib.check(rc, "wfl_lstm_layer")
    _t1(e0, "lstm", f"lstm H{H} B{B} T{T}", T)  # latency-bound: work = serial steps
    _count()
