"""Weight packing: reference ``state_dict`` tensors (fp32, reference layouts) -> the f16 layouts the
sm_100a kernels consume.  Pure tensor reshuffles and folds, done once at load time:

* Conv1d [out, in, k] -> [out][tap][in] so each tap is one K-slab of the implicit GEMM
* eval-mode BatchNorm1d folded into the preceding conv (REF/model.py:33-34)
* GLU value/gate rows interleaved per output tile (REF/model.py:31-32)
* lang_proj split into W_h and a per-language bias  W_e @ emb[l] + b  (REF/model.py:176-180)
* q/k/v projections concatenated (Whisper k_proj has no bias -> zeros)
* hi/lo f16 split of precision-critical tail weights (classifier)
"""
import torch


F16_MAX = 65504.0


def f16(t):
    """fp32 -> the library's 16-bit operand type (IEEE fp16), saturating like the kernels' own conversions."""
    return t.detach().float().clamp(-F16_MAX, F16_MAX).to(torch.float16).contiguous()


def conv_taps(weight, in_to=None):
    """[out, in, k] -> [out, k*in] with tap-major K (tap j multiplies input row t + j*dil - pad).  ``in_to``: zero-pad
    every tap's input channels to this width (one tap = one K slab, whose width must be a multiple of 64)."""
    o, i, k = weight.shape
    w = weight.permute(0, 2, 1)
    if in_to is not None and in_to != i:
        wp = w.new_zeros(o, k, in_to)
        wp[:, :, :i] = w
        w, i = wp, in_to
    return w.reshape(o, k * i).contiguous()


def pad_blocks(t, blocks, width, width_to, dim):
    """Dimension ``dim`` holds ``blocks`` consecutive blocks of ``width`` entries; zero-pad each block to ``width_to``
    (attention heads padded to a built head size, LSTM units padded to a built hidden size)."""
    if width == width_to:
        return t
    shape = list(t.shape)
    assert shape[dim] == blocks * width
    v = t.reshape(shape[:dim] + [blocks, width] + shape[dim + 1:])
    out_shape = list(v.shape)
    out_shape[dim + 1] = width_to
    out = v.new_zeros(out_shape)
    out.narrow(dim + 1, 0, width).copy_(v)
    return out.reshape(shape[:dim] + [blocks * width_to] + shape[dim + 1:])


def fit_size(n, built, what):
    """Smallest built kernel size >= n (operands are zero-padded up to it)."""
    for s in built:
        if s >= n:
            return s
    raise ValueError(f"{what} {n} exceeds the largest built size {built[-1]}")


def pad_k(weight2d, k_to):
    o, k = weight2d.shape
    if k == k_to:
        return weight2d
    out = weight2d.new_zeros(o, k_to)
    out[:, :k] = weight2d
    return out


def fold_batchnorm(conv_w, conv_b, gamma, beta, mean, var, eps=1e-5):
    """y = BN(conv(x)) == conv'(x): w' = w * g/sqrt(var+eps), b' = (b - mean) * g/sqrt(var+eps) + beta."""
    s = gamma / torch.sqrt(var + eps)
    return conv_w * s[:, None, None], (conv_b - mean) * s + beta


def interleave_glu(weight2d, bias, tile_n):
    """Rows [0,d) are GLU values, [d,2d) gates.  Reorder so each block of tile_n rows holds tile_n/2 values
    followed by their tile_n/2 gates -- the layout WFL_OUT_GLU_F16 expects."""
    two_d = weight2d.shape[0]
    d = two_d // 2
    h = tile_n // 2
    assert d % h == 0, "GLU width must be a multiple of tile_n/2"
    idx = []
    for t in range(d // h):
        idx += list(range(t * h, (t + 1) * h)) + list(range(d + t * h, d + (t + 1) * h))
    idx = torch.tensor(idx, device=weight2d.device)
    return weight2d[idx].contiguous(), (bias[idx].contiguous() if bias is not None else None)


def split_hi_lo(weight2d, k_to=None, parts="hhl"):
    """fp32 [n, k] -> f16 [n, 3k] = [hi | hi | lo]: pairs with A = [hi | lo] slabs (cols 0, k, 0).  ``k_to``: zero-pad
    each slab to this width.  ``parts="hl"``: [hi | lo] only (an f16 A operand against split weights, cols 0, 0)."""
    hi = f16(weight2d)
    lo = f16(weight2d.float() - hi.float())
    if k_to is not None:
        hi, lo = pad_k(hi, k_to), pad_k(lo, k_to)
    return torch.cat([hi if p == "h" else lo for p in parts], dim=1).contiguous()


def split_hi_lo_taps(weight, in_to):
    """Conv1d weight fp32 [out, in, k] -> f16 [out, k * 3 * in_to]: per tap the three slabs [hi | hi | lo] (each tap's
    input channels zero-padded to ``in_to``), pairing with A = [hi | lo] rows shifted by the tap."""
    o, i, k = weight.shape
    return torch.cat([split_hi_lo(weight[:, :, j], in_to) for j in range(k)], dim=1).contiguous()


def pad_rows(weight2d, rows_to):
    n, k = weight2d.shape
    if n == rows_to:
        return weight2d
    out = weight2d.new_zeros(rows_to, k)
    out[:n] = weight2d
    return out
