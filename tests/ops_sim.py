"""TEST INFRASTRUCTURE ONLY -- a torch/CPU model of the C-ABI's *semantics* (include/wfl_b200.h), used to check the
host side of the product (weight packing, buffer layouts, launch sequence of ``wfl_asr_b200.engine``) in the
authoring container, which has no GPU.  ``install(monkeypatch)`` swaps the functions of ``wfl_asr_b200.ops`` that
the engine calls for the models below; nothing here is importable from the product path, and the GPU parity
tests (``-m gpu``) never use it -- they run the real kernels.

Each function follows the contract written in the header for the entry point of the same name: operands are
rounded to fp16 where the kernels take fp16, accumulation is fp32, tensor-map reads outside the declared
[a_rows, a_cols] extent return zero."""
import torch
import torch.nn.functional as F

ACT_NONE, ACT_GELU, ACT_RELU = 0, 1, 2
OUT_STORE_F16, OUT_STORE_F32, OUT_ADD_F32, OUT_GLU_F16 = 0, 1, 2, 3


def _strided(t, shape, strides):
    return torch.as_strided(t, shape, strides, t.storage_offset())


def _act(v, act):
    if act == ACT_GELU:
        return F.gelu(v)
    if act == ACT_RELU:
        return F.relu(v)
    return v


def gemm(a, w, out, *, n, slab_k, shifts=(0,), cols=(0,), a_rows, a_cols, a_row_stride, a_batch_stride=0, batches=1,
         m_rows=None, out_row_stride=None, out_batch_stride=0, bias=None, bias_batch_stride=0, act=ACT_NONE,
         out_mode=OUT_STORE_F16, alpha=1.0, tile_n=0, groups=1, a_col_group_stride=0, out_col_group_stride=0):
    assert a.dtype == torch.float16 and w.dtype == torch.float16 and slab_k % 64 == 0 and n % 8 == 0
    m_rows = a_rows if m_rows is None else m_rows
    if groups > 1:  # Conv1d(groups=G): group g = A columns + g * stride, W rows / bias [g*n, (g+1)*n), output columns + g * stride
        assert out_mode != OUT_GLU_F16 and not bias_batch_stride and w.shape[0] == groups * n
        for g in range(groups):
            og = torch.as_strided(out, (out.numel() - g * out_col_group_stride,), (1,),
                                  out.storage_offset() + g * out_col_group_stride)
            gemm(a, w[g * n:(g + 1) * n], og, n=n, slab_k=slab_k, shifts=shifts,
                 cols=[c + g * a_col_group_stride for c in cols], a_rows=a_rows, a_cols=a_cols,
                 a_row_stride=a_row_stride, a_batch_stride=a_batch_stride, batches=batches, m_rows=m_rows,
                 out_row_stride=out_row_stride, out_batch_stride=out_batch_stride,
                 bias=None if bias is None else bias[g * n:(g + 1) * n], act=act, out_mode=out_mode, alpha=alpha,
                 tile_n=tile_n)
        return
    assert w.shape == (n, len(shifts) * slab_k), (tuple(w.shape), n, len(shifts), slab_k)
    av = _strided(a, (batches, a_rows, a_cols), (a_batch_stride, a_row_stride, 1)).float()
    wf = w.float()
    acc = torch.zeros(batches, m_rows, n)
    t = torch.arange(m_rows)
    for s, (shift, col) in enumerate(zip(shifts, cols)):
        rows = t + shift
        ok = (rows >= 0) & (rows < a_rows)
        kk = min(slab_k, max(a_cols - col, 0))  # columns past a_cols are out of the tensor map: zero
        blk = torch.zeros(batches, m_rows, slab_k)
        if kk > 0 and ok.any():
            blk[:, ok, :kk] = av[:, rows[ok], col:col + kk]
        acc += blk @ wf[:, s * slab_k:(s + 1) * slab_k].T
    if bias is not None:
        if bias_batch_stride:
            acc += _strided(bias, (batches, 1, n), (bias_batch_stride, 0, 1))
        else:
            acc += bias[:n]
    out_cols = n // 2 if out_mode == OUT_GLU_F16 else n
    ors = out_cols if out_row_stride is None else out_row_stride
    ov = _strided(out, (batches, m_rows, out_cols), (out_batch_stride, ors, 1))
    if out_mode == OUT_GLU_F16:
        assert tile_n in (128, 256) and n % tile_n == 0 and out.dtype == torch.float16
        h = tile_n // 2
        blocks = acc.view(batches, m_rows, n // tile_n, 2, h)
        ov.copy_((blocks[..., 0, :] * torch.sigmoid(blocks[..., 1, :])).reshape(batches, m_rows, out_cols).half())
        return
    v = _act(acc, act)
    if out_mode == OUT_STORE_F16:
        assert out.dtype == torch.float16
        ov.copy_(v.clamp(-65504, 65504).half())
    elif out_mode == OUT_STORE_F32:
        assert out.dtype == torch.float32
        ov.copy_(v)
    else:
        assert out.dtype == torch.float32
        ov.add_(alpha * v)


def attention(qkv, out, *, B, T, H, hd, scale, q_col, k_col, v_col, rel_bias=None, gate=None):
    assert qkv.dtype == torch.float16 and out.dtype == torch.float16 and (rel_bias is None) == (gate is None)
    assert hd in (64, 256, 384, 512, 640), f"head_dim {hd} is not a built instantiation"
    assert rel_bias is None or hd == 64, "the gated relative-position bias is built for head_dim 64 only"
    x = qkv.float()
    parts = [x[:, :, c:c + H * hd].view(B, T, H, hd).transpose(1, 2) for c in (q_col, k_col, v_col)]
    scores = parts[0] @ parts[1].transpose(-1, -2) * scale
    if rel_bias is not None:  # gate[b,h,q] * rel_bias[h, k - q + T - 1]
        idx = torch.arange(T)[None, :] - torch.arange(T)[:, None] + T - 1
        scores = scores + gate.view(B, H, T, 1) * rel_bias[:, idx].unsqueeze(0)
    p = torch.softmax(scores, dim=-1)
    out.copy_((p @ parts[2]).transpose(1, 2).reshape(B, T, H * hd).half())


def layernorm(x, gamma, beta, *, out_f32=None, out_f16=None, gamma2=None, beta2=None, eps=1e-5, act_f16=ACT_NONE,
              rows=None):
    d = x.shape[-1]
    assert d % 4 == 0
    rows = x.numel() // d if rows is None else rows
    xv = x.reshape(-1)[:rows * d].view(rows, d)
    y = F.layer_norm(xv, (d,), gamma, beta, eps)
    if out_f16 is not None:
        z = _act(y, act_f16)
        if gamma2 is not None:
            z = F.layer_norm(z, (d,), gamma2, beta2, eps)
        out_f16.view(-1)[:rows * d].view(rows, d).copy_(z.half())
    if out_f32 is not None:
        out_f32.view(-1)[:rows * d].view(rows, d).copy_(y)


def split_f16(x, out):
    d = x.shape[-1]
    xv = x.reshape(-1, d)
    hi = xv.clamp(-65504, 65504).half()
    ov = out.view(-1)[:xv.shape[0] * 2 * d].view(-1, 2 * d)
    ov[:, :d] = hi
    ov[:, d:] = (xv - hi.float()).half()


def broadcast_rows(src, dst, batches):
    dst.view(batches, *src.shape).copy_(src.unsqueeze(0).expand(batches, *src.shape))


def rowdot_sigmoid(x_f16, w, b, out):
    d = x_f16.shape[-1]
    assert d % 8 == 0 and x_f16.dtype == torch.float16
    out.view(-1, w.shape[0]).copy_(torch.sigmoid(x_f16.reshape(-1, d).float() @ w.T + b))


def lstm_layer(gx, whh, B, T, H, y_f16=None, y_f32=None):
    """gx fp32 [B, T, 8H] columns [dir][unit][gate i,f,g,o]; whh f16 [2][4H][H] (gate-major rows)."""
    assert H in (192, 256, 384, 512, 640), f"hidden size {H} is not a built instantiation"
    assert gx.dtype == torch.float32 and whh.shape == (2, 4 * H, H)
    g4 = gx.reshape(-1)[:B * T * 8 * H].view(B, T, 2, H, 4)
    y = torch.zeros(B, T, 2 * H)
    for dirn in range(2):
        wt = whh[dirn].float().view(4, H, H)
        h = torch.zeros(B, H)
        c = torch.zeros(B, H)
        order = range(T) if dirn == 0 else range(T - 1, -1, -1)
        for t in order:
            pre = g4[:, t, dirn] + torch.einsum("bk,guk->bug", h.half().float(), wt)
            i_, f_, g_, o_ = pre.unbind(-1)
            c = torch.sigmoid(f_) * c + torch.sigmoid(i_) * torch.tanh(g_)
            h = torch.sigmoid(o_) * torch.tanh(c)
            y[:, t, dirn * H:(dirn + 1) * H] = h
    if y_f16 is not None:
        y_f16.view(-1)[:y.numel()].view_as(y).copy_(y.half())
    if y_f32 is not None:
        y_f32.view(-1)[:y.numel()].view_as(y).copy_(y)


def gather_cols(src, dst, groups, w_in, w_out):
    assert w_in % 4 == 0 and w_out % 4 == 0 and w_out <= w_in
    rows = src.numel() // (groups * w_in)
    dst.view(-1)[:rows * groups * w_out].view(rows, groups, w_out).copy_(src.view(rows, groups, w_in)[:, :, :w_out])


def wavlm_conv0(wave, n_samples, w, gamma, beta, norm_mode, out, out_batch_stride, scratch):
    """Conv1d(1, 512, k10, s5, no bias) + {0: GroupNorm(512, 512) over time | 1: input normalisation + LayerNorm} + GELU."""
    x = wave[:, :n_samples].float()
    if norm_mode == 1:
        x = (x - x.mean(dim=1, keepdim=True)) / torch.sqrt(x.var(dim=1, keepdim=True, unbiased=False) + 1e-7)
    y = F.conv1d(x[:, None], w[:, None, :], stride=5)  # [B, 512, T0]
    if norm_mode == 0:
        y = F.group_norm(y, 512, gamma, beta, 1e-5)
    else:
        y = F.layer_norm(y.transpose(1, 2), (512,), gamma, beta, 1e-5).transpose(1, 2)
    y = F.gelu(y).transpose(1, 2)  # [B, T0, 512]
    B, T0, _ = y.shape
    _strided(out, (B, T0, 512), (out_batch_stride, 512, 1)).copy_(y.half())


def wavlm_gate(x_f16, row_stride, B, T, H, hd, gw, gb, gconst, gate):
    x = _strided(x_f16, (B * T, H, hd), (row_stride, hd, 1)).float()
    proj = (x @ gw.T + gb).view(B, T, H, 2, 4).sum(-1)
    ga, gb_ = torch.sigmoid(proj).unbind(-1)
    gate.copy_((ga * (gb_ * gconst.view(1, 1, H) - 1.0) + 2.0).permute(0, 2, 1))


def mel_power_frames(n_samples, hop):
    return 1 + n_samples // hop


def mel_power_scratch(B, n_samples, hop, device):
    return (None, None, None)


def mel_power(wave, n_samples, hop, basis_split, filters, n_mels, out, scratch):
    """Windowed DFT through the packed split basis (hi + mid), exactly the contraction the kernel runs."""
    assert n_mels % 16 == 0 and 16 <= n_mels <= 128 and hop % 8 == 0 and n_samples > 200
    x = wave[:, :n_samples].float()
    xp = F.pad(x.unsqueeze(1), (200, 200), mode="reflect").squeeze(1)
    frames = xp.unfold(1, 400, hop)  # [B, T, 400]
    T = frames.shape[1]
    assert T == mel_power_frames(n_samples, hop)
    hi = frames.half()
    mid = (frames - hi.float()).half()
    bs = basis_split.float().view(448, 3, 448)  # rows = output column, slabs [W_hi | W_mid | W_hi]
    dft = hi.float() @ bs[:, 0, :400].T + hi.float() @ bs[:, 1, :400].T + mid.float() @ bs[:, 2, :400].T
    power = dft[..., 0:402:2] ** 2 + dft[..., 1:402:2] ** 2  # [B, T, 201]
    ov = _strided(out, (x.shape[0], T, n_mels), (out.stride(0), out.stride(-2), 1))
    ov.copy_(power @ filters)


def logmel_scratch(B, n_mels, device):
    return (None, None, None, None)


def whisper_logmel(wave, n_samples, basis_split, filters, n_mels, out, scratch):
    from oracle.torch_oracle import whisper_log_mel
    feats = whisper_log_mel(wave[:, :n_samples], n_mels).transpose(1, 2)  # [B, 3000, n_mels]
    out.zero_()
    out[:, :, :n_mels] = feats.half()


def install(monkeypatch):
    """Routes ``wfl_asr_b200.ops`` through this module and lets the engine accept CPU tensors (tests only)."""
    from wfl_asr_b200 import engine, ops
    for name in ("gemm", "attention", "layernorm", "split_f16", "broadcast_rows", "rowdot_sigmoid", "lstm_layer",
                 "gather_cols", "mel_power", "mel_power_scratch", "mel_power_frames", "logmel_scratch", "whisper_logmel",
                 "wavlm_conv0", "wavlm_gate"):
        monkeypatch.setattr(ops, name, globals()[name])
    monkeypatch.setattr(engine.Engine, "_require_device", lambda self, wave: None)
