"""Where does the fp16-operand logit error come from?  (CPU experiment, test infrastructure: uses tests/ops_sim.py.)

Runs wfl_asr_b200.engine over the torch model of the C ABI with every 16-bit buffer stored in fp32 and the rounding
to fp16 applied (or not) per PRODUCER class, so that one class at a time can be given exact operands:
    w      weight packing (packing.f16)          ln     LayerNorm outputs
    gelu   GELU/ReLU GEMM outputs (fc1, ff l1, conv-31, conv1)    lin   plain f16 GEMM outputs (qkv, lang_proj)
    glu    GLU output                            attn   attention context
    split  [hi | lo] splits (hi only; lo kept)   mel    log-mel features
Prints the rms / max logit error and tag mismatches against the fp32 oracle for "all rounded" (= the kernels'
arithmetic) and for each class made exact.   python tools/precision_attribution.py [case ...]
"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_forward_golden as mfg  # noqa: E402
import ops_sim  # noqa: E402
from oracle import torch_oracle as to  # noqa: E402
from wfl_asr_b200 import engine, ops, packing  # noqa: E402

EXACT = set()  # producer classes whose outputs are NOT rounded


def r16(x, kind):
    x = x.float()
    return x if kind in EXACT else x.clamp(-65504, 65504).half().float()


class FakeHalf:
    """tensor.half() on fp32-stored buffers is a no-op in this experiment (rounding happened at the producer)."""


def gemm(a, w, out, *, n, slab_k, shifts=(0,), cols=(0,), a_rows, a_cols, a_row_stride, a_batch_stride=0, batches=1,
         m_rows=None, out_row_stride=None, out_batch_stride=0, bias=None, bias_batch_stride=0, act=0, out_mode=0,
         alpha=1.0, tile_n=0, groups=1, a_col_group_stride=0, out_col_group_stride=0):
    m_rows = a_rows if m_rows is None else m_rows
    if groups > 1:
        for g in range(groups):
            og = torch.as_strided(out, (out.numel() - g * out_col_group_stride,), (1,), out.storage_offset() + g * out_col_group_stride)
            gemm(a, w[g * n:(g + 1) * n], og, n=n, slab_k=slab_k, shifts=shifts, cols=[c + g * a_col_group_stride for c in cols],
                 a_rows=a_rows, a_cols=a_cols, a_row_stride=a_row_stride, a_batch_stride=a_batch_stride, batches=batches,
                 m_rows=m_rows, out_row_stride=out_row_stride, out_batch_stride=out_batch_stride,
                 bias=None if bias is None else bias[g * n:(g + 1) * n], act=act, out_mode=out_mode, alpha=alpha, tile_n=tile_n)
        return
    S = ops_sim._strided
    av = S(a, (batches, a_rows, a_cols), (a_batch_stride, a_row_stride, 1)).float()
    wf = w.float()
    acc = torch.zeros(batches, m_rows, n)
    t = torch.arange(m_rows)
    for s, (shift, col) in enumerate(zip(shifts, cols)):
        rows = t + shift
        ok = (rows >= 0) & (rows < a_rows)
        kk = min(slab_k, max(a_cols - col, 0))
        blk = torch.zeros(batches, m_rows, slab_k)
        if kk > 0 and ok.any():
            blk[:, ok, :kk] = av[:, rows[ok], col:col + kk]
        acc += blk @ wf[:, s * slab_k:(s + 1) * slab_k].T
    if bias is not None:
        acc += S(bias, (batches, 1, n), (bias_batch_stride, 0, 1)) if bias_batch_stride else bias[:n]
    out_cols = n // 2 if out_mode == 3 else n
    ors = out_cols if out_row_stride is None else out_row_stride
    ov = S(out, (batches, m_rows, out_cols), (out_batch_stride, ors, 1))
    if out_mode == 3:
        h = tile_n // 2
        blocks = acc.view(batches, m_rows, n // tile_n, 2, h)
        ov.copy_(r16((blocks[..., 0, :] * torch.sigmoid(blocks[..., 1, :])).reshape(batches, m_rows, out_cols), "glu"))
        return
    v = ops_sim._act(acc, act)
    if out_mode == 0:
        ov.copy_(r16(v, "gelu" if act else "lin"))
    elif out_mode == 1:
        ov.copy_(v)
    else:
        ov.add_(alpha * v)


def attention(qkv, out, *, B, T, H, hd, scale, q_col, k_col, v_col, rel_bias=None, gate=None):
    x = qkv.float()
    parts = [x[:, :, c:c + H * hd].view(B, T, H, hd).transpose(1, 2) for c in (q_col, k_col, v_col)]
    scores = parts[0] @ parts[1].transpose(-1, -2) * scale
    if rel_bias is not None:
        idx = torch.arange(T)[None, :] - torch.arange(T)[:, None] + T - 1
        scores = scores + gate.view(B, H, T, 1) * rel_bias[:, idx].unsqueeze(0)
    p = r16(torch.softmax(scores, dim=-1), "p")
    out.copy_(r16((p @ parts[2]).transpose(1, 2).reshape(B, T, H * hd), "attn"))


def layernorm(x, gamma, beta, *, out_f32=None, out_f16=None, gamma2=None, beta2=None, eps=1e-5, act_f16=0, rows=None):
    d = x.shape[-1]
    rows = x.numel() // d if rows is None else rows
    xv = x.reshape(-1)[:rows * d].view(rows, d)
    y = F.layer_norm(xv, (d,), gamma, beta, eps)
    if out_f16 is not None:
        z = ops_sim._act(y, act_f16)
        if gamma2 is not None:
            z = F.layer_norm(z, (d,), gamma2, beta2, eps)
        nm = LN_NAMES.get((gamma2 if gamma2 is not None else gamma).data_ptr(), "?")
        out_f16.view(-1)[:rows * d].view(rows, d).copy_(z if nm.startswith(LN_EXACT_PREFIX) and LN_EXACT_PREFIX else r16(z, "ln"))
    if out_f32 is not None:
        out_f32.view(-1)[:rows * d].view(rows, d).copy_(y)


def split_f16(x, out):
    d = x.shape[-1]
    xv = x.reshape(-1, d)
    hi = r16(xv, "split")
    ov = out.view(-1)[:xv.shape[0] * 2 * d].view(-1, 2 * d)
    ov[:, :d] = hi
    ov[:, d:] = r16(xv - hi, "split_lo")


def whisper_logmel(wave, n_samples, basis_split, filters, n_mels, out, scratch):
    feats = to.whisper_log_mel(wave[:, :n_samples], n_mels).transpose(1, 2)
    out.zero_()
    out[:, :, :n_mels] = r16(feats, "mel")


def rowdot_sigmoid(x_f16, w, b, out):
    d = x_f16.shape[-1]
    out.view(-1, w.shape[0]).copy_(torch.sigmoid(x_f16.reshape(-1, d).float() @ w.T + b))


LN_NAMES = {}
LN_EXACT_PREFIX = ()
W_EXACT_PREFIX = ()  # weights whose engine name starts with one of these are kept exact
LN_NAMES = {}
LN_EXACT_PREFIX = ()  # LayerNorms (by engine name of their gamma) whose f16 output is kept exact
_CUR = [""]
V_ONLY = [False]


class Eng32(engine.Engine):
    def _put(self, name, t, dtype=None):
        _CUR[0] = name
        super()._put(name, t, dtype)
        _CUR[0] = ""

    def _buffers(self, B, T, n_samples=None):
        ws = super()._buffers(B, T, n_samples)
        for k, v in list(ws.items()):
            if torch.is_tensor(v) and v.dtype == torch.float16:
                ws[k] = torch.zeros(v.shape, dtype=torch.float32)
        return ws


def run(name):
    cfg, labels, sd, wave, lang = mfg.case_inputs(name)
    wave, lang = wave[:1], lang[:1]
    ref_l, _ = to.forward(wave, sd, cfg, lang)
    scale = ref_l.abs().max().item()
    rows = []
    top = ([], ["w"], ["ln"], ["gelu"], ["lin"], ["glu"], ["attn", "p"], ["split"], ["mel"],
           ["w", "ln", "gelu", "lin", "glu", "attn", "p", "split", "mel", "split_lo"])
    if os.environ.get("WFL_ATTR_QUICK"):
        top = ([],)
    for exact in top:
        EXACT.clear()
        EXACT.update(exact)
        eng = Eng32(sd, cfg, len(labels), torch.device("cpu"))
        l, _ = eng.forward(wave, lang)
        err = (l - ref_l).abs()
        rows.append((",".join(exact) or "all rounded", err.pow(2).mean().sqrt().item() / scale, err.max().item() / scale,
                     int((l.argmax(-1) != ref_l.argmax(-1)).sum())))
    global W_EXACT_PREFIX
    groups = [("enc",), ("conf0",), ("conf1",), ("conf2",), ("conf3",), ("lang", "cls", "off", "dil", "lstm")]
    groups += [tuple(f"conf{i}.{p}" for i in range(8)) for p in ("ff1", "ff2", "attn", "pw1", "conv.", "pw2")]
    if os.environ.get("WFL_ATTR_QUICK"):
        groups = []
    for pref in groups:
        EXACT.clear()
        W_EXACT_PREFIX = pref
        eng = Eng32(sd, cfg, len(labels), torch.device("cpu"))
        l, _ = eng.forward(wave, lang)
        err = (l - ref_l).abs()
        rows.append(("w:" + "|".join(pref)[:36], err.pow(2).mean().sqrt().item() / scale, err.max().item() / scale,
                     int((l.argmax(-1) != ref_l.argmax(-1)).sum())))
    W_EXACT_PREFIX = ()
    global LN_EXACT_PREFIX
    confs = tuple(f"conf{i}." for i in range(8))
    combos = [((), (), True), (tuple(f"enc{i}.out" for i in range(12)), (), True),
              (tuple(f"enc{i}.out" for i in range(12)) + tuple(f"enc{i}.qkv" for i in range(12)), (), False),
              (tuple(f"enc{i}.out" for i in range(12)) + ("conf0.ff1",), (), True)]
    for wp, lp, vo in combos:
        EXACT.clear()
        V_ONLY[0] = vo
        W_EXACT_PREFIX, LN_EXACT_PREFIX = wp, lp
        eng = Eng32(sd, cfg, len(labels), torch.device("cpu"))
        LN_NAMES.clear()
        LN_NAMES.update({v.data_ptr(): k for k, v in eng.W.items() if k.endswith(".g")})
        l, _ = eng.forward(wave, lang)
        err = (l - ref_l).abs()
        rows.append((("w:" + "|".join(wp))[:20] + " ln:" + "|".join(lp)[:16], err.pow(2).mean().sqrt().item() / scale, err.max().item() / scale,
                     int((l.argmax(-1) != ref_l.argmax(-1)).sum())))
    W_EXACT_PREFIX, LN_EXACT_PREFIX = (), ()
    V_ONLY[0] = False
    base = rows[0][1]
    print(f"== {name}: {ref_l.shape[1]} frames, logit scale {scale:.3f}")
    for label, rms, mx, mism in rows:
        share = max(0.0, 1.0 - (rms / base) ** 2)
        print(f"  exact [{label:>40s}]  rms {rms:.3e}  max {mx:.3e}  mismatched frames {mism}  (variance share of this class {share:5.1%})")


if __name__ == "__main__":
    mp = type("MP", (), {"setattr": staticmethod(setattr)})()
    ops_sim.install(mp)
    for fn in (gemm, attention, layernorm, split_f16, whisper_logmel, rowdot_sigmoid):
        setattr(ops, fn.__name__, fn)
    def _f16(t):
        t = t.detach().float()
        if W_EXACT_PREFIX and _CUR[0].startswith(W_EXACT_PREFIX):
            return t.contiguous()
        r = r16(t, "w")
        if V_ONLY[0] and _CUR[0].endswith(("attn.in.w", "qkv.w")):  # value rows of the packed in_proj exact, q / k rows rounded
            n = t.shape[0] // 3
            r[2 * n:] = t[2 * n:]
        return r.contiguous()
    packing.f16 = _f16
    mfg.CASES["cfg3_depth12"] = (dict(whisper_model="openai/whisper-small", num_conformer_layers=4), 12, 1, 30.0, 33)
    mfg.CASES["cfg2_depth6"] = (dict(enable_bilstm=False, enable_dilated_conv=False, num_conformer_layers=4), 6, 1, 30.0, 34)
    for name in (sys.argv[1:] or ["whisper_base_cfg2"]):
        run(name)
