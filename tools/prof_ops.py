"""Runs single ops at bench shapes (for ncu captures): python tools/prof_ops.py attn64 attn256 ln logmel gemm"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wfl_asr_b200 import ops
from wfl_asr_b200.frontend import whisper_frontend_constants
dev = torch.device("cuda:0")
B, T, d = int(os.environ.get("PROF_B", "32")), int(os.environ.get("PROF_T", "1500")), int(os.environ.get("PROF_D", "512"))
which = sys.argv[1:] or ["attn64"]
reps = int(os.environ.get("REPS", "2"))
g = torch.Generator().manual_seed(0)
def t_ms(fn):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for w in which:
    if w == "attn64b":  # WavLM: head dim 64 with the gated relative-position bias (PROF_B / PROF_T / PROF_D)
        H = d // 64
        qkv = (torch.randn(B, T, 3 * d, generator=g) * 0.5).to(dev).half()
        out = torch.empty(B, T, d, device=dev, dtype=torch.float16)
        tab = (torch.randn(H, 2 * T - 1, generator=g) * 0.5).to(dev)
        gate = (torch.rand(B, H, T, generator=g) * 2).to(dev)
        ms = t_ms(lambda: ops.attention(qkv, out, B=B, T=T, H=H, hd=64, scale=0.125, q_col=0, k_col=d, v_col=2 * d,
                                        rel_bias=tab, gate=gate))
        print(f"{w} B{B} T{T} H{H} ({os.environ.get('WFL_ATTN64', 'default')}): {ms:.3f} ms  {4.0 * T * T * d * B / ms / 1e9:.1f} TFLOP/s")
    elif w.startswith("attn"):
        hd = int(w[4:]); H = d // hd if hd <= 256 else 2
        dd = H * hd
        qkv = (torch.randn(B, T, 3 * dd, generator=g) * 0.5).to(dev).half()
        out = torch.empty(B, T, dd, device=dev, dtype=torch.float16)
        ms = t_ms(lambda: ops.attention(qkv, out, B=B, T=T, H=H, hd=hd, scale=hd ** -0.5, q_col=0, k_col=dd, v_col=2 * dd))
        print(f"{w}: {ms:.3f} ms  {4.0 * T * T * dd * B / ms / 1e9:.1f} TFLOP/s")
    elif w == "ln":
        x = torch.randn(B * T, d, generator=g).to(dev); gm = torch.ones(d, device=dev); bt = torch.zeros(d, device=dev)
        o = torch.empty(B * T, d, device=dev, dtype=torch.float16)
        ms = t_ms(lambda: ops.layernorm(x, gm, bt, out_f16=o))
        print(f"ln: {ms:.4f} ms  {B * T * d * 6 / ms / 1e6:.0f} GB/s")
    elif w == "logmel":
        wave = (torch.randn(B, 480000, generator=g) * 0.1).to(dev)
        basis, filt = whisper_frontend_constants(80, dev)
        out = torch.empty(B, 3000, 128, device=dev, dtype=torch.float16)
        scratch = ops.logmel_scratch(B, 80, dev)
        ms = t_ms(lambda: ops.whisper_logmel(wave, 480000, basis, filt, 80, out, scratch))
        print(f"logmel: {ms:.3f} ms")
    elif w.startswith("conv"):
        # convK_d[_tile]: Conformer conv module shape, K taps over [B, T, d]
        parts = w[4:].split("_"); taps, dd = int(parts[0]), int(parts[1]); tile = int(parts[2]) if len(parts) > 2 else 0
        xin = torch.randn(B, T, dd, generator=g).to(dev).half()
        wt = (torch.randn(dd, taps * dd, generator=g) * (taps * dd) ** -0.5).to(dev).half()
        o = torch.empty(B, T, dd, device=dev, dtype=torch.float16); bias = torch.zeros(dd, device=dev)
        pad = (taps - 1) // 2
        ms = t_ms(lambda: ops.gemm(xin, wt, o, n=dd, slab_k=dd, shifts=[j - pad for j in range(taps)], cols=[0] * taps,
                                   a_rows=T, a_cols=dd, a_row_stride=dd, a_batch_stride=T * dd, batches=B, m_rows=T,
                                   out_row_stride=dd, out_batch_stride=T * dd, bias=bias, act=ops.ACT_GELU, tile_n=tile))
        print(f"{w}: {ms:.4f} ms  {2.0 * B * T * dd * taps * dd / ms / 1e9:.1f} TFLOP/s")
    elif w.startswith("lstm"):
        # lstmH_B : one bidirectional layer, T = 1500
        H, Bl = (int(v) for v in w[4:].split("_"))
        gx = (torch.randn(Bl, T, 8 * H, generator=g) * 0.5).to(dev)
        whh = (torch.randn(2, 4 * H, H, generator=g) * H ** -0.5).to(dev).half()
        y = torch.empty(Bl, T, 2 * H, device=dev)
        ms = t_ms(lambda: ops.lstm_layer(gx, whh, Bl, T, H, y_f32=y))
        print(f"{w}: {ms:.3f} ms  {ms * 1e3 / T:.3f} us/step")
    elif w.startswith("gemm"):
        # gemmN_K[_mode]
        parts = w[4:].split("_"); N, K = int(parts[0]), int(parts[1]); mode = int(parts[2]) if len(parts) > 2 else 0
        tile = int(parts[3]) if len(parts) > 3 else 0
        act = int(parts[4]) if len(parts) > 4 else (ops.ACT_GELU if mode == 0 else 0)
        a = torch.randn(B * T, K, generator=g).to(dev).half(); wt = (torch.randn(N, K, generator=g) * K ** -0.5).to(dev).half()
        o = torch.zeros(B * T, N, device=dev, dtype=torch.float32 if mode in (1, 2) else torch.float16)
        bias = torch.zeros(N, device=dev)
        ms = t_ms(lambda: ops.linear(a, wt, o, bias=bias, out_mode=mode, act=act, tile_n=tile))
        print(f"{w}: {ms:.4f} ms  {2.0 * B * T * N * K / ms / 1e9:.1f} TFLOP/s")
