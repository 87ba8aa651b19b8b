"""Bitwise run-to-run check of single kernels at bench shapes: python tools/op_determinism.py [iters]
(GEMM modes incl. CTA-pair / grouped / conv taps, attention incl. head dim 512, BiLSTM 8 and 16 clips per cluster, LayerNorm)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wfl_asr_b200 import ops
dev = torch.device("cuda:0")
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 100
g = torch.Generator().manual_seed(5)
def rnd(*s, scale=1.0): return (torch.randn(*s, generator=g) * scale).to(dev)
def check(name, fn, outs):
    ref, bad = None, 0
    for it in range(iters):
        fn()
        cur = [o.clone() for o in outs]
        if ref is None: ref = cur
        elif any(not torch.equal(a.view(torch.uint8), b.view(torch.uint8)) for a, b in zip(cur, ref)): bad += 1
    print(f"{name}: {bad} of {iters - 1} launches differ")
B, T = 32, 1500
M = B * T
for (N, K, mode, act) in [(1536, 512, 0, 0), (2048, 512, 0, 1), (512, 2048, 2, 0), (512, 512, 2, 0), (1024, 512, 3, 0), (3072, 768, 1, 0)]:
    a = rnd(M, K).half(); w = rnd(N, K, scale=K ** -0.5).half(); bias = rnd(N)
    base = rnd(M, N) if mode == 2 else None
    out = torch.zeros(M, N // 2 if mode == 3 else N, device=dev, dtype=torch.float32 if mode in (1, 2) else torch.float16)
    def fn():
        if mode == 2: out.copy_(base)
        ops.linear(a, w, out, bias=bias, out_mode=mode, act=act, alpha=0.5 if mode == 2 else 1.0, tile_n=256 if mode == 3 else 0)
    check(f"gemm N{N} K{K} mode{mode} act{act}", fn, [out])
x = rnd(B, T, 512).half(); wt = rnd(512, 31 * 512, scale=(31 * 512) ** -0.5).half(); o = torch.empty(B, T, 512, device=dev, dtype=torch.float16); bz = rnd(512)
check("conv31 d512", lambda: ops.gemm(x, wt, o, n=512, slab_k=512, shifts=[j - 15 for j in range(31)], cols=[0] * 31, a_rows=T, a_cols=512,
      a_row_stride=512, a_batch_stride=T * 512, batches=B, m_rows=T, out_row_stride=512, out_batch_stride=T * 512, bias=bz, act=ops.ACT_GELU), [o])
for (H, hd, Bq) in [(8, 64, 32), (2, 256, 32), (2, 384, 16), (2, 512, 16), (2, 640, 8)]:
    d = H * hd
    qkv = rnd(Bq, T, 3 * d, scale=0.5).half(); out = torch.empty(Bq, T, d, device=dev, dtype=torch.float16)
    check(f"attention hd{hd} H{H} B{Bq}", lambda: ops.attention(qkv, out, B=Bq, T=T, H=H, hd=hd, scale=hd ** -0.5, q_col=0, k_col=d, v_col=2 * d), [out])
for (Hh, Bl) in [(384, 32), (384, 64), (256, 70), (640, 16)]:
    gx = rnd(Bl, 300, 8 * Hh, scale=0.5); whh = rnd(2, 4 * Hh, Hh, scale=Hh ** -0.5).half(); y = torch.empty(Bl, 300, 2 * Hh, device=dev)
    check(f"lstm H{Hh} B{Bl}", lambda: ops.lstm_layer(gx, whh, Bl, 300, Hh, y_f32=y), [y])
xl = rnd(M, 512); gm = rnd(512); bt = rnd(512); o16 = torch.empty(M, 512, device=dev, dtype=torch.float16)
check("layernorm d512", lambda: ops.layernorm(xl, gm, bt, out_f16=o16), [o16])
