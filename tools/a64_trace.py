"""Per-tile clock64 timeline of one attention64 CTA (variant build with -DWFL_A64_TRACE):
   python tools/build_variant.py a64trace --src attention64.cu -DWFL_A64_TRACE
   WFL_LIB=variants/libwfl_a64trace.so python tools/a64_trace.py"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from wfl_asr_b200 import _lib, ops
dev = torch.device("cuda:0")
B, T, H, hd = 32, 1500, 8, 64
d = H * hd
g = torch.Generator().manual_seed(0)
qkv = (torch.randn(B, T, 3 * d, generator=g) * 0.5).to(dev).half()
out = torch.empty(B, T, d, device=dev, dtype=torch.float16)
for _ in range(3):
    ops.attention(qkv, out, B=B, T=T, H=H, hd=hd, scale=hd ** -0.5, q_col=0, k_col=d, v_col=2 * d)
torch.cuda.synchronize()
buf = np.zeros(20 * 16 * 8, dtype=np.int64)
lib = _lib.load()
lib.wfl_debug_a64_trace.argtypes = [ctypes.c_void_p]
rc = lib.wfl_debug_a64_trace(buf.ctypes.data_as(ctypes.c_void_p))
assert rc == 0, rc
tr = buf.reshape(20, 16, 8)
t0 = tr[tr > 0].min()
rel = np.where(tr > 0, tr - t0, -1)
n_kv = 12
print("MMA threads (warp 1 = query tile 0, warp 2 = tile 1): per tile j: s_empty seen, QK(j+1) issued, p_full0 seen, PV0 issued, p_full1 seen, PV1 issued")
for w in (1, 2):
    for j in range(n_kv):
        print(f"  w{w} j{j:2d} " + " ".join(f"{rel[w, j, e]:7d}" for e in range(6)))
print("softmax warps: per tile j: start wait s_full, got it, S loaded, after max exchange, exps0 done, pv_done0 ok, exps1 done, pv_done1 ok")
for w in (4, 8, 12, 16):
    for j in range(n_kv):
        print(f"  w{w:2d} j{j:2d} " + " ".join(f"{rel[w, j, e]:7d}" for e in range(8)))
