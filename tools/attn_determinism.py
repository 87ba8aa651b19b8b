"""Repeats one attention launch on fixed inputs and reports launches whose output differs bitwise from the first:
python tools/attn_determinism.py [bias|nobias] [T] [B] [H] [hd]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wfl_asr_b200 import ops
mode = sys.argv[1] if len(sys.argv) > 1 else "bias"
T, B, H, hd = (int(v) for v in (sys.argv[2:6] + ["1499", "16", "16", "64"][len(sys.argv) - 2:]))
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(1)
d = H * hd
qkv = (torch.randn(B, T, 3 * d, generator=g) * 0.5).to(dev).half()
rel = torch.randn(H, 2 * T - 1, generator=g).to(dev) if mode == "bias" else None
gate = (1.0 + 0.3 * torch.randn(B, H, T, generator=g)).to(dev) if mode == "bias" else None
out = torch.empty(B, T, d, device=dev, dtype=torch.float16)
ref, bad = None, 0
for it in range(int(os.environ.get("ITERS", "200"))):
    out.fill_(float("nan"))
    ops.attention(qkv, out, B=B, T=T, H=H, hd=hd, scale=hd ** -0.5, q_col=0, k_col=d, v_col=2 * d, rel_bias=rel, gate=gate)
    cur = out.clone()
    if ref is None:
        ref = cur
    elif not torch.equal(cur.view(torch.int16), ref.view(torch.int16)):
        bad += 1
        diff = (cur.float() - ref.float()).abs()
        rows = diff.amax(dim=2).nonzero()
        if bad <= 5:
            print(f"iter {it}: {rows.shape[0]} rows differ; clips {sorted(set(rows[:, 0].tolist()))[:6]} frames "
                  f"{rows[:4, 1].tolist()}..{rows[-3:, 1].tolist()} max {diff.max().item():.3e} nan {int(torch.isnan(cur.float()).sum())}")
print(f"{mode} T={T} B={B} H={H} hd={hd}: {bad} differing launches")
