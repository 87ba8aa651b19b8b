"""Runs one workload's forward twice on the same input and reports whether logits/offsets are bitwise equal:
python tools/determinism_check.py cfg3 64   (env: WFL_LSTM_NB, WFL_GEMM_PAIR to isolate kernels)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from wfl_asr_b200 import synth
from wfl_asr_b200.model import BIOPhonemeTagger
wl, B = sys.argv[1], int(sys.argv[2])
dev = torch.device("cuda:0")
cfg = synth.workload_config(wl)
if len(sys.argv) > 3:
    for kv in sys.argv[3:]:
        k, v = kv.split("=")
        cfg["model"][k] = type(cfg["model"].get(k, 0))(eval(v))
# DET_FENCE_AFTER=gemm,attention,...: enqueue a tiny unrelated kernel after every launch of those ops (separates
# adjacent launches without a host sync: localises races between consecutive kernels)
fence_after = [f for f in os.environ.get("DET_FENCE_AFTER", "").split(",") if f]
if fence_after:
    from wfl_asr_b200 import ops
    dummy = torch.zeros(32, device=dev)
    for name in fence_after:
        fn = getattr(ops, name)
        def inner(*a, _fn=fn, **kw):
            r = _fn(*a, **kw)
            dummy.add_(1.0)
            return r
        setattr(ops, name, inner)
model = synth.bench_model(BIOPhonemeTagger, cfg, synth.synth_labels(30)).to(dev).eval()
base = [synth.synth_wave(700 + i, 30.0) for i in range(4)]
wave = torch.from_numpy(np.stack([base[i % 4] * (0.5 + 0.5 * ((i * 7) % 11) / 11.0) for i in range(B)]).astype(np.float32)).to(dev)
lang = torch.tensor([i % 2 for i in range(B)], device=dev)
outs = []
for r in range(6):
    l, o = model(wave, lang)
    outs.append((l.clone(), o.clone()))
for r in range(1, 6):
    dl = (outs[r][0] - outs[0][0]).abs()
    if dl.max().item() > 0 and os.environ.get("DET_VERBOSE"):
        fr = dl.amax(dim=2)  # [B, T]
        for c in sorted(set(dl.amax(dim=(1, 2)).nonzero().flatten().tolist()))[:4]:
            nz = fr[c].nonzero().flatten()
            print(f"   clip {c}: {nz.numel()} of {fr.shape[1]} frames differ, first {nz[:6].tolist()} last {nz[-6:].tolist()}, "
                  f"max at frame {int(fr[c].argmax())}")
    print(f"{wl} B={B} {' '.join(sys.argv[3:])} run{r} vs run0: equal={torch.equal(outs[r][0], outs[0][0])} max|d|={dl.max().item():.3e} "
          f"differing clips={sorted(set(dl.amax(dim=(1, 2)).nonzero().flatten().tolist()))[:12]}")
