"""Detects reads of never-written workspace memory: after a first pass (which allocates the workspaces) every workspace
tensor is filled with NaN, the pass is repeated, and the first launch whose output contains NaN is reported.
python tools/poison_check.py cfg4 16 [model.key=value ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from wfl_asr_b200 import ops, synth
from wfl_asr_b200.model import BIOPhonemeTagger
wl, B = sys.argv[1], int(sys.argv[2])
dev = torch.device("cuda:0")
cfg = synth.workload_config(wl)
for kv in sys.argv[3:]:
    k, v = kv.split("=")
    cfg["model"][k] = type(cfg["model"].get(k, 0))(eval(v))
model = synth.bench_model(BIOPhonemeTagger, cfg, synth.synth_labels(30)).to(dev).eval()
secs = synth.WORKLOADS[wl]["seconds"]
base = [synth.synth_wave(700 + i, secs) for i in range(4)]
wave = torch.from_numpy(np.stack([base[i % 4] * (0.5 + 0.5 * ((i * 7) % 11) / 11.0) for i in range(B)]).astype(np.float32)).to(dev)
lang = torch.tensor([i % 2 for i in range(B)], device=dev)
l0, o0 = model(wave, lang)
l0, o0 = l0.clone(), o0.clone()
eng = model.engine()
n = 0
for key, ws in eng._ws.items():
    for name, t in ws.items():
        ts = t if isinstance(t, (tuple, list)) else [t]
        for u in ts:
            if torch.is_tensor(u) and u.is_floating_point():
                u.fill_(float("nan"))
                n += 1
print(f"poisoned {n} workspace tensors")
LOG = []
def nan_count(t):
    torch.cuda.synchronize()
    return int(torch.isnan(t.float()).sum().item()) if t.is_floating_point() else 0
def wrap(name, out_args):
    fn = getattr(ops, name)
    def inner(*a, **kw):
        r = fn(*a, **kw)
        for spec in out_args:
            t = kw.get(spec) if isinstance(spec, str) else (a[spec] if spec < len(a) else None)
            if torch.is_tensor(t):
                LOG.append((name, kw.get("n"), kw.get("out_mode"), len(kw.get("shifts", ())), kw.get("groups", 1), tuple(t.shape), nan_count(t), t.numel()))
        return r
    setattr(ops, name, inner)
wrap("gemm", [2]); wrap("layernorm", ["out_f32", "out_f16"]); wrap("attention", [1]); wrap("wavlm_conv0", [6])
wrap("wavlm_gate", [9]); wrap("split_f16", [1]); wrap("lstm_layer", ["y_f16", "y_f32"]); wrap("rowdot_sigmoid", [3])
wrap("whisper_logmel", [5]); wrap("broadcast_rows", [1])
l1, o1 = model(wave, lang)
print("NaN in logits after poisoning:", int(torch.isnan(l1).sum().item()), "of", l1.numel(), "| equal to first pass:", torch.equal(l0, l1))
for i, e in enumerate(LOG):
    print(i, e)
    if i > 40: break
