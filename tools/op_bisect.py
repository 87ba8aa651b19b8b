"""Finds the first kernel launch whose output differs between two passes over the same input:
python tools/op_bisect.py cfg4 16 [model.key=value ...]   (wraps every ops.* call, hashes its output buffers after a sync)"""
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
LOG = []
SYNC = os.environ.get("BISECT_SYNC", "0") == "1"
def digest(t):
    """Hash of the output buffer, computed by kernels enqueued on the same stream (no host sync unless BISECT_SYNC=1:
    a sync after every launch hides races between consecutive launches)."""
    if SYNC:
        torch.cuda.synchronize()
    raw = t.contiguous().view(torch.int32 if t.element_size() * t.numel() % 4 == 0 else torch.uint8)
    return raw.flatten().to(torch.int64).cumsum(0)[-1] + (raw.flatten()[::97].to(torch.int64) * 31).sum()
def wrap(name, out_args):
    fn = getattr(ops, name)
    def inner(*a, **kw):
        r = fn(*a, **kw)
        outs = []
        for spec in out_args:
            t = kw.get(spec) if isinstance(spec, str) else (a[spec] if spec < len(a) else None)
            if torch.is_tensor(t):
                outs.append(digest(t))
        shape = tuple(kw.get("out", a[2] if len(a) > 2 else torch.empty(0)).shape) if name == "gemm" else ()
        LOG.append((name, kw.get("n"), kw.get("out_mode"), len(kw.get("shifts", ())), shape, tuple(outs)))
        return r
    setattr(ops, name, inner)
wrap("gemm", [2]); wrap("layernorm", ["out_f32", "out_f16"]); wrap("attention", [1]); wrap("wavlm_conv0", [6])
wrap("wavlm_gate", [9]); wrap("split_f16", [1]); wrap("lstm_layer", ["y_f16", "y_f32"]); wrap("rowdot_sigmoid", [3])
wrap("whisper_logmel", [5]); wrap("broadcast_rows", [1])
runs = []
for r in range(int(os.environ.get("RUNS", "4"))):
    LOG.clear()
    model(wave, lang)
    torch.cuda.synchronize()
    runs.append([e[:5] + (tuple(int(d.item()) for d in e[5]),) for e in LOG])
ref = runs[0]
for r, log in enumerate(runs[1:], 1):
    first = next((i for i, (a, b) in enumerate(zip(ref, log)) if a != b), None)
    if first is None:
        print(f"run{r}: all {len(log)} launches identical to run0")
    else:
        print(f"run{r}: first differing launch #{first} of {len(log)}: {log[first][:5]}  (previous: {log[first - 1][:5] if first else None})")
