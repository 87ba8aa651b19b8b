"""Times one labeling pass (forward + post-processing) launched directly vs replayed from a CUDA graph: python tools/graph_ab.py [batch ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from wfl_asr_b200 import synth
from wfl_asr_b200.model import BIOPhonemeTagger
from wfl_asr_b200.pipeline import Labeler, _GraphedPass
dev = torch.device("cuda:0")
wl = os.environ.get("WFL_BENCH_WORKLOAD", "cfg2")
cfg = synth.workload_config(wl)
labels = synth.synth_labels(30)
model = synth.bench_model(BIOPhonemeTagger, cfg, labels).to(dev).eval()
lab = Labeler(model, median_filter=5, merge_mode="right", confidence_threshold=0.5)
for B in [int(a) for a in sys.argv[1:]] or [1, 32]:
    wave = torch.from_numpy(np.stack([synth.synth_wave(i, 30.0) for i in range(B)]).astype(np.float32)).to(dev)
    lang = torch.zeros(B, dtype=torch.long, device=dev)
    def timed(fn, n=10):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t = time.perf_counter(); e0.record()
        for _ in range(n): fn()
        e1.record(); cpu = (time.perf_counter() - t) / n * 1e3
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n, cpu
    d_gpu, d_cpu = timed(lambda: lab._pass(wave, lang))
    g = _GraphedPass(lab, wave, lang)
    g_gpu, g_cpu = timed(lambda: g.replay(lang))
    print(f"batch {B}: direct {d_gpu:.3f} ms (host issue {d_cpu:.3f} ms) | graph replay {g_gpu:.3f} ms (host issue {g_cpu:.3f} ms)")
