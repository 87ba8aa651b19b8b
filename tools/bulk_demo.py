"""BASELINE configs[3]/[4] flow on N GPUs: synthetic utterances of 2-30 s, length-bucketed, sharded over the ranks with
bulk.label_corpus (one process per GPU, NCCL only for the final gather of segment records), .lab files written by rank 0.

  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bulk_demo.py --workload cfg4 --utts 256
"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from wfl_asr_b200 import bulk, synth, utils
from wfl_asr_b200.model import BIOPhonemeTagger

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg4")
ap.add_argument("--utts", type=int, default=256)
ap.add_argument("--out", default="")
args = ap.parse_args()
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
cfg = synth.workload_config(args.workload)
labels = synth.synth_labels(30)
model = synth.bench_model(BIOPhonemeTagger, cfg, labels).to(dev).eval()
rng = np.random.default_rng(4242)  # lengths U[2, 30] s, seed 4242
secs = rng.uniform(2.0, 30.0, size=args.utts)
base = [synth.synth_wave(i, 30.0).astype(np.float32) for i in range(8)]  # 8 distinct clips, cropped to each length
waves = [base[i % 8][:int(s * 16000)] for i, s in enumerate(secs)]
pp = cfg["postprocess"]
kw = dict(median_filter=pp["median_filter"], merge_mode=pp["merge_segments"], confidence_threshold=pp["confidence_threshold"])
bulk.label_corpus(model, waves[:world * 4], [0] * (world * 4), **kw)  # warm-up (kernel attributes, workspaces)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t = time.perf_counter()
segs = bulk.label_corpus(model, waves, [0] * len(waves), **kw)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
dt = time.perf_counter() - t
if rank == 0:
    assert len(segs) == len(waves) and all(isinstance(s, list) for s in segs)
    n_seg = sum(len(s) for s in segs)
    # (Whisper pads every clip to 30 s and, like the reference, decodes all 1500 frames; WavLM clips end with the audio)
    ok = all(all(0.0 <= a <= b <= (30.0 if model.encoder_type == 'whisper' else len(w) / 16000 + 0.54) for a, b, _ in s)
             for s, w in zip(segs, waves))
    if args.out:
        os.makedirs(args.out, exist_ok=True)
        for i, s in enumerate(segs):
            utils.save_lab(os.path.join(args.out, f"utt{i:05d}.lab"), s)
    print(f"{args.workload}: {len(waves)} utterances, {secs.sum():.0f} audio-s on {world} GPU(s): {dt:.2f} s wall -> "
          f"{secs.sum() / dt:.0f} audio-s/s; {n_seg} segments; times inside clips: {ok}")
if world > 1:
    dist.destroy_process_group()
