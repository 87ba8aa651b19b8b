#!/usr/bin/env python
"""Headline benchmark: audio-seconds labeled per second (RTFx) on BASELINE.json configs[1]
(Whisper-base encoder + 4 Conformer blocks + median smoothing, batch 32 x 30 s, one B200 per rank).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU implementation (baseline/_ref,
                                                             # placed by oracle/place_reference.py; else the oracle port)

One "step" = one pass of the hot path over one batch of synthetic clips: log-mel -> encoder -> Conformer ->
heads -> threshold/argmax -> median -> BIO decode -> merge.  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = os.environ.get("WFL_BENCH_WORKLOAD", "cfg2")
METRIC = "audio_seconds_labeled_per_second"
UNIT = "audio-s/s"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def build_inputs(batch, seconds, rank, n_sets=2):
    from wfl_asr_b200 import synth
    sets = []
    for s in range(n_sets):
        w = np.stack([synth.synth_wave(rank * 100000 + s * batch + i, seconds) for i in range(batch)]).astype(np.float32)
        sets.append(torch.from_numpy(w))
    return sets


MEMORY_KINDS = ("layernorm", "split_f16", "decode_frames", "median_filter", "bio_decode", "merge_segments")
ALL_KINDS = MEMORY_KINDS + ("attention", "lstm")


def line_config(batch, world, wl, cfg, frames=1500):
    """The ``config`` object of the JSON line -- the SAME dict for this repo's arm and the reference arm."""
    pp, m = cfg["postprocess"], cfg["model"]
    arch_txt = (f"{m['whisper_model'] if m['encoder_type'] == 'whisper' else m['wavlm_model']} encoder"
                f"{' + BiLSTM(' + str(m.get('bilstm_num_layer', 1)) + ')' if m.get('enable_bilstm', True) else ''}"
                f" + {m['num_conformer_layers']} Conformer (heads {m['conformer_heads']}, ffx {m['conformer_ff_expansion']}, "
                f"k{m['conformer_kernel_size']}){' + dilated conv stack' if m.get('enable_dilated_conv', True) else ''}")
    return {"workload": f"BASELINE configs[{int(WORKLOAD[3:]) - 1}] ({WORKLOAD}): {arch_txt} + median {pp['median_filter']} + "
                        f"merge {pp['merge_segments']}, batch {batch} x {wl['seconds']:.0f} s per GPU, L=61, lang_id=0, random init",
            "batch_per_gpu": batch, "clip_seconds": wl["seconds"], "frames_per_clip": frames,
            "l2_policy": "inputs alternate between two batches; per-step working set (~1.5 GB of activations) exceeds the 126 MB L2",
            "parallelism": f"utterance-sharded x{world}, no hot-path collective"}


def tabulate(records):
    table = {}
    for kind, tag, work, a, b in records:
        d = table.setdefault((kind, tag), [0, 0.0, 0.0])
        d[0] += 1
        d[1] += a.elapsed_time(b)
        d[2] += work
    return table


def run_ours(args, rank, world, local_rank):
    from wfl_asr_b200 import ops, synth
    from wfl_asr_b200.model import BIOPhonemeTagger
    from wfl_asr_b200.pipeline import Labeler

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    wl = synth.WORKLOADS[WORKLOAD]
    batch = args.batch or wl["batch"]
    cfg = synth.workload_config(WORKLOAD)
    labels = synth.synth_labels(30)
    model = synth.bench_model(BIOPhonemeTagger, cfg, labels).to(dev).eval()
    pp = cfg["postprocess"]
    labeler = Labeler(model, median_filter=pp["median_filter"], merge_mode=pp["merge_segments"],
                      confidence_threshold=pp["confidence_threshold"])
    host_sets = [w.pin_memory() for w in build_inputs(batch, wl["seconds"], rank)]
    dev_sets = [w.to(dev) for w in host_sets]
    lang = torch.zeros(batch, dtype=torch.long, device=dev)
    audio_s_per_step = batch * wl["seconds"]

    def step_resident(i):
        return labeler.label_device(dev_sets[i % len(dev_sets)], lang)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step_resident(i)
    barrier()
    # ---- timed region 1: inputs resident in HBM
    sampler = ClockSampler(local_rank)
    sampler.start()
    # live in the timed region: CUDA events around the dominant kernel only (the 31-tap conv GEMM, 4 launches/step)
    ops.TIMING, ops.TIMING_MIN_SLABS, ops.TIMING_KINDS = [], 16, ()
    launches0 = ops.LAUNCHES
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step_resident(i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ops.LAUNCHES - launches0
    timing, ops.TIMING = ops.TIMING, None
    frames = int(step_resident(0)[2])
    # separate, untimed pass: events around every launch for the per-family tables (serialises launches slightly)
    ops.TIMING, ops.TIMING_MIN_SLABS, ops.TIMING_KINDS = [], 1, ALL_KINDS
    fam_steps = min(args.steps, 3)
    for i in range(fam_steps):
        step_resident(i)
    torch.cuda.synchronize()
    family, ops.TIMING, ops.TIMING_KINDS = ops.TIMING, None, ()
    # ---- timed region 2: end to end through the public call with HOST buffers (pinned H2D in, segments D2H out)
    # Public call: Labeler.label_stream(host batches) -> python segment lists; every step copies its batch host->device
    # from pinned memory and its segment records device->host (overlapped with the neighbouring steps' kernels).
    for out in labeler.label_stream((host_sets[i % len(host_sets)] for i in range(max(args.warmup, 2))), lang):
        pass
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    n_seg = 0
    for out in labeler.label_stream((host_sets[i % len(host_sets)] for i in range(args.steps)), lang):
        n_seg += sum(len(s) for s in out)
    t1.record()
    barrier()
    ms_e2e = t0.elapsed_time(t1)
    clocks = sampler.stop()
    # p50 latency of one 30 s clip, batch 1, host waveform in -> python segments out (synchronous call)
    one = host_sets[0][:1].clone().pin_memory()
    lang1 = lang[:1]
    lat = []
    for i in range(25):
        torch.cuda.synchronize()
        tt = time.perf_counter()
        labeler.label_host(one, lang1)
        lat.append((time.perf_counter() - tt) * 1e3)
    lat_p50 = statistics.median(lat[5:])
    # the same clip through the handle-level C ABI alone (csrc/handle.cu: wfl_forward + wfl_postprocess, its own CUDA
    # graph; no engine.py, no Python launch loop): pinned host waveform -> device -> segments on the host
    lat_native = None
    if cfg["model"]["encoder_type"] in ("whisper", "wavlm"):
        from wfl_asr_b200.native import QUERY_FRAMES, NativeModel
        nm = NativeModel(cfg, labels, {k: v.detach().cpu() for k, v in model.state_dict().items()}, dev, max_batch=1)
        side = torch.cuda.Stream(dev)
        wave1 = torch.empty(1, one.shape[1], device=dev)
        T1 = nm.query(QUERY_FRAMES, one.shape[1])
        out1 = (torch.empty(1, T1, nm.Lp, device=dev), torch.empty(1, T1, 2, device=dev))
        latn = []
        with torch.cuda.stream(side):
            for i in range(25):
                torch.cuda.synchronize()
                tt = time.perf_counter()
                wave1.copy_(one, non_blocking=True)
                nm.forward(wave1, lang1, out=out1)
                segs, nseg = nm.postprocess(out1[0], out1[1], median_filter=pp["median_filter"], merge_mode=pp["merge_segments"],
                                            confidence_threshold=pp["confidence_threshold"])
                n = int(nseg.cpu()[0])
                rec = segs[0, :max(n, 1)].cpu()
                latn.append((time.perf_counter() - tt) * 1e3)
        lat_native = statistics.median(latn[5:])
        nm.close()
    del dev_sets, host_sets, labeler, model
    torch.cuda.empty_cache()
    bulk = None if args.no_bulk else run_bulk(args, rank, world, dev)
    ingest = None
    if world == 1 and not args.no_ingest:
        # does the real-audio ingest keep up?  WAV files on tmpfs -> pinned staging -> H2D ->
        # device PCM decode + peak normalisation (ingest_only), and the same files through infer.infer_folder to .lab
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import ingest_bench
        ingest = ingest_bench.measure(files=args.ingest_files, seconds=wl["seconds"], workers=8, workload=WORKLOAD, dev=dev)
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms, ms_e2e = t.tolist()
    if rank != 0:
        return None

    peaks = measured_peaks()
    # dominant kernel: the Conformer conv-31 implicit GEMM (largest single launch); family totals reported too
    by_tag = tabulate(timing)
    fam_all = tabulate(family)
    fam_tag = {k: v for k, v in fam_all.items() if k[0] == "gemm"}
    if os.environ.get("WFL_BENCH_DEBUG"):
        for (kind, tag), (n, tms, work) in sorted(fam_all.items(), key=lambda kv: -kv[1][1]):
            rate = work / (tms * 1e-3) / 1e12 if tms > 0 else 0.0
            print(f"  {tms / fam_steps:7.3f} ms/step  x{n // fam_steps:3d}  {rate:9.2f} T(FLOP|B)/s  {tag}", file=sys.stderr)
    fam_ms = sum(v[1] for v in fam_tag.values()) * args.steps / fam_steps
    fam_flops = sum(v[2] for v in fam_tag.values()) * args.steps / fam_steps
    dom = max(by_tag.items(), key=lambda kv: kv[1][1]) if by_tag else None
    roofline = None
    if dom is not None:
        (kind, tag), (n, tms, fl) = dom
        achieved = fl / (tms * 1e-3) / 1e12
        # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture of this
        # kernel on this shape (profiles/traffic.json names the capture), else null
        traffic = ncu_traffic(tag)
        roofline = {"bound": "tensor", "kernel": f"gemm_kernel[{tag}]", "achieved": round(achieved, 1),
                    "peak": peaks["tf_sustained"], "peak_source": peaks["src"] + " (sustained bf16 dense)",
                    "peak_burst": peaks["tf_burst"], "frac_of_burst": round(achieved / peaks["tf_burst"], 4),
                    "unit": "TFLOP/s", "frac": round(achieved / peaks["tf_sustained"], 4),
                    "traffic": traffic["bytes"] if traffic else None, "traffic_unit": "bytes/launch",
                    "traffic_source": traffic["source"] if traffic else None,
                    "algorithmic_bytes_per_launch": gemm_algorithmic_bytes(tag),
                    "launches": n, "avg_launch_ms": round(tms / n, 4),
                    "algorithmic_flop_per_launch": fl / n,
                    "gemm_family": {"note": "all wfl_gemm launches, measured in a separate event-instrumented pass",
                                    "launches_per_step": sum(v[0] for v in fam_tag.values()) // fam_steps,
                                    "ms_per_step": round(fam_ms / args.steps, 3),
                                    "achieved": round(fam_flops / (fam_ms * 1e-3) / 1e12, 1),
                                    "share_of_step": round(fam_ms / ms, 3)}}
        att = [(k[1], v) for k, v in fam_all.items() if k[0] == "attention"]
        roofline["attention"] = [{"kernel": t, "launches_per_step": v[0] // fam_steps, "avg_launch_ms": round(v[1] / v[0], 4),
                                  "achieved": round(v[2] / (v[1] * 1e-3) / 1e12, 1), "unit": "TFLOP/s",
                                  "frac_of_burst": round(v[2] / (v[1] * 1e-3) / 1e12 / peaks["tf_burst"], 4)} for t, v in att]
        # memory-bound kernels: achieved GB/s on ALGORITHMIC bytes against the measured HBM copy rate
        mem = []
        for (k, t), v in sorted(fam_all.items(), key=lambda kv: -kv[1][1]):
            if k not in MEMORY_KINDS:
                continue
            us = v[1] / v[0] * 1e3
            gbs = v[2] / v[0] / (us * 1e-6) / 1e9 if v[2] > 0 else None
            mem.append({"kernel": t, "launches_per_step": v[0] // fam_steps, "avg_launch_us": round(us, 2),
                        "algorithmic_bytes_per_launch": int(v[2] / v[0]) if v[2] > 0 else None,
                        "achieved_GBps": round(gbs, 1) if gbs else None,
                        "frac_of_hbm_peak": round(gbs / peaks["hbm"], 4) if gbs else None})
        roofline["memory_bound"] = {"peak_GBps": peaks["hbm"], "peak_source": peaks["src"] + " (HBM copy)", "kernels": mem}
        lstm = [(k[1], v) for k, v in fam_all.items() if k[0] == "lstm"]
        if lstm:
            roofline["lstm"] = [{"kernel": t, "launches_per_step": v[0] // fam_steps,
                                 "us_per_serial_step": round(v[1] * 1e3 / v[2], 3)} for t, v in lstm]
    value = audio_s_per_step * world * args.steps / (ms * 1e-3)
    e2e_value = audio_s_per_step * world * args.steps / (ms_e2e * 1e-3)
    h2d = batch * int(round(wl["seconds"] * 16000)) * 4
    d2h = batch * frames * 24 + batch * 4
    line = {
        "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "fp16",
        "dtype_note": "IEEE fp16 operands (same tcgen05 kind::f16 rate and width as bf16, 11-bit significand), fp32 accumulate, "
                      "fp32 residual stream / normalisation / softmax / LSTM cell",
        "data": "synthetic",
        "config": line_config(batch, world, wl, cfg, frames),
        "e2e": {"value": round(e2e_value, 1), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": round(ms_e2e / args.steps, 3), "segments_per_step": n_seg / args.steps},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
        "latency_p50_ms": {"value": round(lat_p50, 3), "what": f"one {wl['seconds']:g} s clip, batch 1, pinned host waveform -> python "
                           "segments (H2D + forward + post-processing + D2H), median of 20 synchronous calls",
                           "handle_c_abi": None if lat_native is None else round(lat_native, 3),
                           "handle_c_abi_what": "same clip through wfl_forward + wfl_postprocess of the handle-level C ABI only"},
        "bulk": bulk, "ingest": ingest,
    }
    return line


def run_bulk(args, rank, world, dev):
    """BASELINE configs[3] through the real multi-GPU path: a FIXED ragged corpus (10k utterances, lengths U[2,30] s,
    seed 4242 -- strong scaling: the same corpus for every N) is length-bucketed, dealt to the ranks batch by batch
    (shard.plan_batches), staged through pinned memory, labeled by WavLM-large + 6 Conformer, and the segment records
    are gathered to rank 0 over NCCL -- all inside the timed region (CUDA events, max over ranks)."""
    from wfl_asr_b200 import bulk, synth
    from wfl_asr_b200.model import BIOPhonemeTagger
    name = args.bulk_workload
    cfg = synth.workload_config(name)
    labels = synth.synth_labels(30)
    model = synth.bench_model(BIOPhonemeTagger, cfg, labels).to(dev).eval()
    corpus = synth.RaggedCorpus(args.bulk_utts)
    pp = cfg["postprocess"]
    kw = dict(median_filter=pp["median_filter"], merge_mode=pp["merge_segments"], confidence_threshold=pp["confidence_threshold"],
              max_clips=args.bulk_max_clips, max_samples_per_batch=args.bulk_max_clips * 480000 // 2, lengths=None)
    warm = synth.RaggedCorpus(8 * world, seed=7)
    bulk.label_corpus(model, warm, [0] * len(warm), **dict(kw, lengths=warm.lengths))  # kernel attributes, NCCL communicator
    _ = corpus[0]  # builds the base-clip pool outside the timed region (synthesis is not ingest)
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_host = time.perf_counter()
    e0.record()
    segs = bulk.label_corpus(model, corpus, [0] * len(corpus), **dict(kw, lengths=corpus.lengths))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    host_s = time.perf_counter() - t_host
    if world > 1:
        t = torch.tensor([ms, host_s], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms, host_s = t.tolist()
    if rank != 0:
        return None
    assert len(segs) == len(corpus)
    m = cfg["model"]
    return {"metric": METRIC, "value": round(corpus.audio_seconds / (ms * 1e-3), 1), "unit": UNIT, "scaling": "strong",
            "workload": f"BASELINE configs[3] ({name}): {m['wavlm_model'] if m['encoder_type'] == 'wavlm' else m['whisper_model']} + "
                        f"{m['num_conformer_layers']} Conformer, {len(corpus)} utterances U[2,30] s (seed 4242), 0.5 s length "
                        f"buckets, batches of <= {args.bulk_max_clips} dealt to {world} rank(s), host staging + H2D + "
                        f"forward + post-processing + D2H + NCCL gather of the segment records in the timed region",
            "utterances": len(corpus), "audio_seconds": round(corpus.audio_seconds, 1), "n_gpus": world,
            "ms": round(ms, 1), "host_wall_ms": round(host_s * 1e3, 1), "segments": sum(len(s) for s in segs),
            "gather_in_timed_region": True}


def ncu_traffic(tag):
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        return json.load(f).get(tag)


def gemm_algorithmic_bytes(tag):
    """A (f16) + W (f16) read once, output written once (f16; fp32 read-modify-write for the residual-add mode)."""
    try:
        dims, slabs, mode = tag.split("/")[:3]
        M, N, K = (int(v[1:]) for v in dims.split("x"))
        taps = int(slabs[5:])
        mode = int(mode[4:])
    except (ValueError, IndexError):
        return None
    out = {0: 2, 1: 4, 2: 8, 3: 1}.get(mode, 2) * M * N
    return 2 * M * (K // taps) + 2 * N * K + out  # a conv's taps re-read the same activation rows (L2 / smem hits)


def _reference_objects(cfg, labels, sd):
    """(kind, forward(wave, lang) -> (logits, offsets), postprocess(logits[b], offsets[b]) -> segments).
    kind "reference": the UNMODIFIED REF/model.py + REF/infer.py + REF/utils.py placed under baseline/_ref by
    oracle/place_reference.py (strict load_state_dict of the same weights); kind "port": oracle/ restatement."""
    from oracle import place_reference, ref_loader
    pp = cfg["postprocess"]
    if place_reference.placed() and not os.environ.get("WFL_BENCH_FORCE_PORT"):
        ref_loader.REF_DIR = place_reference.DEST
        ref = ref_loader.build_reference_model(cfg, labels, randomize_bn=False)
        ref.load_state_dict(sd, strict=True)
        _, ref_utils, ref_infer = ref_loader.load_reference_modules()
        from scipy.ndimage import median_filter

        def forward(wave, lang):
            with torch.no_grad():
                return ref(wave, lang)

        def post(logits, offsets):
            # REF/infer.py:288-310 (single-chunk branch) with the reference's own functions
            tags = ref_infer.suppress_low_confidence(logits, ref.id2label, threshold=pp["confidence_threshold"])
            ids = [ref.label2id.get(t, ref.label2id["O"]) for t in tags]
            if pp["median_filter"] > 1:
                ids = median_filter(ids, size=pp["median_filter"])
            tags = [ref.id2label[int(i)] for i in ids]
            segs = ref_utils.decode_bio_tags(tags, frame_duration=0.02, offsets=offsets)
            if pp["merge_segments"] != "none":
                segs = ref_utils.merge_adjacent_segments(segs, mode=pp["merge_segments"])
            return segs
        return "reference", forward, post
    from oracle import postproc_oracle as po
    from oracle import torch_oracle as to

    def forward(wave, lang):
        return to.forward(wave, sd, cfg, lang)

    def post(logits, offsets):
        ids, segs = po.postprocess_clip(logits.numpy(), offsets.numpy(), labels, pp["confidence_threshold"],
                                        pp["median_filter"], pp["merge_segments"])
        return po.merge_adjacent_segments(segs, pp["merge_segments"])
    return "port", forward, post


def cpu_reference_run(steps, warmup, sample_clips, seconds):
    """The reference's CPU implementation of the path on the host cores (all threads): batched fp32 forward under
    no_grad + per-clip python post-processing, same weights and synthetic clips as the CUDA arm."""
    from wfl_asr_b200 import synth
    from wfl_asr_b200.model import BIOPhonemeTagger
    cfg = synth.workload_config(WORKLOAD)
    labels = synth.synth_labels(30)
    model = synth.bench_model(BIOPhonemeTagger, cfg, labels)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    del model
    torch.set_num_threads(os.cpu_count() or 1)
    kind, forward, post = _reference_objects(cfg, labels, sd)
    wave = torch.from_numpy(np.stack([synth.synth_wave(i, seconds) for i in range(sample_clips)]).astype(np.float32))
    lang = torch.zeros(sample_clips, dtype=torch.long)

    def step():
        logits, offsets = forward(wave, lang)
        return sum(len(post(logits[b], offsets[b])) for b in range(sample_clips))

    for _ in range(warmup):
        step()
    t = time.perf_counter()
    for _ in range(steps):
        n_seg = step()
    dt = time.perf_counter() - t
    return sample_clips * seconds * steps / dt, dt / steps, torch.get_num_threads(), kind, n_seg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--cpu-sample-clips", type=int, default=0, help="clips per CPU step (0 = the workload's batch for "
                    "--impl reference, 8 for the cpu_baseline leg of the default run)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-bulk", action="store_true", help="skip the ragged-corpus (BASELINE configs[3]) record")
    ap.add_argument("--no-ingest", action="store_true", help="skip the file-ingest record (N = 1 only)")
    ap.add_argument("--ingest-files", type=int, default=128)
    ap.add_argument("--bulk-workload", default="cfg4")
    ap.add_argument("--bulk-utts", type=int, default=10000)
    ap.add_argument("--bulk-max-clips", type=int, default=32)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from wfl_asr_b200 import synth
    wl = synth.WORKLOADS[WORKLOAD]
    batch = args.batch or wl["batch"]

    if args.impl == "reference":
        if rank != 0:
            return 0
        clips = args.cpu_sample_clips or batch
        v, s_per_step, cores, kind, n_seg = cpu_reference_run(max(args.steps, 1), max(args.warmup, 1), clips, wl["seconds"])
        what = ("unmodified REF/model.py forward + REF/infer.py / REF/utils.py post-processing (baseline/_ref)" if kind == "reference"
                else "oracle/ restatement of the reference (baseline/_ref not placed)")
        sample = (f"{clips} clips x {wl['seconds']:.0f} s per step ({'the whole batch of the workload' if clips == batch else 'bounded sample of the batch-' + str(batch) + ' workload'}), "
                  f"fp32, torch.no_grad, {cores} host threads; {what}")
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": round(v, 2), "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(s_per_step * 1e3, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": line_config(batch, args.gpus, wl, synth.workload_config(WORKLOAD)),
            "cpu_baseline": {"value": round(v, 2), "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": round(v, 2), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "segments_per_step": n_seg}), flush=True)
        return 0

    if world > 1:
        # stdout carries exactly one JSON line: NCCL's own log (the driver may set NCCL_DEBUG=INFO to check the
        # communicator's rank count) goes to stderr instead of being silenced
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    line = run_ours(args, rank, world, local_rank)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            clips = args.cpu_sample_clips or 8
            v, s_per_step, cores, kind, _ = cpu_reference_run(1, 1, clips, wl["seconds"])
            line["cpu_baseline"] = {"value": round(v, 2), "unit": UNIT, "cores": cores, "kind": kind,
                                    "sample": f"{clips} clips x {wl['seconds']:.0f} s, 1 warm-up + 1 timed pass "
                                              f"({s_per_step:.1f} s), fp32 no_grad forward + python post-processing"}
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
