"""Bulk labeling driver: many utterances -> segments, sharded over the GPUs of one box (one process per GPU).

This is the batched, multi-GPU form of REF/infer.py's per-file loop (REF/infer.py:330-357): utterances are
assigned to ranks by ``shard.plan_shards``, grouped into length buckets, pushed through the forward kernels and
the device post-processing, and the per-rank results are gathered to rank 0 at the end (the only collective).
"""
import numpy as np
import torch
import torch.distributed as dist

from . import shard
from .pipeline import SEG_DTYPE, Labeler


def label_corpus(model, waves, lang_ids=None, *, median_filter=1, merge_mode="right", confidence_threshold=0.0,
                 max_clips=32, max_samples_per_batch=32 * 480000, bucket_samples=8000):
    """waves: list of 1-D float32 numpy arrays (16 kHz, already peak-normalised, each <= 30 s).
    Returns on rank 0 a list (per utterance) of [(start, end, phoneme)]; None on other ranks."""
    dev = next(model.parameters()).device
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    lens = [int(len(w)) for w in waves]
    if any(n > 480000 for n in lens):
        raise ValueError("label_corpus takes clips of at most 30 s; split longer files first (infer.split_audio)")
    etype = model.encoder_type
    plan = shard.plan_shards(lens, world, etype)
    labeler = Labeler(model, median_filter=median_filter, merge_mode=merge_mode,
                      confidence_threshold=confidence_threshold)
    bucket = 480000 if etype == "whisper" else bucket_samples  # Whisper pads every clip to 30 s itself
    local = []
    pending = None  # (group, T, pinned records, pinned counts, event) of the batch whose results are still in flight

    def collect(item):
        group, T, rec, cnt, done = item
        done.synchronize()
        raw = rec.numpy().reshape(-1).view(SEG_DTYPE)
        counts = cnt.numpy()
        for j, i in enumerate(group):
            local.append((i, raw[j * T:j * T + int(counts[j])].copy()))

    # two pinned staging buffers, sized for the largest batch once (cudaHostAlloc per batch costs more than the batch's
    # kernels); buffer k % 2 is free again when batch k - 2 has been collected
    batches = shard.bucket_batches(plan[rank], lens, max_clips, max_samples_per_batch, bucket)
    cap = max((len(g) * p for p, g in batches), default=0)
    staging = [torch.empty(cap, dtype=torch.float32).pin_memory() for _ in range(2)] if cap else []
    for k, (padded, group) in enumerate(batches):
        host = staging[k & 1][:len(group) * padded].view(len(group), padded)
        host.zero_()
        for j, i in enumerate(group):
            host[j, :lens[i]] = torch.from_numpy(np.asarray(waves[i], dtype=np.float32))
        wave = host.to(dev, non_blocking=True)
        lang = None
        if lang_ids is not None:
            lang = torch.tensor([lang_ids[i] for i in group], dtype=torch.long).pin_memory().to(dev, non_blocking=True)
        logits, offsets = model(wave, lang)
        T = logits.shape[1]
        # like the reference's batched caller (REF/train.py:485-495): decode each item on its own frame count
        valid = torch.tensor([min(T, shard.frames_for(lens[i], etype, model.arch.get("hop", 320))) for i in group],
                             dtype=torch.int32).pin_memory().to(dev, non_blocking=True)
        _, merged, nout, fcb, n_files = labeler.postprocess(logits, offsets, valid)
        # results leave through pinned buffers; the host decodes batch k-1 while batch k runs on the GPU
        rec = torch.empty(merged.shape, dtype=torch.uint8).pin_memory()
        cnt = torch.empty(n_files, dtype=torch.int32).pin_memory()
        rec.copy_(merged, non_blocking=True)
        cnt.copy_(nout[:n_files], non_blocking=True)
        done = torch.cuda.Event()
        done.record()
        if pending is not None:
            collect(pending)
        pending = (group, T, rec, cnt, done)
    if pending is not None:
        collect(pending)
    gathered = shard.gather_segments(local, dev)
    if gathered is None:
        return None
    names = labeler.out_names
    out = []
    for i in range(len(waves)):  # .tolist() converts a whole column at once (69 k segments: 70 ms -> 15 ms)
        rec = gathered[i]
        out.append(list(zip(rec["start"].tolist(), rec["end"].tolist(), [names[p] for p in rec["ph"].tolist()])))
    return out
