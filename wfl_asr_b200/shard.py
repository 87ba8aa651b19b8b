"""Utterance sharding for multi-GPU bulk labeling (one process per GPU, no hot-path collective).

Utterances (and the 30 s chunks of long files, REF/infer.py:19-28) are independent, so the path shards by
data: a longest-processing-time assignment of utterances to ranks on an estimated cost ``a*T + b*T^2``
(T = encoder frames: the quadratic term is attention), then length-bucketed batches inside each rank.
The only communication is the end-of-run gather of variable-length segment records to rank 0
(``all_gather`` of counts + ``all_gather`` of padded record tensors; NCCL on GPUs, gloo in the CPU tests).
"""
import heapq

import numpy as np
import torch
import torch.distributed as dist

from .pipeline import SEG_DTYPE


def frames_for(n_samples, encoder_type, hop=320):
    if encoder_type == "whisper":
        return 1500  # Whisper pads/truncates every clip to 30 s
    if encoder_type in ("none", "null"):
        return 1 + n_samples // hop  # MelSpectrogram with center=True (REF/model.py:85-90)
    n = n_samples
    for k, s in zip((10, 3, 3, 3, 3, 2, 2), (5, 2, 2, 2, 2, 2, 2)):
        n = (n - k) // s + 1
    return max(n, 0)


def cost(frames, a=1.0, b=1.0 / 1500.0):
    """Relative cost of one utterance: linear (GEMM/conv) + quadratic (attention) in the frame count."""
    return a * frames + b * frames * frames


def plan_shards(n_samples_per_utt, world_size, encoder_type="whisper"):
    """Deterministic LPT assignment -> list (per rank) of utterance indices, each sorted by length (desc).
    Every rank computes the same plan from the same lengths, so no communication is needed to agree on it."""
    costs = [cost(frames_for(n, encoder_type)) for n in n_samples_per_utt]
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    heap = [(0.0, r) for r in range(world_size)]
    heapq.heapify(heap)
    shards = [[] for _ in range(world_size)]
    for i in order:
        load, r = heapq.heappop(heap)
        shards[r].append(i)
        heapq.heappush(heap, (load + costs[i], r))
    return shards


def bucket_batches(indices, n_samples_per_utt, max_clips, max_samples_per_batch, bucket_samples=8000):
    """Groups a rank's utterances into batches of similar length.  Lengths are rounded UP to a multiple of
    ``bucket_samples`` (0.5 s at 16 kHz) and clips are zero-padded to the bucket length -- the batched semantics of
    the reference's own batched caller (REF/train.py:22-36 collate zero-pads, no masks).  ``bucket_samples=1`` gives
    exact-length groups (bit-identical to per-file REF/infer.py for WavLM, at the price of tiny batches)."""
    by_len = {}
    for i in indices:
        n = n_samples_per_utt[i]
        padded = -(-n // bucket_samples) * bucket_samples
        by_len.setdefault(padded, []).append(i)
    batches = []
    for padded in sorted(by_len, reverse=True):
        group = by_len[padded]
        cap = max(1, min(max_clips, max_samples_per_batch // padded))
        for s in range(0, len(group), cap):
            batches.append((padded, group[s:s + cap]))
    return batches


def plan_batches(n_samples_per_utt, world_size, encoder_type="whisper", max_clips=32, max_samples_per_batch=32 * 480000,
                 bucket_samples=8000):
    """Shards BATCHES instead of utterances: the corpus is length-bucketed into batches first (independent of the
    world size), then whole batches are dealt to ranks longest-processing-time first on the summed utterance cost.
    Dealing single utterances (``plan_shards``) spreads every length bucket over all ranks, so at 8 ranks each rank
    is left with 1/8 of every bucket and runs small, badly filled batches; dealing batches keeps them full on every
    rank, and -- because a clip's padded length and the kernels' arithmetic do not depend on its batch mates -- the
    results are the same for every world size.  Returns a list (per rank) of (padded_length, [utterance indices])."""
    batches = bucket_batches(range(len(n_samples_per_utt)), n_samples_per_utt, max_clips, max_samples_per_batch,
                             bucket_samples)
    hop = 320
    costs = [sum(cost(frames_for(padded, encoder_type, hop)) for _ in group) for padded, group in batches]
    order = sorted(range(len(batches)), key=lambda k: (-costs[k], k))
    heap = [(0.0, r) for r in range(world_size)]
    heapq.heapify(heap)
    per_rank = [[] for _ in range(world_size)]
    for k in order:
        load, r = heapq.heappop(heap)
        per_rank[r].append(k)
        heapq.heappush(heap, (load + costs[k], r))
    # inside a rank: longest first, so the largest workspaces are sized once and the tail is short batches
    return [[batches[k] for k in sorted(ks, key=lambda k: (-batches[k][0], k))] for ks in per_rank]


def gather_segments(local, device):
    """``local``: list of (utterance index, numpy structured array of SEG_DTYPE records) on this rank.
    Returns on rank 0 a dict {utterance index: records}; None elsewhere.  Payload is KBs: latency, not bandwidth."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    idx = np.asarray([i for i, _ in local], dtype=np.int64)
    cnt = np.asarray([len(r) for _, r in local], dtype=np.int64)
    recs = np.concatenate([r for _, r in local]) if len(local) and cnt.sum() else np.zeros(0, dtype=SEG_DTYPE)
    if world == 1:
        out, pos = {}, 0
        for i, c in zip(idx, cnt):
            out[int(i)] = recs[pos:pos + c]
            pos += c
        return out
    sizes = torch.tensor([len(idx), len(recs)], dtype=torch.int64, device=device)
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes)
    max_utt = int(max(s[0].item() for s in all_sizes))
    max_rec = int(max(s[1].item() for s in all_sizes))
    meta = torch.zeros(max(max_utt, 1), 2, dtype=torch.int64, device=device)
    if len(idx):
        meta[:len(idx), 0] = torch.from_numpy(idx).to(device)
        meta[:len(idx), 1] = torch.from_numpy(cnt).to(device)
    payload = torch.zeros(max(max_rec, 1), SEG_DTYPE.itemsize, dtype=torch.uint8, device=device)
    if len(recs):
        payload[:len(recs)] = torch.from_numpy(recs.view(np.uint8).reshape(-1, SEG_DTYPE.itemsize)).to(device)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    payloads = [torch.zeros_like(payload) for _ in range(world)]
    dist.all_gather(metas, meta)
    dist.all_gather(payloads, payload)
    if rank != 0:
        return None
    out = {}
    for r in range(world):
        n_utt = int(all_sizes[r][0].item())
        m = metas[r][:n_utt].cpu().numpy()
        raw = payloads[r].cpu().numpy().reshape(-1).view(SEG_DTYPE)
        pos = 0
        for i, c in m:
            out[int(i)] = raw[pos:pos + int(c)].copy()
            pos += int(c)
    return out
