"""CPU: multi-rank host logic of the bulk-labeling driver -- deterministic LPT sharding, length bucketing and the
end-of-run segment gather -- exercised with world_size 2 on the gloo backend."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from wfl_asr_b200 import shard
from wfl_asr_b200.pipeline import SEG_DTYPE


def _lengths(n=200, seed=4242):
    rng = np.random.default_rng(seed)
    return (rng.uniform(2.0, 30.0, n) * 16000).astype(np.int64).tolist()


def test_plan_is_balanced_complete_and_deterministic():
    lens = _lengths()
    for world in (1, 2, 4, 8):
        plan = shard.plan_shards(lens, world, "wavlm")
        assert sorted(i for s in plan for i in s) == list(range(len(lens)))
        assert plan == shard.plan_shards(lens, world, "wavlm")
        loads = [sum(shard.cost(shard.frames_for(lens[i], "wavlm")) for i in s) for s in plan]
        assert max(loads) <= 1.03 * (sum(loads) / world)
    # whisper: every clip costs the same (always 1500 frames) -> counts differ by at most one
    plan = shard.plan_shards(lens, 8, "whisper")
    assert max(map(len, plan)) - min(map(len, plan)) <= 1


def test_bucket_batches_respect_caps():
    lens = _lengths(300)
    idx = list(range(300))
    batches = shard.bucket_batches(idx, lens, max_clips=16, max_samples_per_batch=16 * 480000)
    assert sorted(i for _, g in batches for i in g) == idx
    for padded, group in batches:
        assert len(group) <= 16 and padded % 8000 == 0
        assert all(padded - 8000 < lens[i] <= padded for i in group)
    exact = shard.bucket_batches(idx, lens, 16, 16 * 480000, bucket_samples=1)
    assert all(all(lens[i] == p for i in g) for p, g in exact)


def test_plan_batches_is_world_size_independent_and_balanced():
    """Batch-level sharding: the SET of batches does not depend on the world size (so results cannot either), every
    utterance is in exactly one batch, and the LPT deal balances the ranks' estimated cost."""
    lens = _lengths(3000)
    one = shard.plan_batches(lens, 1, "wavlm", max_clips=16, max_samples_per_batch=16 * 480000)[0]
    key = lambda b: (b[0], tuple(b[1]))
    for world in (2, 4, 8):
        plan = shard.plan_batches(lens, world, "wavlm", max_clips=16, max_samples_per_batch=16 * 480000)
        assert plan == shard.plan_batches(lens, world, "wavlm", max_clips=16, max_samples_per_batch=16 * 480000)
        flat = [b for r in plan for b in r]
        assert sorted(map(key, flat)) == sorted(map(key, one))
        assert sorted(i for _, g in flat for i in g) == list(range(len(lens)))
        loads = [sum(shard.cost(shard.frames_for(p, "wavlm")) * len(g) for p, g in r) for r in plan]
        assert max(loads) <= 1.03 * (sum(loads) / world)
        # batches stay full: at most one partial batch per length bucket in the whole corpus, not one per rank
        partial = sum(1 for p, g in flat if len(g) < min(16, 16 * 480000 // p))
        assert partial <= len({p for p, _ in flat})
    for p, g in one:
        assert all(p - 8000 < lens[i] <= p for i in g)


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lens = _lengths(37)
    plan = shard.plan_shards(lens, world, "wavlm")
    local = []
    for i in plan[rank]:
        n = i % 5  # utterance i produced i % 5 segments (some have none)
        rec = np.zeros(n, dtype=SEG_DTYPE)
        rec["start"] = np.arange(n) + i
        rec["end"] = np.arange(n) + i + 0.5
        rec["ph"] = i
        local.append((i, rec))
    out = shard.gather_segments(local, torch.device("cpu"))
    if rank == 0:
        assert sorted(out) == list(range(37))
        for i, rec in out.items():
            assert len(rec) == i % 5
            assert np.array_equal(rec["start"], np.arange(i % 5) + i) and (rec["ph"] == i).all()
        open(os.path.join(tmp, "ok"), "w").write("ok")
    else:
        assert out is None
    dist.destroy_process_group()


def test_gather_segments_gloo_world2(tmp_path):
    port = 29600 + os.getpid() % 300
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").exists()


def test_gather_single_process():
    rec = np.zeros(3, dtype=SEG_DTYPE)
    rec["ph"] = [1, 2, 3]
    out = shard.gather_segments([(7, rec), (9, rec[:0])], torch.device("cpu"))
    assert list(out[7]["ph"]) == [1, 2, 3] and len(out[9]) == 0


def test_frames_for_every_encoder_type():
    """Frame counts the sharding cost model and the per-utterance decode lengths use: Whisper pads to 30 s, WavLM's
    conv chain (TF/models/wavlm: k 10,3,3,3,3,2,2 / s 5,2,2,2,2,2,2), the centred STFT of encoder_type "none"."""
    assert shard.frames_for(16000, "whisper") == 1500 and shard.frames_for(480000, "whisper") == 1500
    assert shard.frames_for(160000, "wavlm") == 499 and shard.frames_for(480000, "wavlm") == 1499
    assert shard.frames_for(399, "wavlm") == 0
    assert shard.frames_for(32000, "none") == 101 and shard.frames_for(480000, "none") == 1501
    assert shard.frames_for(32000, "none", hop=160) == 201
