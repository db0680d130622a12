"""N>1 host logic under gloo, world_size 2, on CPU: shard plan -> per-rank slice -> merge in canonical order.
The per-shard engine here is the CPU oracle (the checker standing in for a device); on a GPU box the same
function runs with DeduplicateGPU (test_gpu_parity.py::test_sharded_matches_unsharded)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as O
from umigpu import shard, synth


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d, cfg = synth.generate_config("C2", device="cpu", scale=0.001)
    h = {k: v.numpy() for k, v in d.items()}

    def engine(tid, pos, rev, umi, score):
        kept, _, _ = O.dedup(tid, pos, rev, umi, score, O.ALGO_DIR, O.MERGE_AVGQUAL, 1, 0.5)
        return kept

    merged, cost = shard.dedup_distributed(h, engine, dist)
    if rank == 0:
        full, _, _ = O.dedup(h["tid"], h["pos"], h["rev"], h["umi"], h["score"], O.ALGO_DIR, O.MERGE_AVGQUAL, 1, 0.5)
        q.put((merged.tolist() == full.tolist(), [int(c) for c in cost]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gloo_sharded_equals_unsharded():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
        assert p.exitcode == 0
    ok, cost = q.get()
    assert ok
    assert len(cost) == 2 and min(cost) > 0


def test_merge_and_take_shard_roundtrip():
    rng = np.random.default_rng(3)
    n = 5000
    arrays = dict(tid=rng.integers(0, 3, n).astype(np.int32), pos=rng.integers(0, 50, n).astype(np.int64),
                  rev=rng.integers(0, 2, n).astype(np.uint8), umi=rng.integers(65, 70, (n, 6)).astype(np.uint8), score=None)
    from umigpu import shard_plan
    plan, _ = shard_plan(arrays["tid"], arrays["pos"], arrays["rev"], 3)
    seen = []
    for s in range(3):
        sub, idx = shard.take_shard(arrays, plan, s)
        assert sub["score"] is None and len(sub["tid"]) == len(idx)
        assert (arrays["pos"][idx] == sub["pos"]).all()
        seen.append(idx)
    assert shard.merge_kept(seen).tolist() == list(range(n))
    assert shard.merge_kept([[], []]).tolist() == []


# ---- contiguous slices of a coordinate-sorted stream (umigpu_shard_plan_sorted: pure host code) ----
def _plan_case(scale=0.004, seed=None, **kw):
    d, cfg = synth.generate_config("C2", device="cpu", scale=scale, seed=seed, **kw)
    return {k: v.numpy() for k, v in d.items()}


def test_plan_sorted_cuts_are_bucket_boundaries_and_balanced():
    import umigpu
    h = _plan_case()
    n = len(h["tid"])
    key = h["tid"].astype(np.int64) * (1 << 36) + h["pos"]
    assert (np.diff(key) >= 0).all()
    lib = umigpu.load()
    assert lib.umigpu_pos_key(3, -7) == 3 * (1 << 36) - 7
    for N in (1, 2, 3, 4, 8, 16):
        cuts, keys, hot, cost = umigpu.shard_plan_sorted(h["tid"], h["pos"], h["rev"], N, hot_min_reads=2000)
        cuts = cuts.astype(np.int64)
        assert cuts[0] == 0 and cuts[-1] == n and (np.diff(cuts) >= 0).all()
        for s in range(1, N):
            c = int(cuts[s])
            if 0 < c < n:
                assert key[c - 1] < key[c], "a cut must fall on the first read of a (contig, position)"
                assert int(keys[s]) == int(key[c])
        assert int(keys[0]) == -2 ** 63 and int(keys[-1]) == 2 ** 63 - 1
        # the hot bucket: the most frequent (position, strand), found from <= 65536 probes
        assert hot.present == 1
        vals, counts = np.unique(key * 2 + h["rev"], return_counts=True)
        top = int(vals[np.argmax(counts)])
        r = int(hot.read_index)
        assert int(key[r]) * 2 + int(h["rev"][r]) == top
        assert cuts[hot.owner] <= r < cuts[hot.owner + 1]
        assert abs(int(hot.reads_est) - int(counts.max())) <= 0.1 * counts.max() + 8
        if N > 1:       # balanced up to the one bucket that cannot be cut (it may outweigh the fair share)
            hot_cost = 0.27 * counts.max()
            assert cost.max() <= max(1.6 * cost.mean(), hot_cost + 1.3 * cost.mean()), (N, cost)
            assert (cost > 0).sum() >= min(N, 8) - 1, (N, cost)
    # no bucket big enough: nothing is split
    _, _, hot, _ = umigpu.shard_plan_sorted(h["tid"], h["pos"], h["rev"], 4, hot_min_reads=10 ** 9)
    assert hot.present == 0
    _, _, hot, _ = umigpu.shard_plan_sorted(h["tid"], h["pos"], h["rev"], 4, hot_min_reads=2 ** 64 - 1)
    assert hot.present == 0
    # empty input and more shards than positions
    cuts, keys, hot, _ = umigpu.shard_plan_sorted(np.zeros(0, np.int32), np.zeros(0, np.int64), np.zeros(0, np.uint8), 3)
    assert cuts.tolist() == [0, 0, 0, 0] and hot.present == 0
    cuts, _, _, _ = umigpu.shard_plan_sorted(np.zeros(50, np.int32), np.zeros(50, np.int64), np.zeros(50, np.uint8), 4)
    assert cuts[0] == 0 and cuts[-1] == 50 and set(cuts.tolist()) <= {0, 50}


def _slice_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import umigpu
    h = _plan_case(0.001, seed=8, n_contigs=3)
    cuts, keys, hot, _ = umigpu.shard_plan_sorted(h["tid"], h["pos"], h["rev"], world, hot_min_reads=500)
    a, b = int(cuts[rank]), int(cuts[rank + 1])
    kept, _, _ = O.dedup(h["tid"][a:b], h["pos"][a:b], h["rev"][a:b], h["umi"][a:b], h["score"][a:b], O.ALGO_DIR, O.MERGE_AVGQUAL, 1, 0.5)
    parts = [None] * world
    dist.all_gather_object(parts, (kept + a).tolist())
    if rank == 0:
        full, _, _ = O.dedup(h["tid"], h["pos"], h["rev"], h["umi"], h["score"], O.ALGO_DIR, O.MERGE_AVGQUAL, 1, 0.5)
        q.put(sum(parts, []) == full.tolist())          # contiguous slices: concatenation in rank order IS the merge
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gloo_contiguous_slices_concatenate():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_slice_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
        assert p.exitcode == 0
    assert q.get()
