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
