"""Multi-GPU sharding of the path: buckets are independent, so every device takes a
load-balanced slice of whole buckets and there is no collective on the data path.

  plan      umigpu_shard_plan (C ABI, host): LPT over per-bucket cost reads^2 + 64 reads
  engine    any callable (tid, pos, rev, umi, score) -> ascending kept indices *within the slice*
            (DeduplicateGPU.dedup_arrays on a GPU; the tests inject the CPU oracle to exercise this
            host logic under gloo)
  merge     survivors of all shards, mapped back to input indices, sorted ascending = canonical order
"""
from __future__ import annotations

import threading

import numpy as np

from .api import Cli, DeduplicateGPU, shard_plan


def take_shard(arrays: dict, shard_of_read: np.ndarray, shard: int):
    idx = np.nonzero(shard_of_read == shard)[0]
    return {k: (None if v is None else np.ascontiguousarray(v[idx])) for k, v in arrays.items()}, idx


def merge_kept(kept_global_lists) -> np.ndarray:
    parts = [np.asarray(k, dtype=np.int64) for k in kept_global_lists if len(k)]
    if not parts:
        return np.zeros(0, np.int64)
    return np.sort(np.concatenate(parts))


def dedup_sharded_inprocess(args: Cli, arrays: dict, devices: list[int]):
    """One process, one host thread + one umigpu context per device (the C ABI releases the GIL)."""
    plan, cost = shard_plan(arrays["tid"], arrays["pos"], arrays["rev"], len(devices))
    out = [None] * len(devices)

    def work(s, dev):
        sub, idx = take_shard(arrays, plan, s)
        kept, _, ctr = DeduplicateGPU(args, device=dev).dedup_arrays(sub["tid"], sub["pos"], sub["rev"], sub["umi"], sub["score"])
        out[s] = (idx[kept.astype(np.int64)], ctr)

    threads = [threading.Thread(target=work, args=(s, d)) for s, d in enumerate(devices)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    return merge_kept([o[0] for o in out]), [o[1] for o in out], cost


def dedup_distributed(arrays: dict, engine, dist, n_shards=None):
    """One process per device under torch.distributed (nccl on GPUs, gloo on CPU): every rank derives the
    same plan from the same keys, runs `engine` on its slice and rank 0 receives the merged result.
    Only the kept indices travel (gather_object); no collective touches the data path."""
    rank, world = dist.get_rank(), dist.get_world_size()
    plan, cost = shard_plan(arrays["tid"], arrays["pos"], arrays["rev"], n_shards or world)
    sub, idx = take_shard(arrays, plan, rank)
    kept_local = np.asarray(engine(sub["tid"], sub["pos"], sub["rev"], sub["umi"], sub["score"]), dtype=np.int64)
    kept_global = idx[kept_local]
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(kept_global, gathered, dst=0)
    return (merge_kept(gathered) if rank == 0 else None), cost
