"""One dataset over several devices: contiguous slices of the coordinate-sorted stream, the hot bucket's
neighbour search split over the group through exchange windows.  Every case is compared with the CPU oracle and with the
unsharded CUDA run.  On a one-GPU box the group is several contexts on device 0 (the exchange then runs over the same
code with device-local copies); with more GPUs the real devices are used as well."""
import os
import socket

import numpy as np
import pytest

import oracle_lib as O
import umigpu
from umigpu import synth

pytestmark = pytest.mark.gpu

CTR_SUM = ("total_reads", "n_buckets", "total_umis", "n_kept", "unordered_pairs")


def small(name, scale, seed=None, **kw):
    d, cfg = synth.generate_config(name, seed=seed, device="cpu", scale=scale, **kw)
    return {k: v.numpy() for k, v in d.items()}, cfg


def device_lists():
    import torch
    nd = torch.cuda.device_count()
    out = [[0], [0, 0], [0, 0, 0], [0] * 8]
    if nd > 1:
        out.append(list(range(nd)))
    return out


def oracle(d, algo=O.ALGO_DIR, k=1):
    return O.dedup(d["tid"], d["pos"], d["rev"], d["umi"], d["score"], algo, O.MERGE_AVGQUAL, k, 0.5)


@pytest.fixture
def hot_env(monkeypatch):
    monkeypatch.setenv("UMIGPU_HOT_MIN_READS", "2000")


def test_group_hot_bucket_split_matches_oracle(hot_env):
    d, cfg = small("C2", 0.004)
    okept, _, octr = oracle(d)
    cuts, keys, hot, cost = umigpu.shard_plan_sorted(d["tid"], d["pos"], d["rev"], 4, hot_min_reads=2000)
    assert hot.present and hot.reads_est > 2000, "this input must have a bucket worth splitting"
    for devices in device_lists():
        with umigpu.Group(cfg["umi_len"], devices) as g:
            for rep in range(2):                      # second call: windows and child contexts are reused
                kept, ctr, ms = g.dedup(d["tid"], d["pos"], d["rev"], d["umi"], d["score"])
                assert kept.astype(np.int64).tolist() == okept.tolist(), devices
                for key in CTR_SUM:
                    assert ctr[key] == octr[key], (devices, key)
                assert ctr["max_umis"] == octr["max_umis"]


def test_group_reused_for_different_datasets(hot_env):
    """One group, three different datasets in a row (different hot buckets, sizes, owners): nothing of an earlier call —
    window contents, hand-over words, child buffers — may leak into the next."""
    devs = device_lists()[-1]
    if len(devs) < 3:
        devs = [0, 0, 0]
    with umigpu.Group(12, devs) as g:
        for seed, scale in ((1, 0.004), (2, 0.002), (3, 0.006), (1, 0.004)):
            d, cfg = small("C2", scale, seed=seed)
            okept, _, octr = oracle(d)
            kept, ctr, _ = g.dedup(d["tid"], d["pos"], d["rev"], d["umi"], d["score"])
            assert kept.astype(np.int64).tolist() == okept.tolist(), (seed, scale)
            assert ctr["total_umis"] == octr["total_umis"]


@pytest.mark.parametrize("algo,oalgo,k,L", [(umigpu.ALGO_CC, O.ALGO_CC, 2, 16), (umigpu.ALGO_ADJ_UPSTREAM, O.ALGO_ADJ_UPSTREAM, 1, 12),
                                            (umigpu.ALGO_ADJ, O.ALGO_ADJ_REF, 1, 12)])
def test_group_other_algorithms(hot_env, algo, oalgo, k, L):
    d, cfg = small("C2", 0.002, seed=11, umi_len=L)
    okept, _, octr = oracle(d, oalgo, k)
    with umigpu.Group(L, [0, 0, 0], k=k, algo=algo) as g:
        kept, ctr, _ = g.dedup(d["tid"], d["pos"], d["rev"], d["umi"], d["score"])
    assert kept.astype(np.int64).tolist() == okept.tolist()
    assert ctr["n_buckets"] == octr["n_buckets"] and ctr["total_umis"] == octr["total_umis"]


def test_group_with_n_bases(hot_env):
    d, cfg = small("C2", 0.002, seed=5, n_rate=0.01)
    okept, _, _ = oracle(d)
    with umigpu.Group(cfg["umi_len"], [0, 0]) as g:
        kept, _, _ = g.dedup(d["tid"], d["pos"], d["rev"], d["umi"], d["score"])
    assert kept.astype(np.int64).tolist() == okept.tolist()


def test_group_unsorted_input_is_detected_and_rerouted(hot_env):
    """A stream that is not coordinate-sorted cannot be cut into contiguous slices: the devices notice (range check of
    the cuts) and the call goes through the hash plan instead — never a wrong answer."""
    d, cfg = small("C2", 0.001, seed=3)
    rng = np.random.default_rng(1)
    perm = rng.permutation(len(d["tid"]))
    d = {k: np.ascontiguousarray(v[perm]) for k, v in d.items()}
    okept, _, octr = oracle(d)
    with umigpu.Group(cfg["umi_len"], [0, 0, 0]) as g:
        kept, ctr, _ = g.dedup(d["tid"], d["pos"], d["rev"], d["umi"], d["score"])
    assert kept.astype(np.int64).tolist() == okept.tolist()
    assert ctr["n_buckets"] == octr["n_buckets"]


def test_group_many_contigs_and_empty_slices(hot_env):
    d, cfg = small("C1", 0.002, seed=2, n_contigs=7)
    okept, _, _ = oracle(d)
    with umigpu.Group(cfg["umi_len"], [0] * 8) as g:
        kept, _, _ = g.dedup(d["tid"], d["pos"], d["rev"], d["umi"], d["score"])
    assert kept.astype(np.int64).tolist() == okept.tolist()
    # fewer distinct positions than devices: some slices are empty
    n = 3000
    rng = np.random.default_rng(4)
    tid = np.zeros(n, np.int32); pos = np.sort(rng.integers(0, 2, n)).astype(np.int64) * 100; rev = np.zeros(n, np.uint8)
    umi = rng.choice(np.frombuffer(b"ACGT", np.uint8), (n, 8)); score = rng.integers(2, 40, n).astype(np.int32)
    ok2, _, _ = O.dedup(tid, pos, rev, umi, score, O.ALGO_DIR, O.MERGE_AVGQUAL, 1, 0.5)
    with umigpu.Group(8, [0] * 6) as g:
        kept, _, _ = g.dedup(tid, pos, rev, umi, score)
    assert kept.astype(np.int64).tolist() == ok2.tolist()


def test_per_rank_api_with_threads(hot_env):
    """The calls a multi-process host makes (plan, window, attach, push the slice, run_sharded, fetch), here from threads."""
    import threading
    d, cfg = small("C2", 0.004, seed=21)
    okept, _, _ = oracle(d)
    N = 3
    cuts, keys, hot, _ = umigpu.shard_plan_sorted(d["tid"], d["pos"], d["rev"], N, hot_min_reads=2000)
    assert hot.present
    ctxs = [umigpu.Context(cfg["umi_len"], 1, 0.5, umigpu.ALGO_DIR, umigpu.MERGE_AVGQUAL, 0) for _ in range(N)]
    for r, c in enumerate(ctxs):
        c.xchg_create(r, N, int(hot.reads_est) + 1024, 1 << 20)
    for c in ctxs:
        c.xchg_attach_local(ctxs)
    out, errs = [None] * N, []

    def work(r):
        try:
            a, b = int(cuts[r]), int(cuts[r + 1])
            c = ctxs[r]
            c.reset()
            if b > a:
                c.push_reads(d["tid"][a:b], d["pos"][a:b], d["rev"][a:b], d["umi"][a:b], d["score"][a:b], None, a)
            c.run_sharded(hot, int(keys[r]), int(keys[r + 1]))
            out[r] = c.fetch()[0]
        except Exception as e:      # noqa: BLE001
            errs.append(repr(e))

    for rep in range(2):
        th = [threading.Thread(target=work, args=(r,)) for r in range(N)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        assert not errs, errs
        merged = np.concatenate(out).astype(np.int64)
        assert merged.tolist() == okept.tolist()
    for c in ctxs:
        c.close()


def test_exchange_window_too_small_is_an_error(hot_env):
    d, cfg = small("C2", 0.004, seed=21)
    cuts, keys, hot, _ = umigpu.shard_plan_sorted(d["tid"], d["pos"], d["rev"], 1, hot_min_reads=2000)
    with umigpu.Context(cfg["umi_len"], 1, 0.5, umigpu.ALGO_DIR, umigpu.MERGE_AVGQUAL, 0) as c:
        c.xchg_create(0, 1, 64, 1024)
        c.xchg_attach_local([c])
        c.push_reads(d["tid"], d["pos"], d["rev"], d["umi"], d["score"])
        with pytest.raises(umigpu.UmiGpuError) as e:
            c.run_sharded(hot)
        assert e.value.code == umigpu._lib.ERR_UNSUPPORTED and "window" in str(e.value)
        # a wrong key range is rejected before anything is exchanged
        c.reset()
        c.push_reads(d["tid"], d["pos"], d["rev"], d["umi"], d["score"])
        with pytest.raises(umigpu.UmiGpuError) as e:
            c.run_sharded(None, int(keys[0]), int(umigpu.load().umigpu_pos_key(0, int(d["pos"][len(d["pos"]) // 2]))))
        assert "range-partitioned" in str(e.value)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _ipc_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["UMIGPU_XCHG_TIMEOUT_S"] = "60"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dev = rank % torch.cuda.device_count()
    d, cfg = synth.generate_config("C2", device="cpu", scale=0.004, seed=33)
    d = {k: v.numpy() for k, v in d.items()}
    cuts, keys, hot, _ = umigpu.shard_plan_sorted(d["tid"], d["pos"], d["rev"], world, hot_min_reads=2000)
    ok = True
    with umigpu.Context(cfg["umi_len"], 1, 0.5, umigpu.ALGO_DIR, umigpu.MERGE_AVGQUAL, dev) as c:
        handle = c.xchg_create(rank, world, int(hot.reads_est) + 1024, 1 << 20)
        handles = [None] * world
        dist.all_gather_object(handles, handle)
        c.xchg_attach_ipc(handles)
        dist.barrier()
        a, b = int(cuts[rank]), int(cuts[rank + 1])
        for rep in range(2):
            c.reset()
            if b > a:
                c.push_reads(d["tid"][a:b], d["pos"][a:b], d["rev"][a:b], d["umi"][a:b], d["score"][a:b], None, a)
            c.run_sharded(hot, int(keys[rank]), int(keys[rank + 1]))
            kept = c.fetch()[0].astype(np.int64)
            full, _, _ = O.dedup(d["tid"], d["pos"], d["rev"], d["umi"], d["score"], O.ALGO_DIR, O.MERGE_AVGQUAL, 1, 0.5)
            mine = full[(full >= a) & (full < b)]
            ok = ok and bool(hot.present) and kept.tolist() == mine.tolist()
        dist.barrier()
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_ipc_ranks_in_separate_processes():
    """One process per rank (what bench.py does under torchrun): windows exchanged as CUDA IPC handles through gloo."""
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_ipc_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(500)
        assert p.exitcode == 0
    res = dict(q.get() for _ in range(world))
    assert res == {0: True, 1: True}


def test_frontier_clustering_matches_oracle(monkeypatch):
    """K6 frontier form (edges sorted by source, rounds over the lowered UMIs only), forced on small inputs."""
    monkeypatch.setenv("UMIGPU_FRONTIER_FORCE", "1")
    monkeypatch.setenv("UMIGPU_FRONTIER_MIN_EDGES", "1")
    for name, scale, seed in (("C2", 0.004, 2), ("C1", 0.05, 1), ("C4", 0.002, 4)):
        d, cfg = small(name, scale, seed=seed)
        for algo, oalgo in ((umigpu.ALGO_DIR, O.ALGO_DIR), (umigpu.ALGO_CC, O.ALGO_CC)):
            with umigpu.Context(cfg["umi_len"], 1, 0.5, algo, umigpu.MERGE_AVGQUAL, 0, umigpu.FLAG_LABELS) as ctx:
                ctx.push_reads(d["tid"], d["pos"], d["rev"], d["umi"], d["score"])
                kept, roots, ctr = ctx.finish()
            okept, oroots, _ = O.dedup(d["tid"], d["pos"], d["rev"], d["umi"], d["score"], oalgo, O.MERGE_AVGQUAL, 1, 0.5, want_roots=True)
            assert kept.astype(np.int64).tolist() == okept.tolist(), (name, algo)
            assert roots.astype(np.int64).tolist() == oroots.tolist(), (name, algo)
    # a long chain of frequency-1 UMIs (every edge mutual): one hop per round without in-place luck
    L, n = 16, 600
    umis = np.full((n, L), ord("A"), np.uint8)
    for i in range(1, n):
        umis[i] = umis[i - 1]
        umis[i, i % L] = b"ACGT"[(b"ACGT".index(bytes([umis[i, i % L]])) + 1) % 4]
    umis = np.unique(umis, axis=0)
    n = len(umis)
    tid = np.zeros(n, np.int32); pos = np.zeros(n, np.int64); rev = np.zeros(n, np.uint8); score = np.arange(n).astype(np.int32) % 37
    with umigpu.Context(L, 1, 0.5, umigpu.ALGO_DIR, umigpu.MERGE_AVGQUAL, 0) as ctx:
        ctx.push_reads(tid, pos, rev, umis, score)
        kept, _, ctr = ctx.finish()
    okept, _, _ = O.dedup(tid, pos, rev, umis, score, O.ALGO_DIR, O.MERGE_AVGQUAL, 1, 0.5)
    assert kept.astype(np.int64).tolist() == okept.tolist()


def test_full_size_c2_sharded_equals_unsharded():
    """C2 at FULL size (50 M reads) over 4 ranks — all the box's GPUs when it has several, else four contexts on device 0:
    contiguous slices + hot-bucket split must keep exactly the reads of the one-context run."""
    import threading

    import torch
    d, cfg = synth.generate_config("C2", device="cuda:0", scale=1.0)
    with umigpu.Context(cfg["umi_len"], 1, 0.5, umigpu.ALGO_DIR, umigpu.MERGE_AVGQUAL, 0) as c:
        c.push_reads(d["tid"], d["pos"], d["rev"], d["umi"], d["score"])
        full, _, fctr = c.finish()
    h = {k: v.cpu().numpy() for k, v in d.items()}
    del d
    nd = torch.cuda.device_count()
    N = 4
    devices = [r % nd for r in range(N)]
    cuts, keys, hot, _ = umigpu.shard_plan_sorted(h["tid"], h["pos"], h["rev"], N)
    assert hot.present and hot.reads_est > 1_000_000
    ctxs = [umigpu.Context(cfg["umi_len"], 1, 0.5, umigpu.ALGO_DIR, umigpu.MERGE_AVGQUAL, devices[r]) for r in range(N)]
    for r, c in enumerate(ctxs):
        c.xchg_create(r, N, int(hot.reads_est) + 4096, 4 * int(hot.reads_est))
    for c in ctxs:
        c.xchg_attach_local(ctxs)
    out, errs = [None] * N, []

    def work(r):
        try:
            a, b = int(cuts[r]), int(cuts[r + 1])
            c = ctxs[r]
            c.push_reads(h["tid"][a:b], h["pos"][a:b], h["rev"][a:b], h["umi"][a:b], h["score"][a:b], None, a)
            c.run_sharded(hot, int(keys[r]), int(keys[r + 1]))
            out[r] = c.fetch()
        except Exception as e:      # noqa: BLE001
            errs.append(repr(e))

    th = [threading.Thread(target=work, args=(r,)) for r in range(N)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs
    merged = np.concatenate([o[0] for o in out])
    assert np.array_equal(merged, full)
    assert sum(o[2]["n_buckets"] for o in out) == fctr["n_buckets"] and sum(o[2]["n_edges"] for o in out) == fctr["n_edges"]
    for c in ctxs:
        c.close()
