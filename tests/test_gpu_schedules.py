"""Schedules that changed late in round 2, each against the one it replaced (selected by an environment variable) and against
the CPU oracle: keys + block table in one kernel (UMIGPU_NO_FUSED_SUMMARY), the one-pass merge scan (UMIGPU_UNIQUE_TWO_PASS),
the speculative block-pair expansion (UMIGPU_EXPAND_TWICE), big-bucket counts without their own read-backs (UMIGPU_BIG_READBACK)."""
import numpy as np
import pytest

import oracle_lib as O
import umigpu
from umigpu import synth

pytestmark = pytest.mark.gpu

FORMER = ("UMIGPU_NO_FUSED_SUMMARY", "UMIGPU_UNIQUE_TWO_PASS", "UMIGPU_EXPAND_TWICE", "UMIGPU_BIG_READBACK")


def run(d, L, algo=umigpu.ALGO_DIR, k=1, merge=umigpu.MERGE_AVGQUAL, runs=1):
    with umigpu.Context(L, k, 0.5, algo, merge, 0, umigpu.FLAG_LABELS) as ctx:
        for _ in range(runs):                       # a second run re-uses the first one's buffers (the speculative sizes)
            ctx.reset()
            ctx.push_reads(d["tid"], d["pos"], d["rev"], d["umi"], d["score"])
            kept, roots, ctr = ctx.finish()
            kept, roots = kept.copy(), roots.copy()
    return kept, roots, ctr


def all_schedules(monkeypatch, d, L, algo=umigpu.ALGO_DIR, oalgo=O.ALGO_DIR, k=1, oracle=True):
    for v in FORMER:
        monkeypatch.delenv(v, raising=False)
    kept, roots, ctr = run(d, L, algo, k, runs=2)
    for v in FORMER:
        monkeypatch.setenv(v, "1")
        k2, r2, c2 = run(d, L, algo, k)
        monkeypatch.delenv(v)
        assert np.array_equal(kept, k2) and np.array_equal(roots, r2), v
        for key in ("n_buckets", "total_umis", "max_umis", "n_kept", "unordered_pairs", "n_edges"):
            assert ctr[key] == c2[key], (v, key)
    if oracle:
        okept, oroots, octr = O.dedup(d["tid"], d["pos"], d["rev"], d["umi"], d["score"], oalgo, O.MERGE_AVGQUAL, k, 0.5, want_roots=True)
        assert kept.astype(np.int64).tolist() == okept.tolist()
        assert roots.astype(np.int64).tolist() == oroots.tolist()
        assert ctr["n_buckets"] == octr["n_buckets"] and ctr["total_umis"] == octr["total_umis"]
    return ctr


def small(name, scale, **kw):
    d, cfg = synth.generate_config(name, device="cpu", scale=scale, **kw)
    return {k: v.numpy() for k, v in d.items()}, cfg


@pytest.mark.parametrize("name,scale,kw", [("C1", 0.05, {}), ("C2", 0.004, {}), ("C2", 0.004, dict(n_contigs=5)), ("C3", 0.002, {}),
                                           ("C2", 0.002, dict(n_rate=0.02))])
def test_baseline_shapes_under_every_schedule(monkeypatch, name, scale, kw):
    d, cfg = small(name, scale, **kw)
    algo, oalgo = (umigpu.ALGO_CC, O.ALGO_CC) if cfg["algo"] == "cc" else (umigpu.ALGO_DIR, O.ALGO_DIR)
    all_schedules(monkeypatch, d, cfg["umi_len"], algo, oalgo, cfg["k"])


def test_uniques_that_span_many_scan_tiles(monkeypatch):
    """A handful of UMIs with thousands of reads each: a unique's reads cover several 2048-read tiles of the merge scan, so
    tiles consist of continuation reads only and the representative (first read with the best score) is decided by values that
    travel through the per-tile carry.  Sorted and shuffled input; the best score is planted early, late and in the middle."""
    rng = np.random.default_rng(5)
    L = 8
    letters = np.frombuffer(b"ACGT", dtype=np.uint8)
    for shuffled in (False, True):
        counts = [9000, 1, 2047, 2048, 2049, 30000, 3, 4096, 1, 1, 12000, 5]
        umis = letters[rng.integers(0, 4, size=(len(counts), L))]
        rows, pos, score = [], [], []
        for j, c in enumerate(counts):
            rows.append(np.repeat(umis[j:j + 1], c, axis=0))
            pos.append(np.full(c, 100 + (j // 4), dtype=np.int64))            # four UMIs share a position: buckets of several uniques
            s = rng.integers(2, 30, size=c).astype(np.int32)
            for at in (0, c // 2, c - 1):                                      # ties for the maximum: the FIRST one must win
                if rng.random() < 0.7:
                    s[at] = 40
            score.append(s)
        umi = np.concatenate(rows); pos = np.concatenate(pos); score = np.concatenate(score)
        n = len(pos)
        d = dict(tid=np.zeros(n, np.int32), pos=pos, rev=np.zeros(n, np.uint8), umi=umi, score=score)
        if shuffled:
            p = rng.permutation(n)
            d = {k: np.ascontiguousarray(v[p]) for k, v in d.items()}
        else:
            o = np.argsort(d["pos"], kind="stable")
            d = {k: np.ascontiguousarray(v[o]) for k, v in d.items()}
        ctr = all_schedules(monkeypatch, d, L)
        assert ctr["total_umis"] == len({bytes(u) + bytes([j // 4]) for j, u in enumerate(umis)})


def test_big_buckets_under_every_schedule(monkeypatch):
    """Buckets beyond the multi-index threshold (4096 unique UMIs): the block-pair list is expanded speculatively, grown when
    the first guess was too small, and re-used by the second run; the big-bucket counts come with the bucket statistics."""
    d, cfg = small("C2", 0.02)                      # 1 M reads, hottest locus ~30 k unique UMIs
    ctr = all_schedules(monkeypatch, d, cfg["umi_len"])
    assert ctr["max_umis"] > 20000 and ctr["n_block_pairs"] > 0


def test_reset_needs_no_host_synchronisation():
    """umigpu_reset uploads the scalar block from its own pinned image: back-to-back batches on one context, no fetch between
    a run and the next reset, must not see each other's counters."""
    a, cfg = small("C1", 0.02)
    b, _ = small("C1", 0.01, seed=77)
    L = cfg["umi_len"]
    with umigpu.Context(L, 1, 0.5, umigpu.ALGO_DIR, umigpu.MERGE_AVGQUAL, 0) as ctx:
        outs = []
        for d in (a, b, a, b):
            ctx.reset()
            ctx.push_reads(d["tid"], d["pos"], d["rev"], d["umi"], d["score"])
            ctx.run()
            kept, _, ctr = ctx.fetch()
            outs.append((kept.copy(), ctr["total_reads"], ctr["total_umis"], ctr["unordered_pairs"]))
    assert np.array_equal(outs[0][0], outs[2][0]) and outs[0][1:] == outs[2][1:]
    assert np.array_equal(outs[1][0], outs[3][0]) and outs[1][1:] == outs[3][1:]
    assert outs[0][1] == len(a["pos"]) and outs[1][1] == len(b["pos"])
