"""K2 for coordinate-sorted input (seg_sort.cuh): segments of equal (contig, position) are sorted in place — windows in
shared memory for small segments, batched one-sweep passes over (strand, UMI) for big ones.  Every shape is compared with
the CPU oracle and with the generic LSD sort (UMIGPU_NO_SEG_SORT=1): kept reads, cluster roots and counters must agree."""
import numpy as np
import pytest

import oracle_lib as O
import umigpu
from umigpu import synth

pytestmark = pytest.mark.gpu


def run(d, L, algo=umigpu.ALGO_DIR, k=1, labels=True, tlen=None):
    with umigpu.Context(L, k, 0.5, algo, umigpu.MERGE_AVGQUAL, 0, umigpu.FLAG_LABELS if labels else 0) as ctx:
        ctx.push_reads(d["tid"], d["pos"], d["rev"], d["umi"], d["score"], tlen=tlen)
        kept, roots, ctr = ctx.finish()
    return kept, roots, ctr


def both_sorts(monkeypatch, d, L, algo=umigpu.ALGO_DIR, oalgo=O.ALGO_DIR, k=1, tlen=None, oracle=True):
    monkeypatch.delenv("UMIGPU_NO_SEG_SORT", raising=False)
    kept, roots, ctr = run(d, L, algo, k, tlen=tlen)
    monkeypatch.setenv("UMIGPU_NO_SEG_SORT", "1")
    gkept, groots, gctr = run(d, L, algo, k, tlen=tlen)
    monkeypatch.delenv("UMIGPU_NO_SEG_SORT", raising=False)
    assert np.array_equal(kept, gkept) and np.array_equal(roots, groots)
    for key in ("n_buckets", "total_umis", "max_umis", "n_kept", "unordered_pairs", "n_edges"):
        assert ctr[key] == gctr[key], key
    if oracle:
        okept, oroots, octr = O.dedup(d["tid"], d["pos"], d["rev"], d["umi"], d["score"], oalgo, O.MERGE_AVGQUAL, k, 0.5, want_roots=True, tlen=tlen)
        assert kept.astype(np.int64).tolist() == okept.tolist()
        if algo != umigpu.ALGO_ADJ:
            assert roots.astype(np.int64).tolist() == oroots.tolist()
        assert ctr["n_buckets"] == octr["n_buckets"] and ctr["total_umis"] == octr["total_umis"]
    return ctr


def small(name, scale, seed=None, **kw):
    d, cfg = synth.generate_config(name, seed=seed, device="cpu", scale=scale, **kw)
    return {k: v.numpy() for k, v in d.items()}, cfg


@pytest.mark.parametrize("name,scale,kw", [("C1", 0.05, {}), ("C2", 0.004, {}), ("C2", 0.004, dict(n_contigs=5)), ("C3", 0.002, {}),
                                           ("C4", 0.002, {}), ("C2", 0.002, dict(n_rate=0.02)), ("C2", 0.0004, dict(umi_len=19))])
def test_baseline_shapes(monkeypatch, name, scale, kw):
    d, cfg = small(name, scale, **kw)
    algo, oalgo = (umigpu.ALGO_CC, O.ALGO_CC) if cfg["algo"] == "cc" else (umigpu.ALGO_DIR, O.ALGO_DIR)
    both_sorts(monkeypatch, d, cfg["umi_len"], algo, oalgo, cfg["k"])


def test_segment_sizes_around_every_threshold(monkeypatch):
    """Segments of 1 .. 20000 reads laid end to end so that heads fall on every offset of the 2048-position blocks: windows
    with one and with thousands of segments, big segments of one tile and of several, a big segment at the very start and at
    the very end, both strands interleaved inside a position."""
    rng = np.random.default_rng(11)
    L = 9
    sizes = [1, 1, 1, 2, 31, 32, 33, 2047, 1, 2048, 2049, 1, 1, 4095, 4096, 4097, 6143, 6144, 6145, 1, 12289, 20000, 3] + [1] * 3000 + [5, 700, 2048, 2050]
    for order in (sizes, sizes[::-1]):
        pos = np.repeat(np.arange(len(order), dtype=np.int64) * 3 - 7, order)
        n = len(pos)
        d = dict(tid=np.zeros(n, np.int32), pos=pos, rev=rng.integers(0, 2, n).astype(np.uint8),
                 umi=rng.choice(np.frombuffer(b"ACGT", np.uint8), (n, L)), score=rng.integers(2, 41, n).astype(np.int32))
        # few distinct UMIs per position so that counting / merging sees long runs of equal keys too
        d["umi"][:, :5] = ord("A")
        ctr = both_sorts(monkeypatch, d, L)
        assert ctr["max_umis"] > 32


def test_unsorted_and_nearly_sorted_inputs_take_the_generic_sort(monkeypatch):
    d, cfg = small("C2", 0.002, seed=5)
    n = len(d["tid"])
    rng = np.random.default_rng(2)
    # (a) clip-displaced: reverse-strand reads carry end_pos + clips (utils/mod.rs:96-104), so a BAM sorted by leftmost
    #     coordinate is only nearly sorted by the bucket coordinate
    e = {k: v.copy() for k, v in d.items()}
    e["pos"] = e["pos"] + np.where(e["rev"] == 1, rng.integers(0, 150, n), 0)
    both_sorts(monkeypatch, e, cfg["umi_len"])
    # (b) one swapped pair of positions is enough to leave the segmented path
    f = {k: v.copy() for k, v in d.items()}
    i = n // 2
    while f["pos"][i] == f["pos"][i + 1]:
        i += 1
    for k in f:
        f[k][[i, i + 1]] = f[k][[i + 1, i]]
    both_sorts(monkeypatch, f, cfg["umi_len"])
    # (c) fully shuffled
    perm = rng.permutation(n)
    g = {k: np.ascontiguousarray(v[perm]) for k, v in d.items()}
    both_sorts(monkeypatch, g, cfg["umi_len"])


def test_paired_keys_and_wide_umis(monkeypatch):
    """tlen widens S; a 21-nt UMI with tlen does not fit the packed element any more -> generic sort, same answer."""
    rng = np.random.default_rng(8)
    for L, n in ((12, 40000), (21, 9000)):
        pos = np.sort(rng.integers(0, 40, n)).astype(np.int64) * 5
        pool = rng.choice(np.frombuffer(b"ACGT", np.uint8), (300, L))
        d = dict(tid=np.zeros(n, np.int32), pos=pos, rev=rng.integers(0, 2, n).astype(np.uint8), umi=pool[rng.integers(0, 300, n)],
                 score=rng.integers(2, 41, n).astype(np.int32))
        tlen = rng.choice(np.array([-300, 150, 151, 152, 4000], np.int64), n)
        both_sorts(monkeypatch, d, L, tlen=tlen)


def test_full_size_c2_segmented_equals_generic(monkeypatch):
    d, cfg = synth.generate_config("C2", device="cuda", scale=1.0)
    outs = []
    for env in (None, "1"):
        if env:
            monkeypatch.setenv("UMIGPU_NO_SEG_SORT", env)
        with umigpu.Context(cfg["umi_len"], 1, 0.5, umigpu.ALGO_DIR, umigpu.MERGE_AVGQUAL, 0) as ctx:
            ctx.push_reads(d["tid"], d["pos"], d["rev"], d["umi"], d["score"])
            kept, _, ctr = ctx.finish()
            outs.append((kept, ctr, ctx.stage_ms()["sort"]))
    assert np.array_equal(outs[0][0], outs[1][0])
    for key in ("n_buckets", "total_umis", "max_umis", "n_kept", "n_edges"):
        assert outs[0][1][key] == outs[1][1][key]
    print("sort ms: segmented %.2f generic %.2f" % (outs[0][2], outs[1][2]))


def test_bulk_copy_staging_variant_of_the_block_pair_kernel(monkeypatch):
    """hamming_blocks<..., BULK>: the next pair's column words arrive by cp.async.bulk + mbarrier (double buffered).  Same
    edges, same survivors as the plain-load variant and the oracle, for every k and a few lengths."""
    ran = 0
    for L, k, name, scale in ((12, 1, "C2", 0.004), (16, 2, "C3", 0.002), (8, 3, "C2", 0.001), (12, 1, "C4", 0.003), (30, 1, "C2", 0.001)):
        d, cfg = small(name, scale, seed=L, umi_len=L)
        outs = []
        for env in ("0", "1"):
            monkeypatch.setenv("UMIGPU_K5_BULK", env)
            with umigpu.Context(L, k, 0.5, umigpu.ALGO_DIR, umigpu.MERGE_AVGQUAL, 0) as ctx:
                ctx.push_reads(d["tid"], d["pos"], d["rev"], d["umi"], d["score"])
                kept, _, ctr = ctx.finish()
            outs.append((kept, ctr))
        monkeypatch.delenv("UMIGPU_K5_BULK")
        assert np.array_equal(outs[0][0], outs[1][0]) and outs[0][1]["n_edges"] == outs[1][1]["n_edges"]
        ran += outs[1][1]["n_block_pairs"] > 0
        okept, _, _ = O.dedup(d["tid"], d["pos"], d["rev"], d["umi"], d["score"], O.ALGO_DIR, O.MERGE_AVGQUAL, k, 0.5)
        assert outs[1][0].astype(np.int64).tolist() == okept.tolist()
    assert ran >= 1, "the block-pair kernel must have run in some of the cases"
