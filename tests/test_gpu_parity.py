"""GPU parity tests: the CUDA path, called through the C ABI (ctypes), against the CPU oracle on the same
seeded inputs, against the committed golden fixtures, and through size-independent properties.
Bit-exact everywhere (integer / index work)."""
import json
import os
import random

import numpy as np
import pytest

import oracle_lib as O
import umigpu
from umigpu import synth

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "cases.json")))
ORACLE_ALGO = {umigpu.ALGO_DIR: O.ALGO_DIR, umigpu.ALGO_ADJ: O.ALGO_ADJ_REF, umigpu.ALGO_ADJ_UPSTREAM: O.ALGO_ADJ_UPSTREAM,
               umigpu.ALGO_CC: O.ALGO_CC}


def arr(umis):
    return np.frombuffer("".join(umis).encode(), dtype=np.uint8).reshape(len(umis), len(umis[0]))


def gpu_dedup(tid, pos, rev, umi, score, algo, merge, k, p, flags=0, chunk=0, labels=False, tlen=None):
    ctx = umigpu.Context(umi.shape[1], k, p, algo, merge, 0, flags | (umigpu.FLAG_LABELS if labels else 0))
    n = len(tid)
    tid = np.asarray(tid, np.int32); pos = np.asarray(pos, np.int64); rev = np.asarray(rev, np.uint8)
    score = None if score is None else np.asarray(score, np.int32)
    tlen = None if tlen is None else np.asarray(tlen, np.int64)
    if chunk:
        for s in range(0, n, chunk):
            e = min(n, s + chunk)
            ctx.push_reads(tid[s:e], pos[s:e], rev[s:e], umi[s:e], None if score is None else score[s:e], None, s,
                           tlen=None if tlen is None else tlen[s:e])
    else:
        ctx.push_reads(tid, pos, rev, umi, score, None, 0, tlen=tlen)
    kept, roots, ctr = ctx.finish()
    ctx.close()
    return kept, roots, ctr


def check_against_oracle(d, algo, merge, k, p, flags=0, chunk=0, labels=False):
    kept, roots, ctr = gpu_dedup(d["tid"], d["pos"], d["rev"], d["umi"], d["score"], algo, merge, k, p, flags, chunk, labels)
    okept, oroots, octr = O.dedup(d["tid"], d["pos"], d["rev"], d["umi"], d["score"], ORACLE_ALGO[algo], merge, k, p, want_roots=labels)
    assert kept.astype(np.int64).tolist() == okept.tolist()
    for key in ("total_reads", "n_buckets", "total_umis", "max_umis", "n_kept", "unordered_pairs"):
        assert ctr[key] == octr[key], key
    if labels and algo != umigpu.ALGO_ADJ:
        assert roots.astype(np.int64).tolist() == oroots.tolist()
    return ctr


def small(name, scale, seed=None, **kw):
    d, cfg = synth.generate_config(name, seed=seed, device="cpu", scale=scale, **kw)
    return {k: v.numpy() for k, v in d.items()}, cfg


# ---------------------------------------------------------------- golden fixtures
@pytest.mark.parametrize("case", GOLD["reads"], ids=lambda c: c["name"])
def test_golden_reads(case):
    algo = {0: umigpu.ALGO_DIR, 1: umigpu.ALGO_ADJ, 2: umigpu.ALGO_ADJ_UPSTREAM, 3: umigpu.ALGO_CC}[case["algo"]]
    kept, _, ctr = gpu_dedup(case["tid"], case["pos"], case["rev"], arr(case["umi"]), case["score"], algo, case["merge"],
                             case["k"], case["p"], tlen=case.get("tlen"))
    assert kept.astype(np.int64).tolist() == case["kept"]
    for key, v in case["counters"].items():
        assert ctr[key] == v, key


@pytest.mark.parametrize("case", GOLD["buckets"], ids=lambda c: c["name"])
def test_golden_buckets(case):
    algo = {0: umigpu.ALGO_DIR, 1: umigpu.ALGO_ADJ, 2: umigpu.ALGO_ADJ_UPSTREAM, 3: umigpu.ALGO_CC}[case["algo"]]
    with umigpu.Context(case["umi_len"], case["k"], case["p"], algo, umigpu.MERGE_ANY) as ctx:
        keep, label = ctx.cluster_bucket(arr(case["umis"]), np.array(case["freq"], np.int32))
    assert keep.tolist() == case["keep"]
    if algo != umigpu.ALGO_ADJ:
        assert label.tolist() == case["label"]


# ---------------------------------------------------------------- the five BASELINE configs, scaled to oracle size
@pytest.mark.parametrize("name,scale", [("C1", 0.05), ("C2", 0.002), ("C3", 0.001), ("C4", 0.001), ("C5", 0.0005)])
def test_baseline_configs_scaled(name, scale):
    d, cfg = small(name, scale)
    algo = {"dir": umigpu.ALGO_DIR, "cc": umigpu.ALGO_CC}[cfg["algo"]]
    ctr = check_against_oracle(d, algo, umigpu.MERGE_AVGQUAL, cfg["k"], 0.5, labels=True)
    assert ctr["total_reads"] == len(d["tid"])
    if name == "C3":      # config 3 also names --algo adj (as written and upstream-intended)
        check_against_oracle(d, umigpu.ALGO_ADJ, umigpu.MERGE_AVGQUAL, cfg["k"], 0.5)
        check_against_oracle(d, umigpu.ALGO_ADJ_UPSTREAM, umigpu.MERGE_AVGQUAL, cfg["k"], 0.5, labels=True)


@pytest.mark.parametrize("algo", [umigpu.ALGO_DIR, umigpu.ALGO_ADJ, umigpu.ALGO_ADJ_UPSTREAM, umigpu.ALGO_CC])
@pytest.mark.parametrize("merge", [umigpu.MERGE_ANY, umigpu.MERGE_AVGQUAL, umigpu.MERGE_MAPQUAL])
def test_algo_merge_matrix(algo, merge):
    d, _ = small("C1", 0.02, seed=algo * 7 + merge)
    check_against_oracle(d, algo, merge, 1, 0.5, labels=True)


@pytest.mark.parametrize("flags", [0, umigpu.FLAG_KERNEL_DIRECT, umigpu.FLAG_NO_CULL, umigpu.FLAG_KERNEL_DIRECT | umigpu.FLAG_NO_CULL,
                                   umigpu.FLAG_KERNEL_TILES, umigpu.FLAG_KERNEL_TILES | umigpu.FLAG_NO_CULL, umigpu.FLAG_NO_MULTI_INDEX])
def test_large_bucket_multi_tile(flags):
    """One hot locus whose unique UMIs span several 2048-wide tiles (diagonal + off-diagonal tiles)."""
    d, _ = small("C2", 0.0008, n_loci=3, zipf_s=2.0, family=1.5, umi_len=8)
    ctr = check_against_oracle(d, umigpu.ALGO_DIR, umigpu.MERGE_AVGQUAL, 1, 0.5, flags=flags, labels=True)
    assert ctr["max_umis"] > 3 * 2048


@pytest.mark.parametrize("k,L", [(0, 8), (1, 6), (2, 8), (3, 10), (4, 12), (1, 1), (1, 32), (2, 31), (1, 17)])
def test_k_and_length_sweep(k, L):
    d, _ = small("C1", 0.004, umi_len=L, n_loci=20, seed=k * 100 + L, err=0.3)
    for algo in (umigpu.ALGO_DIR, umigpu.ALGO_CC):
        check_against_oracle(d, algo, umigpu.MERGE_AVGQUAL, k, 0.5, labels=True)


@pytest.mark.parametrize("p", [0.1, 0.3, 0.5, 0.75, 1.0])
def test_percentage_sweep(p):
    d, _ = small("C1", 0.01, umi_len=6, n_loci=10, seed=int(p * 100), family=2.0)
    check_against_oracle(d, umigpu.ALGO_DIR, umigpu.MERGE_AVGQUAL, 1, p, labels=True)


def test_n_bases_and_both_kernels():
    for flags in (0, umigpu.FLAG_KERNEL_DIRECT, umigpu.FLAG_KERNEL_TILES):
        d, _ = small("C1", 0.01, umi_len=9, n_loci=8, n_rate=0.05, seed=77)
        assert (d["umi"] == ord("N")).any()
        for algo in (umigpu.ALGO_DIR, umigpu.ALGO_CC, umigpu.ALGO_ADJ_UPSTREAM):
            check_against_oracle(d, algo, umigpu.MERGE_AVGQUAL, 1, 0.5, flags=flags, labels=True)
            check_against_oracle(d, algo, umigpu.MERGE_AVGQUAL, 2, 0.5, flags=flags)


def test_wide_keys_multi_contig_negative_positions():
    """tid spread + large coordinates + 16-nt UMIs force the two-word (128-bit) sort key; unclipped positions
    can be negative (utils/mod.rs:96-104)."""
    rng = np.random.default_rng(4)
    n = 20000
    tid = rng.integers(0, 3000, n).astype(np.int32)
    pos = (rng.integers(0, 40, n) * 6_000_000 - 90).astype(np.int64)
    rev = rng.integers(0, 2, n).astype(np.uint8)
    codes = rng.integers(0, 6, (n, 16))
    umi = np.array([65, 67, 71, 84, 65, 65], np.uint8)[codes]
    tid[: n // 2] = 7; pos[: n // 2] = -90          # one populated bucket at a negative coordinate
    score = rng.integers(0, 60, n).astype(np.int32)
    d = dict(tid=tid, pos=pos, rev=rev, umi=umi, score=score)
    check_against_oracle(d, umigpu.ALGO_DIR, umigpu.MERGE_MAPQUAL, 2, 0.5, labels=True)


def test_edge_cases_empty_single_duplicates():
    ctx = umigpu.Context(12)
    kept, roots, ctr = ctx.finish()                  # nothing pushed
    assert len(kept) == 0 and ctr["total_reads"] == 0 and ctr["n_kept"] == 0
    ctx.reset()
    one = dict(tid=np.zeros(1, np.int32), pos=np.array([5], np.int64), rev=np.zeros(1, np.uint8),
               umi=arr(["ACGTACGTACGT"]), score=np.array([30], np.int32))
    ctx.push_reads(**one)
    kept, _, ctr = ctx.finish()
    assert kept.tolist() == [0] and ctr["n_buckets"] == 1 and ctr["total_umis"] == 1
    ctx.close()
    # all reads identical: one survivor = first read with the best score
    n = 5000
    d = dict(tid=np.zeros(n, np.int32), pos=np.full(n, 9, np.int64), rev=np.ones(n, np.uint8), umi=arr(["ACGTAC"] * n),
             score=(np.arange(n) % 37).astype(np.int32))
    kept, _, ctr = gpu_dedup(d["tid"], d["pos"], d["rev"], d["umi"], d["score"], umigpu.ALGO_DIR, umigpu.MERGE_AVGQUAL, 1, 0.5)
    assert kept.tolist() == [36] and ctr["total_umis"] == 1
    kept, _, _ = gpu_dedup(d["tid"], d["pos"], d["rev"], d["umi"], d["score"], umigpu.ALGO_DIR, umigpu.MERGE_ANY, 1, 0.5)
    assert kept.tolist() == [0]


def test_bad_base_is_an_error_like_the_reference_panic():
    with umigpu.Context(4) as ctx:
        ctx.push_reads(np.zeros(2, np.int32), np.zeros(2, np.int64), np.zeros(2, np.uint8), arr(["ACGT", "ACGx"]), None)
        with pytest.raises(umigpu.UmiGpuError) as e:
            ctx.finish()
        assert e.value.code == -3 and "Unknown character" in str(e.value)


def test_chunked_push_and_first_read_index():
    d, _ = small("C1", 0.01, seed=5)
    a = check_against_oracle(d, umigpu.ALGO_DIR, umigpu.MERGE_AVGQUAL, 1, 0.5, chunk=777, labels=True)
    assert a["total_reads"] == len(d["tid"])
    # non-contiguous numbering: the host skipped reads (unmapped etc.) between chunks
    n = len(d["tid"]); h = n // 2
    with umigpu.Context(d["umi"].shape[1]) as ctx:
        ctx.push_reads(d["tid"][:h], d["pos"][:h], d["rev"][:h], d["umi"][:h], d["score"][:h], None, 100)
        ctx.push_reads(d["tid"][h:], d["pos"][h:], d["rev"][h:], d["umi"][h:], d["score"][h:], None, 100 + h + 1000)
        kept, _, _ = ctx.finish()
    okept, _, _ = O.dedup(d["tid"], d["pos"], d["rev"], d["umi"], d["score"], O.ALGO_DIR, O.MERGE_AVGQUAL, 1, 0.5)
    expect = [100 + i if i < h else 100 + 1000 + i for i in okept.tolist()]
    assert kept.astype(np.int64).tolist() == expect


def test_device_resident_inputs():
    import torch
    d, cfg = synth.generate_config("C2", device="cuda", scale=0.004)
    with umigpu.Context(cfg["umi_len"]) as ctx:
        ctx.push_reads(d["tid"], d["pos"], d["rev"], d["umi"], d["score"])
        kept, _, ctr = ctx.finish()
    h = {k: v.cpu().numpy() for k, v in d.items()}
    okept, _, octr = O.dedup(h["tid"], h["pos"], h["rev"], h["umi"], h["score"], O.ALGO_DIR, O.MERGE_AVGQUAL, 1, 0.5)
    assert kept.astype(np.int64).tolist() == okept.tolist() and ctr["n_buckets"] == octr["n_buckets"]


# ---------------------------------------------------------------- trait-shaped entries
def test_algorithm_apply_shape_random_buckets():
    rng = random.Random(2)
    for trial in range(40):
        L = rng.choice([5, 8, 12, 20])
        n = rng.randint(1, min(300, 3 ** L))
        s = set()
        while len(s) < n:
            s.add("".join(rng.choice("ACG") for _ in range(L)) if rng.random() < 0.5 else "A" * (L - 2) + rng.choice("ACGT") + rng.choice("ACGT"))
        umis = sorted(s); rng.shuffle(umis)
        freq = np.array([rng.choice([1, 1, 1, 2, 3, 8, 50]) for _ in umis], np.int32)
        for algo in (umigpu.ALGO_DIR, umigpu.ALGO_CC, umigpu.ALGO_ADJ_UPSTREAM, umigpu.ALGO_ADJ):
            k, p = rng.choice([1, 2]), rng.choice([0.5, 0.4])
            with umigpu.Context(L, k, p, algo, umigpu.MERGE_ANY) as ctx:
                keep, label = ctx.cluster_bucket(arr(umis), freq)
            okeep, olabel, _ = O.cluster_bucket(arr(umis), freq, ORACLE_ALGO[algo], k, p)
            assert keep.tolist() == okeep.tolist()
            if algo != umigpu.ALGO_ADJ:
                assert label.tolist() == olabel.tolist()


def test_reference_shaped_objects():
    """Directional.apply / Naive.remove_near / Naive.contains read like the reference's call sites."""
    args = umigpu.Cli(k=1, percentage=0.5)
    umis = [b"ACGT", b"TCGT", b"CCGT", b"ACAT", b"ACAG", b"AAAT"]
    freq = [456, 2, 2, 72, 1, 90]
    reads = {u: umigpu.ReadFreq(read=f"read_of_{u.decode()}", freq=f) for u, f in zip(umis, freq)}
    assert umigpu.Directional(args).apply(reads, None, 4) == ["read_of_ACGT", "read_of_AAAT"]       # freq-descending order
    assert umigpu.ConnectedComponents(args).apply(reads, None, 4) == ["read_of_ACGT"]
    assert len(umigpu.Adjacency(args).apply(reads, None, 4)) == 6
    data = umigpu.Naive.new(dict(zip(umis, freq)), 4, 1)
    assert data.contains(b"ACGT")
    near = data.remove_near(b"ACGT", 1, (456 + 1) // 2)                 # directional.rs:38-39
    assert near == {b"ACGT", b"TCGT", b"CCGT", b"ACAT"} and not data.contains(b"ACAT") and data.contains(b"ACAG")
    assert data.remove_near(b"ACAT", 1, (72 + 1) // 2) == {b"ACAG"}     # ACAT itself is already gone; AAAT (90) is too frequent
    assert data.remove_near(b"AAAT", 1, 0) == {b"AAAT"}                 # adjacency.rs:56: max_freq 0 removes only the query


def test_remove_near_matches_oracle():
    rng = random.Random(8)
    for L, alpha in ((6, "ACGT"), (12, "ACGTN"), (32, "ACGT"), (21, "ACGTN")):
        umis = sorted({"".join(rng.choice(alpha[:3] if rng.random() < 0.8 else alpha) for _ in range(L)) for _ in range(800)})
        a = arr(umis)
        freq = np.array([rng.choice([1, 2, 5, 30]) for _ in umis], np.int32)
        with umigpu.Context(L) as ctx:
            for k, mf in ((0, 5), (1, 0), (1, 2), (2, 100), (3, 1)):
                q = rng.choice(umis).encode()
                assert ctx.remove_near(a, freq, q, k, mf).tolist() == O.remove_near(a, freq, q, k, mf).tolist()


def test_neighbours_csr_matches_bruteforce():
    rng = random.Random(6)
    L = 7
    umis = sorted({"".join(rng.choice("ACGT") for _ in range(L)) for _ in range(5000)})
    rng.shuffle(umis)
    a = arr(umis)
    freq = np.array([rng.choice([1, 1, 2, 3, 9]) for _ in umis], np.int32)
    n = len(umis)
    dist = (a[:, None, :] != a[None, :, :]).sum(-1)
    thr = np.array([O.lib().oracle_dir_threshold(0.5, int(f)) for f in freq])
    for k, rule in ((1, True), (1, False), (2, True)):
        with umigpu.Context(L, k, 0.5) as ctx:
            row_ptr, col = ctx.neighbours(a, freq, apply_rule=rule)
        adj = (dist <= k) & ~np.eye(n, dtype=bool)
        if rule:
            adj &= freq[None, :] <= thr[:, None]
        assert row_ptr[-1] == adj.sum() == len(col)
        exp_col = np.nonzero(adj)[1]
        assert np.array_equal(np.diff(row_ptr.astype(np.int64)), adj.sum(1))
        assert np.array_equal(col, exp_col.astype(np.uint32))


def test_avg_qual_kernel_matches_reference_formula():
    rng = np.random.default_rng(1)
    lens = np.concatenate([rng.integers(0, 300, 500), [0, 1, 70000, 65536, 65537]])
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    q = rng.integers(0, 94, int(offs[-1])).astype(np.uint8)
    with umigpu.Context(8) as ctx:
        out = ctx.avg_qual(q, offs)
    for i in range(len(lens)):
        assert out[i] == O.avg_qual(q[int(offs[i]): int(offs[i + 1])]), i


# ---------------------------------------------------------------- size-independent properties at larger sizes
def test_properties_at_scale():
    """No oracle at this size: idempotence, monotonicity between algorithms, and exact-UMI dedup for adj."""
    import torch
    d, cfg = synth.generate_config("C2", device="cuda", scale=0.1)       # 5M reads
    n = d["tid"].shape[0]
    res = {}
    for algo in (umigpu.ALGO_ADJ, umigpu.ALGO_DIR, umigpu.ALGO_CC):
        with umigpu.Context(cfg["umi_len"], 1, 0.5, algo, umigpu.MERGE_AVGQUAL) as ctx:
            ctx.push_reads(d["tid"], d["pos"], d["rev"], d["umi"], d["score"])
            res[algo] = ctx.finish()
    kadj, _, cadj = res[umigpu.ALGO_ADJ]; kdir, _, cdir = res[umigpu.ALGO_DIR]; kcc, _, ccc = res[umigpu.ALGO_CC]
    assert cadj["n_kept"] == cadj["total_umis"]
    assert (np.diff(kdir.astype(np.int64)) > 0).all()                     # canonical order: ascending, unique
    assert set(kcc.tolist()) <= set(kdir.tolist()) <= set(kadj.tolist())
    assert cdir["pairs_evaluated"] > 0 and cdir["n_tile_candidates"] >= cdir["n_tile_items"] > 0
    # idempotence: deduplicating the survivors of cc again keeps all of them under adj, and dir survivors of dir
    idx = torch.from_numpy(kdir.astype(np.int64)).cuda()
    sub = {k: v[idx].contiguous() for k, v in d.items()}
    with umigpu.Context(cfg["umi_len"], 1, 0.5, umigpu.ALGO_ADJ, umigpu.MERGE_AVGQUAL) as ctx:
        ctx.push_reads(sub["tid"], sub["pos"], sub["rev"], sub["umi"], sub["score"])
        k2, _, c2 = ctx.finish()
    assert c2["n_kept"] == len(kdir) == c2["total_umis"]                  # every survivor is a distinct (bucket, UMI)
    # representative rule on a sample of survivors: best score, earliest read among its (bucket, UMI)
    h = {k: v.cpu().numpy() for k, v in d.items()}
    key = h["pos"].astype(np.int64) * 2 + h["rev"]
    samp = kadj[:: max(1, len(kadj) // 200)].astype(np.int64)
    for r in samp.tolist():
        same = np.nonzero((key == key[r]) & (h["umi"] == h["umi"][r]).all(1))[0]
        best = same[h["score"][same] == h["score"][same].max()].min()
        assert best == r


def test_sharded_matches_unsharded():
    """whole buckets per device, no collective; two contexts (two devices when the box has them)."""
    import torch
    from umigpu import shard
    d, cfg = small("C2", 0.004)
    devs = [0, 1] if torch.cuda.device_count() > 1 else [0, 0]
    args = umigpu.Cli(k=1, algo_str="dir", merge_str="avgqual")
    merged, ctrs, cost = shard.dedup_sharded_inprocess(args, d, devs)
    okept, _, octr = O.dedup(d["tid"], d["pos"], d["rev"], d["umi"], d["score"], O.ALGO_DIR, O.MERGE_AVGQUAL, 1, 0.5)
    assert merged.tolist() == okept.tolist()
    assert sum(c["n_buckets"] for c in ctrs) == octr["n_buckets"] and sum(c["total_reads"] for c in ctrs) == len(d["tid"])


def test_randomised_configurations_against_oracle():
    """Many small random workloads over the whole parameter space (length, k, p, algo, merge, alphabet, bucket shape):
    the CUDA path must agree with the oracle on the kept reads, the cluster roots and the counters every time."""
    rng = random.Random(20261018)
    algos = [umigpu.ALGO_DIR, umigpu.ALGO_DIR, umigpu.ALGO_CC, umigpu.ALGO_ADJ_UPSTREAM, umigpu.ALGO_ADJ]
    for trial in range(60):
        L = rng.choice([3, 4, 5, 6, 7, 9, 11, 13, 18, 21, 25, 32])
        alphabet = rng.choice(["ACGT", "ACGT", "AC", "ACGTN"]) if L <= 21 else rng.choice(["ACGT", "AC"])
        n = rng.choice([1, 2, 33, 500, 3000, 9000])
        n_pos = rng.choice([1, 1, 2, 7, 60])
        pool = ["".join(rng.choice(alphabet) for _ in range(L)) for _ in range(rng.choice([1, 3, 40, 400]))]
        tid = np.array([rng.randrange(2) for _ in range(n)], np.int32)
        pos = np.array([rng.randrange(n_pos) * 1000 - 5 for _ in range(n)], np.int64)
        rev = np.array([rng.randrange(2) for _ in range(n)], np.uint8)
        umis = []
        for _ in range(n):
            u = list(rng.choice(pool))
            for _ in range(rng.choice([0, 0, 1, 2])):
                u[rng.randrange(L)] = rng.choice(alphabet)
            umis.append("".join(u))
        score = np.array([rng.randrange(0, 5) for _ in range(n)], np.int32)
        d = dict(tid=tid, pos=pos, rev=rev, umi=arr(umis), score=score)
        algo, merge = rng.choice(algos), rng.choice([umigpu.MERGE_ANY, umigpu.MERGE_AVGQUAL, umigpu.MERGE_MAPQUAL])
        k, p = rng.choice([0, 1, 1, 2, 3, 4]), rng.choice([0.5, 0.5, 0.2, 0.9])
        flags = rng.choice([0, 0, umigpu.FLAG_NO_CULL, umigpu.FLAG_KERNEL_TILES, umigpu.FLAG_KERNEL_DIRECT])
        check_against_oracle(d, algo, merge, k, p, flags=flags, chunk=rng.choice([0, 0, 257]), labels=True)


@pytest.mark.parametrize("k,L,alphabet_n", [(1, 9, 0.0), (2, 9, 0.0), (3, 10, 0.0), (2, 12, 0.0), (1, 11, 0.02), (2, 10, 0.02)])
def test_multi_index_passes_big_buckets(k, L, alphabet_n):
    """Buckets above the multi-index threshold (4096 unique UMIs) are searched in k+1 passes, one per UMI part; every
    pair within k must be found exactly once whatever part its mismatches fall in (edge count, kept set, roots)."""
    d, _ = small("C2", 0.0012, n_loci=3, zipf_s=1.5, family=1.3, umi_len=L, err=0.3, n_rate=alphabet_n, seed=100 * k + L)
    res = {}
    for flags in (0, umigpu.FLAG_NO_MULTI_INDEX):
        for algo in (umigpu.ALGO_DIR, umigpu.ALGO_CC):
            ctr = check_against_oracle(d, algo, umigpu.MERGE_AVGQUAL, k, 0.5, flags=flags, labels=True)
            res[(flags, algo)] = ctr
    assert res[(0, umigpu.ALGO_DIR)]["max_umis"] > 4096
    for algo in (umigpu.ALGO_DIR, umigpu.ALGO_CC):
        assert res[(0, algo)]["n_edges"] == res[(umigpu.FLAG_NO_MULTI_INDEX, algo)]["n_edges"]      # no pair lost, none reported twice


def test_c2_shape_one_million_reads_exact():
    """The bench workload's generator at 1 M reads (hottest locus ~30 k unique UMIs: multi-index passes, block-pair
    kernel, small-bucket kernel and the sort all take their production paths), exact against the oracle."""
    d, cfg = small("C2", 0.02)
    ctr = check_against_oracle(d, umigpu.ALGO_DIR, umigpu.MERGE_AVGQUAL, 1, 0.5, labels=True)
    assert ctr["max_umis"] > 20000 and ctr["n_block_pairs"] > 0


def test_one_call_sharded_over_devices():
    """umigpu_dedup_sharded: LPT shard plan, one context per device on its own host thread, merge in input order."""
    import torch
    d, cfg = small("C2", 0.004, seed=9)
    nd = torch.cuda.device_count()
    for devices in ([0], [0, 0, 0], list(range(nd)) if nd > 1 else [0, 0]):
        kept, ctr = umigpu.dedup_sharded(cfg["umi_len"], devices, d["tid"], d["pos"], d["rev"], d["umi"], d["score"])
        okept, _, octr = O.dedup(d["tid"], d["pos"], d["rev"], d["umi"], d["score"], O.ALGO_DIR, O.MERGE_AVGQUAL, 1, 0.5)
        assert kept.astype(np.int64).tolist() == okept.tolist()
        for key in ("total_reads", "n_buckets", "total_umis", "max_umis", "n_kept", "unordered_pairs"):
            assert ctr[key] == octr[key], key


def test_labels_expose_cluster_root_and_umi_representative():
    """FLAG_LABELS: read_cluster_root (ClusterTracker) and read_umi_rep (the best read of the read's own UMI group)."""
    d, _ = small("C1", 0.01, seed=41)
    with umigpu.Context(d["umi"].shape[1], 1, 0.5, umigpu.ALGO_DIR, umigpu.MERGE_AVGQUAL, 0, umigpu.FLAG_LABELS) as ctx:
        ctx.push_reads(d["tid"], d["pos"], d["rev"], d["umi"], d["score"])
        kept, roots, _ = ctx.finish()
        urep = ctx.last_umi_rep
    groups = {}
    for i in range(len(d["tid"])):
        groups.setdefault((int(d["tid"][i]), int(d["pos"][i]), int(d["rev"][i]), bytes(d["umi"][i])), []).append(i)
    for members in groups.values():
        best = min(m for m in members if d["score"][m] == max(d["score"][x] for x in members))
        assert all(int(urep[m]) == best for m in members)
    assert set(roots.tolist()) == set(kept.tolist())


def test_error_convention_and_call_order():
    """Every misuse returns a negative code with a message (never aborts): the Rust shim turns these into panics."""
    import ctypes as C
    from umigpu import _lib as L
    lib = umigpu.load()
    ctx = umigpu.Context(6)
    one = (np.zeros(3, np.int32), np.zeros(3, np.int64), np.zeros(3, np.uint8), arr(["ACGTAC", "ACGTAA", "TTTTTT"]), np.array([1, 2, 3], np.int32))
    res = L.Result()
    assert lib.umigpu_fetch(ctx._h, C.byref(res)) == L.ERR_STATE and b"fetch before run" in lib.umigpu_last_error(ctx._h)
    ctx.push_reads(*one)
    with pytest.raises(umigpu.UmiGpuError, match="score/weight"):
        ctx.push_reads(one[0], one[1], one[2], one[3], None, None, 10)               # score given for one chunk only
    with pytest.raises(umigpu.UmiGpuError, match="ascending"):
        ctx.push_reads(*one, None, 1)                                                # overlapping read-index range
    kept, _, _ = ctx.finish()
    assert kept.tolist() == [1, 2]          # ACGTAA / ACGTAC are one cluster of two freq-1 UMIs, root = canonical first (ACGTAA, read 1)
    with pytest.raises(umigpu.UmiGpuError, match="reset"):
        ctx.push_reads(*one, None, 100)                                              # push after run
    assert lib.umigpu_run(ctx._h) == L.ERR_STATE
    ctx.reset()
    ctx.push_reads(*one)
    assert ctx.finish()[0].tolist() == [1, 2]
    assert lib.umigpu_push_reads(ctx._h, 3, None, None, None, None, None, None, 0) == L.ERR_ARG
    assert lib.umigpu_stage_ms(ctx._h, 99, C.byref(C.c_float())) == L.ERR_ARG
    ctx.close()


def _run_flags(d, cfg, algo, flags):
    with umigpu.Context(cfg["umi_len"], cfg["k"], 0.5, algo, umigpu.MERGE_AVGQUAL, flags=flags) as ctx:
        ctx.push_reads(d["tid"], d["pos"], d["rev"], d["umi"], d["score"])
        return ctx.finish()


@pytest.mark.parametrize("name,brute", [("C2", umigpu.FLAG_KERNEL_DIRECT | umigpu.FLAG_NO_CULL | umigpu.FLAG_NO_MULTI_INDEX),
                                        ("C4", umigpu.FLAG_KERNEL_TILES | umigpu.FLAG_NO_CULL | umigpu.FLAG_NO_MULTI_INDEX)])
def test_full_size_baseline_configs_production_equals_brute_force(name, brute):
    """BASELINE.json configs at FULL size (C2: 50 M reads, hottest locus 1.5 M unique UMIs; C4: one bucket of 11.7 M).  The
    oracle cannot run here, so the production neighbour search (multi-index passes, exact culling, block-pair lists)
    is held against the all-pairs kernels that are themselves oracle-checked at small sizes: every one of the
    ~1.7e12 (C2) / 6.8e13 (C4) unordered pairs is evaluated, nothing is culled — kept lists and edge counts must be identical.
    Plus the size-independent properties: partition invariance (buckets are independent), ascending unique output."""
    import torch
    d, cfg = synth.generate_config(name, device="cuda", scale=1.0)
    kept, _, ctr = _run_flags(d, cfg, umigpu.ALGO_DIR, 0)
    bkept, _, bctr = _run_flags(d, cfg, umigpu.ALGO_DIR, brute)
    assert bctr["pairs_evaluated"] >= bctr["unordered_pairs"] > 1e12          # really all pairs
    assert ctr["pairs_evaluated"] < bctr["pairs_evaluated"] / 20               # and the production path really culls
    assert ctr["n_edges"] == bctr["n_edges"] and ctr["n_kept"] == bctr["n_kept"]
    assert np.array_equal(kept, bkept)
    assert (np.diff(kept.astype(np.int64)) > 0).all()
    for key in ("total_reads", "n_buckets", "total_umis", "max_umis"):
        assert ctr[key] == bctr[key]
    if name == "C2":
        # partition invariance: deduplicating the reads of even and odd positions separately gives the same survivors
        parts = []
        for par in (0, 1):
            sel = torch.nonzero((d["pos"] & 1) == par).squeeze(1)
            sub = {k2: v[sel].contiguous() for k2, v in d.items()}
            pk, _, _ = _run_flags(sub, cfg, umigpu.ALGO_DIR, 0)
            parts.append(sel.cpu().numpy()[pk.astype(np.int64)])
        merged = np.sort(np.concatenate(parts)).astype(np.uint64)
        assert np.array_equal(merged, kept)


def test_two_phase_clustering_forced_matches_oracle(monkeypatch):
    """K6's second scheme (hook + jump on mutual edges, contracted propagation, expand) only takes over on graphs with
    >= 8 Mi edges that plain sweeps do not settle — sizes the oracle cannot reach.  The test knobs force it on small
    inputs, where it must reproduce the oracle's kept reads and cluster roots exactly (long chains, ties, saturated
    UMI spaces)."""
    monkeypatch.setenv("UMIGPU_SV_MIN_EDGES", "0")
    monkeypatch.setenv("UMIGPU_PLAIN_ROUNDS", "0")
    rng = random.Random(77)
    total_sweeps = 0
    for trial in range(24):
        L = rng.choice([4, 5, 6, 8, 10])
        alphabet = rng.choice(["ACGT", "AC", "ACG"])
        n = rng.choice([300, 3000, 9000])
        pool = ["".join(rng.choice(alphabet) for _ in range(L)) for _ in range(rng.choice([2, 30, 300]))]
        umis = []
        for _ in range(n):
            u = list(rng.choice(pool))
            for _ in range(rng.choice([0, 1, 1, 2])):
                u[rng.randrange(L)] = rng.choice(alphabet)
            umis.append("".join(u))
        d = dict(tid=np.zeros(n, np.int32), pos=np.array([rng.randrange(3) for _ in range(n)], np.int64),
                 rev=np.zeros(n, np.uint8), umi=arr(umis), score=np.array([rng.randrange(0, 40) for _ in range(n)], np.int32))
        for algo in (umigpu.ALGO_DIR, umigpu.ALGO_CC):
            ctr = check_against_oracle(d, algo, umigpu.MERGE_AVGQUAL, rng.choice([1, 1, 2]), rng.choice([0.5, 0.9]), labels=True)
            total_sweeps += ctr["n_sweeps"]
    assert total_sweeps > 0
    # a long chain of frequency-1 UMIs (each one substitution from the next): the worst case for plain sweeps
    L = 12
    chain = ["A" * L]
    for i in range(1, 400):
        u = list(chain[-1]); u[i % L] = "ACGT"[("ACGT".index(u[i % L]) + 1) % 4]; chain.append("".join(u))
    n = len(chain)
    d = dict(tid=np.zeros(n, np.int32), pos=np.zeros(n, np.int64), rev=np.zeros(n, np.uint8), umi=arr(chain), score=np.full(n, 30, np.int32))
    for algo in (umigpu.ALGO_DIR, umigpu.ALGO_CC):
        check_against_oracle(d, algo, umigpu.MERGE_AVGQUAL, 1, 0.5, labels=True)


def test_full_size_two_phase_equals_plain_sweeps(monkeypatch):
    """C5 at full size (200 M reads, hottest locus 5.1 M unique UMIs, 3e7 edges, 40 plain sweeps): the two-phase
    clustering (engaged after 8 sweeps through the test knob; production keeps sweeping up to 64) and plain label sweeps
    run to their fixpoint must keep exactly the same reads."""
    d, cfg = synth.generate_config("C5", device="cuda", scale=1.0)
    monkeypatch.setenv("UMIGPU_PLAIN_ROUNDS", "2")
    kept, _, ctr = _run_flags(d, cfg, umigpu.ALGO_DIR, 0)
    monkeypatch.setenv("UMIGPU_SV_MIN_EDGES", str(1 << 62))
    pkept, _, pctr = _run_flags(d, cfg, umigpu.ALGO_DIR, 0)
    assert ctr["n_edges"] == pctr["n_edges"] > (8 << 20)
    assert np.array_equal(kept, pkept) and ctr["n_kept"] == pctr["n_kept"]
    assert pctr["n_sweeps"] != ctr["n_sweeps"]        # really two different schedules


def test_many_contigs_use_the_linear_coordinate_layout(monkeypatch):
    """Thousands of contigs with positions up to 2^27 and 13-nt UMIs: [tid | pos | strand | UMI] needs 12 + 27 + 1 + 26 = 66
    bits (two-word keys); laying the contigs' occupied ranges end to end keeps the key in one word.  Same survivors,
    same counters either way, and both match the oracle."""
    rng = np.random.default_rng(17)
    n, L, n_tid = 60000, 13, 3000
    tid = rng.integers(0, n_tid, n).astype(np.int32)
    # every contig is occupied over a narrow window somewhere below 2^27 (plus a negative corner case on contig 0)
    lo = rng.integers(0, (1 << 27) - 5000, n_tid)
    pos = (lo[tid] + rng.integers(0, 40, n) * 97).astype(np.int64)
    pos[tid == 0] -= (1 << 20)
    tid[:2] = (0, n_tid - 1); pos[0] = -(1 << 20); pos[1] = (1 << 27) - 1
    rev = rng.integers(0, 2, n).astype(np.uint8)
    fam = rng.integers(0, 4, (200, L))
    umi = np.frombuffer(b"ACGT", np.uint8)[np.where(rng.random((n, L)) < 0.05, rng.integers(0, 4, (n, L)), fam[rng.integers(0, 200, n)])]
    score = rng.integers(0, 41, n).astype(np.int32)
    d = dict(tid=tid, pos=pos, rev=rev, umi=umi, score=score)
    ctr_lin = check_against_oracle(d, umigpu.ALGO_DIR, umigpu.MERGE_AVGQUAL, 1, 0.5, labels=True)
    assert ctr_lin["key_bits"] <= 64
    monkeypatch.setenv("UMIGPU_NO_LINEAR_KEYS", "1")
    ctr_plain = check_against_oracle(d, umigpu.ALGO_DIR, umigpu.MERGE_AVGQUAL, 1, 0.5, labels=True)
    assert ctr_plain["key_bits"] > 64
    for key in ("n_buckets", "total_umis", "max_umis", "n_kept", "n_edges"):
        assert ctr_lin[key] == ctr_plain[key]


@pytest.mark.parametrize("L,alphabet", [(8, "ACGT"), (12, "ACGTN"), (16, "ACGT"), (17, "ACGTN"), (20, "ACGT"), (32, "ACGT")])
def test_compact_host_format_equals_ascii_push(L, alphabet):
    """umigpu_push_reads_packed (int32 positions, 2-bit UMIs, uint8 scores: 14-18 B/read over PCIe) must give exactly what
    the ASCII entry gives, and so the oracle's answer."""
    rng = random.Random(L)
    n = 5000
    pool = ["".join(rng.choice(alphabet) for _ in range(L)) for _ in range(60)]
    umis = []
    for _ in range(n):
        u = list(rng.choice(pool))
        if rng.random() < 0.3:
            u[rng.randrange(L)] = rng.choice(alphabet)
        umis.append("".join(u))
    d = dict(tid=np.array([rng.randrange(3) for _ in range(n)], np.int32), pos=np.array([rng.randrange(9) * 13 - 20 for _ in range(n)], np.int64),
             rev=np.array([rng.randrange(2) for _ in range(n)], np.uint8), umi=arr(umis), score=np.array([rng.randrange(0, 94) for _ in range(n)], np.int32))
    ctr = check_against_oracle(d, umigpu.ALGO_DIR, umigpu.MERGE_AVGQUAL, 1, 0.5)
    kept, _, _ = gpu_dedup(d["tid"], d["pos"], d["rev"], d["umi"], d["score"], umigpu.ALGO_DIR, umigpu.MERGE_AVGQUAL, 1, 0.5)
    code, nm = umigpu.pack_umis(d["umi"])
    assert (nm is not None) == ("N" in alphabet)
    with umigpu.Context(L, 1, 0.5, umigpu.ALGO_DIR, umigpu.MERGE_AVGQUAL) as ctx:
        half = n // 2 + 1
        ctx.push_reads_packed(d["tid"][:half], d["pos"][:half].astype(np.int32), d["rev"][:half], code[:half], None if nm is None else nm[:half],
                              d["score"][:half].astype(np.uint8), 0)
        ctx.push_reads_packed(d["tid"][half:], d["pos"][half:].astype(np.int32), d["rev"][half:], code[half:], None if nm is None else nm[half:],
                              d["score"][half:].astype(np.uint8), half)
        pkept, _, pctr = ctx.finish()
    assert np.array_equal(kept, pkept)
    for key in ("total_reads", "n_buckets", "total_umis", "max_umis", "n_kept", "n_edges"):
        assert ctr[key] == pctr[key]
    if 2 * L < code.dtype.itemsize * 8:      # a code with bits beyond 2L is rejected like an unknown UMI byte
        bad = code.copy(); bad[7] |= bad.dtype.type(1) << bad.dtype.type(2 * L)
        with umigpu.Context(L) as ctx:
            ctx.push_reads_packed(d["tid"], d["pos"].astype(np.int32), d["rev"], bad, nm, d["score"].astype(np.uint8), 0)
            with pytest.raises(umigpu.UmiGpuError, match="Unknown character"):
                ctx.finish()


def test_read_order_does_not_change_the_surviving_groups():
    """Size-independent property (no oracle needed): the (bucket, UMI) groups that survive depend on frequencies and UMIs
    only, never on the order in which the reads arrive; shuffling 2.5 M reads must keep exactly the same groups."""
    d, cfg = small("C2", 0.05, seed=9)
    n = len(d["tid"])
    perm = np.random.default_rng(1).permutation(n)
    def groups(dd):
        kept, _, ctr = gpu_dedup(dd["tid"], dd["pos"], dd["rev"], dd["umi"], dd["score"], umigpu.ALGO_DIR, umigpu.MERGE_AVGQUAL, 1, 0.5)
        k = kept.astype(np.int64)
        key = np.concatenate([dd["tid"][k, None].astype(np.int64), dd["pos"][k, None], dd["rev"][k, None].astype(np.int64), dd["umi"][k].astype(np.int64)], axis=1)
        return {tuple(r) for r in key.tolist()}, ctr
    g0, c0 = groups(d)
    g1, c1 = groups({kk: v[perm] for kk, v in d.items()})
    assert g0 == g1 and len(g0) == c0["n_kept"] == c1["n_kept"]
    for key in ("n_buckets", "total_umis", "max_umis", "n_edges"):
        assert c0[key] == c1[key]


def test_n_bases_in_long_umis():
    """N is a fifth letter at any UMI length the path supports (BitSet + n_bits in the reference, utils/bitset.rs:9-14): with
    N the sort code takes 3 bits per base, beyond 21 nt it spills into the second key word (dual 12-nt UMIs = 24 nt, 32 nt)."""
    rng = random.Random(5)
    for L in (22, 24, 27, 32):
        for trial in range(3):
            n = rng.choice([400, 6000])
            pool = ["".join(rng.choice("ACGT") for _ in range(L)) for _ in range(rng.choice([5, 60, 700]))]
            umis = []
            for _ in range(n):
                u = list(rng.choice(pool))
                for _ in range(rng.choice([0, 0, 1, 2])):
                    u[rng.randrange(L)] = rng.choice("ACGTN")
                umis.append("".join(u))
            d = dict(tid=np.zeros(n, np.int32), pos=np.array(sorted(rng.randrange(4) * 100 for _ in range(n)), np.int64),
                     rev=np.array([rng.randrange(2) for _ in range(n)], np.uint8), umi=arr(umis),
                     score=np.array([rng.randrange(0, 40) for _ in range(n)], np.int32))
            assert (d["umi"] == ord("N")).any()
            for algo in (umigpu.ALGO_DIR, umigpu.ALGO_CC, umigpu.ALGO_ADJ_UPSTREAM):
                check_against_oracle(d, algo, umigpu.MERGE_AVGQUAL, rng.choice([1, 2]), 0.5, labels=True)
    # a single N read in a large N-free batch of dual 12-nt UMIs (the case the CLI twin used to die on)
    d, cfg = small("C2", 0.002, seed=3, umi_len=24)
    d["umi"][len(d["umi"]) // 2, 7] = ord("N")
    check_against_oracle(d, umigpu.ALGO_DIR, umigpu.MERGE_AVGQUAL, 1, 0.5)


@pytest.mark.parametrize("algo", [umigpu.ALGO_CC, umigpu.ALGO_ADJ_UPSTREAM])
def test_full_size_c3_production_equals_brute_force(algo):
    """C3 at FULL size (50 M reads, 16-nt UMIs, k = 2: three multi-index passes, hamming_blocks<16,2,0>) for both algorithms
    BASELINE.json names (cc and the upstream-intended adjacency): the production search against the all-pairs direct kernel
    (3.6e10 pairs, nothing culled) — same edges, same survivors."""
    d, cfg = synth.generate_config("C3", device="cuda", scale=1.0)
    kept, _, ctr = _run_flags(d, cfg, algo, 0)
    bkept, _, bctr = _run_flags(d, cfg, algo, umigpu.FLAG_KERNEL_DIRECT | umigpu.FLAG_NO_CULL | umigpu.FLAG_NO_MULTI_INDEX)
    assert bctr["pairs_evaluated"] >= bctr["unordered_pairs"] > 1e10
    assert ctr["pairs_evaluated"] < bctr["pairs_evaluated"]
    assert ctr["n_edges"] == bctr["n_edges"] > 0 and ctr["n_kept"] == bctr["n_kept"]
    assert np.array_equal(kept, bkept)
    assert (np.diff(kept.astype(np.int64)) > 0).all()


def test_full_size_c5_production_equals_brute_force():
    """C5 at FULL size (200 M reads, hottest locus 5.1 M unique UMIs, 2.1e13 unordered pairs): production (segmented sort,
    multi-index passes, culling, two-phase clustering) against the dense bit-sliced tile kernel with culling and multi-index
    switched off and against the generic sort: same edge count, same survivors."""
    d, cfg = synth.generate_config("C5", device="cuda", scale=1.0)
    kept, _, ctr = _run_flags(d, cfg, umigpu.ALGO_DIR, 0)
    bkept, _, bctr = _run_flags(d, cfg, umigpu.ALGO_DIR, umigpu.FLAG_KERNEL_TILES | umigpu.FLAG_NO_CULL | umigpu.FLAG_NO_MULTI_INDEX)
    assert bctr["pairs_evaluated"] >= bctr["unordered_pairs"] > 1e13
    assert ctr["n_edges"] == bctr["n_edges"] > 1e7 and ctr["n_kept"] == bctr["n_kept"]
    assert np.array_equal(kept, bkept)
    for key in ("total_reads", "n_buckets", "total_umis", "max_umis"):
        assert ctr[key] == bctr[key]
