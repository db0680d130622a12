"""CPU tests: the oracle against golden fixtures, known answers, an independent literal restatement and
the order-independent invariants of the reference.  No GPU, no libumigpu compute."""
import json
import os
import random

import numpy as np
import pytest

import oracle_lib as O
import ref_literal as R

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "cases.json")))


def arr(umis):
    return np.frombuffer("".join(umis).encode(), dtype=np.uint8).reshape(len(umis), len(umis[0]))


def plain_hamming(a: bytes, b: bytes) -> int:
    return sum(x != y for x, y in zip(a, b))


@pytest.mark.parametrize("L", [1, 4, 10, 12, 16, 21, 22, 23, 30, 32, 42, 43, 44])
def test_reference_distance_formula_is_plain_hamming(L):
    """bitset.rs:77-91 + mod.rs:24-26 == Hamming over {A,C,G,T,N}, including bases straddling 64-bit words."""
    rng = random.Random(L)
    for _ in range(300):
        a = "".join(rng.choice("ACGTN") for _ in range(L)).encode()
        b = bytearray(a)
        for _ in range(rng.randint(0, 4)):
            b[rng.randrange(L)] = ord(rng.choice("ACGTN"))
        b = bytes(b)
        assert O.umi_dist(a, b) == plain_hamming(a, b)
        assert R.umi_dist(R.to_bitset(a), R.to_bitset(b)) == plain_hamming(a, b)


def test_unknown_base_is_an_error():
    assert O.lib().oracle_umi_dist(b"ACGX", b"ACGT", 4) == -1      # reference panics, utils/mod.rs:78
    with pytest.raises(ValueError):
        R.to_bitset(b"acgt")


def test_threshold_f32_semantics():
    """directional.rs:38; with p = 0.5 the rule is freq_u >= 2 freq_v - 1 for every freq < 2^24."""
    rng = random.Random(7)
    for f in list(range(0, 300)) + [rng.randrange(1, 1 << 24) for _ in range(2000)]:
        t = O.lib().oracle_dir_threshold(0.5, f)
        assert t == (f + 1) // 2 == R.dir_threshold(0.5, f)
    for p in (0.1, 0.3, 0.75, 0.9, 1.0):
        for f in (1, 2, 3, 10, 99, 1000, 123456):
            assert O.lib().oracle_dir_threshold(p, f) == R.dir_threshold(p, f)


def test_known_answer_umi_tools_network():
    """Hand-derived: the classic directional example (ACGT 456 ... AAAT 90)."""
    umis = ["ACGT", "TCGT", "CCGT", "ACAT", "ACAG", "AAAT"]
    freq = [456, 2, 2, 72, 1, 90]
    a = arr(umis)
    keep, label, _ = O.cluster_bucket(a, freq, O.ALGO_DIR, 1, 0.5)
    assert keep.tolist() == [1, 0, 0, 0, 0, 1]
    assert label.tolist() == [0, 0, 0, 0, 0, 5]
    keep, label, _ = O.cluster_bucket(a, freq, O.ALGO_CC, 1, 0.5)
    assert keep.tolist() == [1, 0, 0, 0, 0, 0] and set(label.tolist()) == {0}
    keep, _, _ = O.cluster_bucket(a, freq, O.ALGO_ADJ_UPSTREAM, 1, 0.5)
    assert keep.tolist() == [1, 0, 0, 0, 1, 1]
    keep, _, _ = O.cluster_bucket(a, freq, O.ALGO_ADJ_REF, 1, 0.5)      # nothing but the query is removed
    assert keep.tolist() == [1] * 6


def test_known_answer_singleton_chain_and_ties():
    """freq-1 UMIs absorb each other (1 >= 2*1-1); the canonical tie-break picks the smallest UMI."""
    umis = ["AAAA", "AAAC", "AACC", "ACCC", "TTTT"]
    keep, label, _ = O.cluster_bucket(arr(umis), [1] * 5, O.ALGO_DIR, 1, 0.5)
    assert keep.tolist() == [1, 0, 0, 0, 1] and label.tolist() == [0, 0, 0, 0, 4]
    # a freq-2 UMI does not absorb a freq-2 neighbour (2 >= 3 is false) but absorbs freq-1 ones
    keep, _, _ = O.cluster_bucket(arr(["AAAA", "AAAC", "AAAG"]), [2, 2, 1], O.ALGO_DIR, 1, 0.5)
    assert keep.tolist() == [1, 1, 0]


@pytest.mark.parametrize("case", GOLD["buckets"], ids=lambda c: c["name"])
def test_golden_buckets(case):
    keep, label, _ = O.cluster_bucket(arr(case["umis"]), case["freq"], case["algo"], case["k"], case["p"])
    assert keep.tolist() == case["keep"]
    assert label.tolist() == case["label"]


@pytest.mark.parametrize("case", GOLD["reads"], ids=lambda c: c["name"])
def test_golden_reads(case):
    kept, _, ctr = O.dedup(case["tid"], case["pos"], case["rev"], arr(case["umi"]), case["score"], case["algo"],
                           case["merge"], case["k"], case["p"], tlen=case.get("tlen"))
    assert kept.tolist() == case["kept"]
    for key, v in case["counters"].items():
        assert ctr[key] == v, key


def test_c_oracle_matches_literal_python_on_random_buckets():
    rng = random.Random(11)
    for trial in range(150):
        L = rng.choice([4, 5, 6, 8, 21, 22, 23])
        n = rng.randint(1, 50)
        alpha = "ACGT" if trial % 3 else "ACGTN"
        s = set()
        n = min(n, 2 ** L)
        while len(s) < n:
            s.add("".join(rng.choice(alpha[: rng.choice([2, 4, len(alpha)])]) for _ in range(L)))
        umis = sorted(s)
        rng.shuffle(umis)
        freq = [rng.choice([1, 1, 1, 2, 3, 5, 10, 40]) for _ in umis]
        for algo in range(4):
            k, p = rng.choice([0, 1, 2]), rng.choice([0.5, 0.5, 0.3, 0.9])
            a = R.cluster_bucket([u.encode() for u in umis], freq, algo, k, p)
            b = O.cluster_bucket(arr(umis), freq, algo, k, p)
            assert a[0] == b[0].tolist() and a[1] == b[1].tolist() and a[2] == b[2]


def _random_reads(rng, n, L, n_pos):
    tid = [rng.randrange(2) for _ in range(n)]
    pos = [rng.randrange(n_pos) - 3 for _ in range(n)]
    rev = [rng.randrange(2) for _ in range(n)]
    umi = ["".join(rng.choice("ACG") for _ in range(L)) for _ in range(n)]
    score = [rng.randrange(0, 42) for _ in range(n)]
    return tid, pos, rev, umi, score


def test_invariants_of_the_reference():
    """properties that hold for every hash order of the reference."""
    rng = random.Random(5)
    tid, pos, rev, umi, score = _random_reads(rng, 3000, 5, 6)
    a = arr(umi)
    # adj as written keeps exactly one read per unique (bucket, UMI)
    kept, _, ctr = O.dedup(tid, pos, rev, a, score, O.ALGO_ADJ_REF, O.MERGE_AVGQUAL, 1, 0.5)
    uniq = {}
    for i in range(len(umi)):
        uniq.setdefault((tid[i], pos[i], rev[i], umi[i]), []).append(i)
    assert ctr["n_kept"] == len(uniq) == ctr["total_umis"]
    assert ctr["n_buckets"] == len({(t, p, r) for t, p, r in zip(tid, pos, rev)})
    # representative = first read in input order attaining the max score (deduplicate_sam.rs:165-174)
    expect = sorted(min(i for i in idx if score[i] == max(score[j] for j in idx)) for idx in uniq.values())
    assert kept.tolist() == expect
    # MERGE_ANY: first read of each unique
    kept_any, _, _ = O.dedup(tid, pos, rev, a, score, O.ALGO_ADJ_REF, O.MERGE_ANY, 1, 0.5)
    assert kept_any.tolist() == sorted(min(idx) for idx in uniq.values())
    # directional output count <= cc... and kept sets nest: cc roots are dir roots
    kd, _, cd = O.dedup(tid, pos, rev, a, score, O.ALGO_DIR, O.MERGE_AVGQUAL, 1, 0.5)
    kc, _, cc = O.dedup(tid, pos, rev, a, score, O.ALGO_CC, O.MERGE_AVGQUAL, 1, 0.5)
    assert cc["n_kept"] <= cd["n_kept"] <= ctr["n_kept"]
    assert set(kc.tolist()) <= set(kd.tolist()) <= set(kept.tolist())
    # k = 0 never merges distinct UMIs
    k0, _, c0 = O.dedup(tid, pos, rev, a, score, O.ALGO_DIR, O.MERGE_AVGQUAL, 0, 0.5)
    assert k0.tolist() == kept.tolist()


def test_min_label_propagation_equals_reference_dfs():
    """The formulation the CUDA clustering uses (cluster.cuh): keep v <=> no earlier-visited UMI reaches v over
    edges u->v (dist<=k, freq_v <= thr_u); root = earliest UMI reaching v.  Checked against the DFS oracle."""
    rng = random.Random(3)
    for trial in range(300):
        L = rng.choice([4, 5, 6])
        alpha = "AC" if trial % 2 else "ACGT"
        n = rng.randint(1, min(40, len(alpha) ** L))
        s = set()
        while len(s) < n:
            s.add("".join(rng.choice(alpha) for _ in range(L)))
        umis = sorted(s)
        freq = [rng.choice([1, 1, 1, 2, 2, 3, 5, 9, 30]) for _ in umis]
        k, p = rng.choice([1, 2]), rng.choice([0.5, 0.5, 0.3, 0.8])
        for algo in (O.ALGO_DIR, O.ALGO_CC):
            keep, label, _ = O.cluster_bucket(arr(umis), freq, algo, k, p)
            order = sorted(range(n), key=lambda i: (-freq[i], umis[i].translate(str.maketrans("ACGTN", "01234"))))
            rank = {u: r for r, u in enumerate(order)}
            thr = [2**31 - 1 if algo == O.ALGO_CC else R.dir_threshold(p, f) for f in freq]
            lab = [rank[i] for i in range(n)]
            changed = True
            while changed:
                changed = False
                for u in range(n):
                    for v in range(n):
                        if u != v and plain_hamming(umis[u].encode(), umis[v].encode()) <= k and freq[v] <= thr[u] and lab[u] < lab[v]:
                            lab[v] = lab[u]; changed = True
            assert [int(lab[i] == rank[i]) for i in range(n)] == keep.tolist()
            assert [order[lab[i]] for i in range(n)] == label.tolist()


def test_avg_qual_and_unclipped_pos():
    rng = random.Random(9)
    for _ in range(200):
        q = np.array([rng.randrange(0, 94) for _ in range(rng.randint(1, 400))], dtype=np.uint8)
        assert O.avg_qual(q) == int(q.sum()) // len(q) == R.avg_qual(bytes(q))
    assert O.avg_qual(np.zeros(0, np.uint8)) == 0
    # utils/mod.rs:96-104 examples from its doc comment: start 100, 4 clipped -> 96; end 100, 7 clipped -> 107
    assert O.unclipped_pos(100, 0, [(4, 4), (0, 50)]) == 96
    assert O.unclipped_pos(100, 0, [(5, 3), (4, 1), (0, 50)]) == 96
    assert O.unclipped_pos(51, 1, [(0, 50), (4, 7)]) == 51 + 50 - 1 + 7
    assert O.unclipped_pos(51, 1, [(0, 20), (2, 5), (1, 3), (0, 25), (4, 2), (5, 5)]) == 51 + 50 - 1 + 7
    for cig in ([(0, 10)], [(4, 2), (0, 10), (4, 3)], [(5, 1), (0, 5), (3, 100), (0, 5), (5, 2)]):
        for rev in (0, 1):
            assert O.unclipped_pos(1000, rev, cig) == R.unclipped_pos(1000, bool(rev), cig)


def test_paired_key_c_oracle_matches_literal():
    """--paired: the bucket key gains the template length (PairedAlignment, deduplicate_sam.rs:545-565)."""
    rng = random.Random(21)
    for trial in range(20):
        tid, pos, rev, umi, score = _random_reads(rng, 300, 5, 6)
        tlen = [rng.choice([-200, 0, 150, 151]) for _ in tid]
        for algo in range(4):
            a_kept, a_ctr = R.dedup(tid, pos, rev, [u.encode() for u in umi], score, algo, O.MERGE_AVGQUAL, 1, 0.5, tlen=tlen)
            b_kept, _, b_ctr = O.dedup(tid, pos, rev, arr(umi), score, algo, O.MERGE_AVGQUAL, 1, 0.5, tlen=tlen)
            assert a_kept == b_kept.tolist()
            assert a_ctr["n_buckets"] == b_ctr["n_buckets"] and a_ctr["total_umis"] == b_ctr["total_umis"]
        u_kept, _, u_ctr = O.dedup(tid, pos, rev, arr(umi), score, 0, O.MERGE_AVGQUAL, 1, 0.5)
        assert u_ctr["n_buckets"] < b_ctr["n_buckets"]


def test_all_core_variant_equals_single_thread():
    """oracle_set_threads(n): buckets are independent, so handing them to n threads changes nothing but the time."""
    rng = random.Random(31)
    tid, pos, rev, umi, score = _random_reads(rng, 4000, 6, 40)
    a = O.dedup(tid, pos, rev, arr(umi), score, O.ALGO_DIR, O.MERGE_AVGQUAL, 1, 0.5, want_roots=True)
    O.set_threads(4)
    try:
        b = O.dedup(tid, pos, rev, arr(umi), score, O.ALGO_DIR, O.MERGE_AVGQUAL, 1, 0.5, want_roots=True)
    finally:
        O.set_threads(1)
    assert a[0].tolist() == b[0].tolist() and a[1].tolist() == b[1].tolist() and a[2] == b[2]


def test_hypothesis_properties_of_the_path():
    """Property tests (hypothesis) of the oracle, i.e. of the reference's algorithm under the canonical tie-break:
    (1) the kept (bucket, UMI) groups do not depend on the order of the reads; (2) splitting the input by bucket and
    deduplicating the parts separately gives the union (buckets are independent — what the multi-GPU sharding relies on);
    (3) adding an exact duplicate of a read never changes which (bucket, UMI) groups survive for adj as written; for cc it
    never changes HOW MANY survive (the components are the same; the extra read can make another UMI of a component the most
    frequent one, i.e. its representative).  derandomize: the same examples on every run."""
    from hypothesis import given, settings, strategies as st

    read = st.tuples(st.integers(0, 1), st.integers(-2, 3), st.integers(0, 1), st.text("ACG", min_size=4, max_size=4), st.integers(0, 40))

    def groups(reads, algo):
        tid, pos, rev, umi, score = (list(x) for x in zip(*reads))
        kept, _, _ = O.dedup(tid, pos, rev, arr(umi), score, algo, O.MERGE_AVGQUAL, 1, 0.5)
        return {(tid[i], pos[i], rev[i], umi[i]) for i in kept.tolist()}, kept.tolist()

    @settings(max_examples=120, deadline=None, derandomize=True)
    @given(st.lists(read, min_size=1, max_size=60), st.randoms(use_true_random=False), st.sampled_from([O.ALGO_DIR, O.ALGO_CC, O.ALGO_ADJ_REF, O.ALGO_ADJ_UPSTREAM]))
    def check(reads, rnd, algo):
        g, _ = groups(reads, algo)
        shuffled = list(reads); rnd.shuffle(shuffled)
        assert groups(shuffled, algo)[0] == g                                  # (1)
        even = [r for r in reads if r[1] % 2 == 0]; odd = [r for r in reads if r[1] % 2 != 0]
        parts = set()
        for part in (even, odd):
            if part:
                parts |= groups(part, algo)[0]
        assert parts == g                                                      # (2)
        if algo == O.ALGO_ADJ_REF:
            assert groups(reads + [reads[0]], algo)[0] == g                     # (3)
        elif algo == O.ALGO_CC:
            assert len(groups(reads + [reads[0]], algo)[0]) == len(g)           # (3)

    check()
