"""Generates tests/golden/cases.json from the pure-Python literal restatement (oracle/ref_literal.py).

The reference ships no golden vectors and cannot be built here (no cargo), so these fixtures come
from the line-by-line Python restatement, NOT from the reference binary: parity stays "unpinned"
(see DESIGN.md).  They pin the C oracle and the CUDA path against an independently written
implementation and against regressions.  Run:  python tests/golden/make_golden.py
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import ref_literal as R  # noqa: E402


def rand_umi(rng, L, alphabet):
    return "".join(rng.choice(alphabet) for _ in range(L))


def make_reads_case(rng, name, n, L, n_pos, algo, merge, k, p, alphabet="ACGT", n_tid=1, neg=False, paired=False):
    pool = {}
    tid, pos, rev, umi, score = [], [], [], [], []
    tlen = [] if paired else None
    for _ in range(n):
        t = rng.randrange(n_tid)
        q = rng.randrange(n_pos) * 7 - (50 if neg else 0)
        r = rng.randrange(2)
        key = (t, q, r)
        fam = pool.setdefault(key, [rand_umi(rng, L, alphabet) for _ in range(rng.randint(1, 6))])
        u = list(rng.choice(fam))
        if rng.random() < 0.25:
            u[rng.randrange(L)] = rng.choice(alphabet)
        tid.append(t); pos.append(q); rev.append(r); umi.append("".join(u)); score.append(rng.randint(0, 41))
        if paired:
            tlen.append(rng.choice([-310, -150, 0, 150, 151, 310]))      # --paired: PairedAlignment's fourth field
    kept, ctr = R.dedup(tid, pos, rev, [u.encode() for u in umi], score, algo, merge, k, p, tlen=tlen)
    ctr.pop("dist_calls")
    case = dict(name=name, umi_len=L, algo=algo, merge=merge, k=k, p=p, tid=tid, pos=pos, rev=rev, umi=umi,
                score=score, kept=kept, counters=ctr)
    if paired:
        case["tlen"] = tlen
    return case


def make_bucket_case(rng, name, n, L, algo, k, p, alphabet="ACGT"):
    s = set()
    base = rand_umi(rng, L, alphabet)
    while len(s) < n:
        u = list(base if rng.random() < 0.7 else rand_umi(rng, L, alphabet))
        for _ in range(rng.randint(0, 3)):
            u[rng.randrange(L)] = rng.choice(alphabet)
        s.add("".join(u))
    umis = sorted(s)
    rng.shuffle(umis)
    freq = [rng.choice([1, 1, 1, 1, 2, 2, 3, 4, 7, 20, 100]) for _ in umis]
    keep, label, _ = R.cluster_bucket([u.encode() for u in umis], freq, algo, k, p)
    return dict(name=name, umi_len=L, algo=algo, k=k, p=p, umis=umis, freq=freq, keep=keep, label=label)


def main():
    rng = random.Random(20261018)
    reads_cases, bucket_cases = [], []
    i = 0
    for algo in (R.ALGO_DIR, R.ALGO_ADJ_REF, R.ALGO_ADJ_UPSTREAM, R.ALGO_CC):
        for merge in (R.MERGE_ANY, R.MERGE_AVGQUAL):
            for (L, k, alphabet, n_tid, neg) in ((6, 1, "ACGT", 1, False), (10, 1, "ACGTN", 3, True), (16, 2, "ACGT", 2, False)):
                reads_cases.append(make_reads_case(rng, f"reads{i}", 400, L, 12, algo, merge, k,
                                                   0.5 if i % 3 else 0.3, alphabet, n_tid, neg))
                i += 1
    # --paired cases from their own generator state, so that the cases above stay byte-identical
    prng = random.Random(20261019)
    for j, algo in enumerate((R.ALGO_DIR, R.ALGO_ADJ_REF, R.ALGO_ADJ_UPSTREAM, R.ALGO_CC)):
        reads_cases.append(make_reads_case(prng, f"paired{j}", 500, 8, 6, algo, R.MERGE_AVGQUAL, 1, 0.5, "ACGT", 2, j % 2 == 1, paired=True))
    i = 0
    for algo in (R.ALGO_DIR, R.ALGO_ADJ_REF, R.ALGO_ADJ_UPSTREAM, R.ALGO_CC):
        for (L, k, alphabet) in ((5, 1, "ACGT"), (8, 2, "ACGT"), (12, 1, "ACGTN"), (22, 3, "ACGT"), (32, 2, "ACGT"), (21, 1, "ACGTN")):
            bucket_cases.append(make_bucket_case(rng, f"bucket{i}", 60, L, algo, k, [0.5, 0.5, 0.25, 0.8][i % 4], alphabet))
            i += 1
    with open(os.path.join(HERE, "cases.json"), "w") as f:
        json.dump(dict(reads=reads_cases, buckets=bucket_cases), f, separators=(",", ":"))
    print(len(reads_cases), "read cases,", len(bucket_cases), "bucket cases")


if __name__ == "__main__":
    main()
