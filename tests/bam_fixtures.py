"""Synthetic BAM records for the host-feed tests (shared by CPU and GPU tests)."""
import random

from umigpu import bamio


def random_cigar(rng, read_len):
    """A valid CIGAR consuming read_len query bases: optional hard/soft clips at both ends, M/I/D/N/=/X inside."""
    ops = []
    lead_h = rng.choice([0, 0, 0, 5]); lead_s = rng.choice([0, 0, 3, 7])
    trail_s = rng.choice([0, 0, 2, 9]); trail_h = rng.choice([0, 0, 0, 4])
    inner = read_len - lead_s - trail_s
    if lead_h: ops.append((5, lead_h))
    if lead_s: ops.append((4, lead_s))
    left = inner
    first = True
    while left > 0:
        op = rng.choice([0, 0, 0, 7, 8]) if first else rng.choice([0, 0, 1, 2, 3, 7, 8])
        first = False
        if op in (2, 3):
            ops.append((op, rng.randint(1, 50)))
            continue
        ln = min(left, rng.randint(1, 40))
        ops.append((op, ln)); left -= ln
    if ops[-1][0] in (2, 3):
        ops.pop()
    if trail_s: ops.append((4, trail_s))
    if trail_h: ops.append((5, trail_h))
    return ops


def make_bam(rng, n, umi_len=8, n_pos=30, n_tid=3, unmapped_rate=0.05, alphabet="ACGT", extra_suffix=True):
    """Returns (header bytes, list of record bytes, list of dict with the ground-truth fields)."""
    header = bamio.make_header([f"chr{t}" for t in range(n_tid)], [10_000_000] * n_tid)
    recs, truth = [], []
    pools = {}
    for i in range(n):
        tid = rng.randrange(n_tid)
        pos = rng.randrange(n_pos) * 11 + 100
        rev = rng.random() < 0.4
        key = (tid, pos, rev)
        fam = pools.setdefault(key, ["".join(rng.choice(alphabet) for _ in range(umi_len)) for _ in range(rng.randint(1, 5))])
        u = list(rng.choice(fam))
        if rng.random() < 0.2:
            u[rng.randrange(umi_len)] = rng.choice(alphabet)
        umi = "".join(u)
        unmapped = rng.random() < unmapped_rate
        read_len = rng.choice([20, 36, 50, 75])
        cigar = [] if unmapped else random_cigar(rng, read_len)
        flag = (4 if unmapped else 0) | (16 if rev else 0)
        qual = bytes(rng.randrange(0, 42) for _ in range(read_len))
        mapq = rng.randrange(0, 61)
        qname = f"read{i}_{umi}".encode() + (b"_x1" if extra_suffix and i % 3 == 0 else b"")
        recs.append(bamio.make_record(-1 if unmapped else tid, -1 if unmapped else pos, flag, mapq, qname, cigar, read_len, qual))
        truth.append(dict(tid=tid, pos=pos, rev=rev, umi=umi, unmapped=unmapped, cigar=cigar, qual=qual, mapq=mapq))
    return header, recs, truth


def make_paired_bam(rng, n_templates, umi_len=8, n_pos=12, n_tid=3, alphabet="ACGT"):
    """Paired-end fixture for deduplicate_sam.rs:96-129 and UcWriter::write_reversed (:409-462): every template has a
    first read and (mostly) a mate with the same name; a few reads are unpaired, unmapped, have an unmapped mate or a
    mate on another reference (chimeric); template lengths come from a small pool so that (tid, pos, strand) buckets
    split by tlen.  Records are shuffled (coordinate order is not required by the path)."""
    header = bamio.make_header([f"chr{t}" for t in range(n_tid)], [10_000_000] * n_tid)
    recs = []
    pools = {}
    for i in range(n_templates):
        tid = rng.randrange(n_tid)
        pos = rng.randrange(n_pos) * 11 + 100
        rev = rng.random() < 0.4
        tl = rng.choice([150, 151, 200, -150, 320])
        key = (tid, pos, rev)
        fam = pools.setdefault(key, ["".join(rng.choice(alphabet) for _ in range(umi_len)) for _ in range(rng.randint(1, 4))])
        u = list(rng.choice(fam))
        if rng.random() < 0.2:
            u[rng.randrange(umi_len)] = rng.choice(alphabet)
        qname = f"t{i}_{''.join(u)}".encode()
        kind = rng.random()
        read_len = rng.choice([20, 36, 50])
        q1 = bytes(rng.randrange(0, 42) for _ in range(read_len)); q2 = bytes(rng.randrange(0, 42) for _ in range(read_len))
        mq = rng.randrange(0, 61)
        mpos = pos + abs(tl) - read_len
        sflag = 16 if rev else 0
        mflag = 0 if rev else 16          # mate on the other strand
        if kind < 0.08:                   # single-end read inside a paired file
            recs.append(bamio.make_record(tid, pos, sflag, mq, qname, random_cigar(rng, read_len), read_len, q1))
        elif kind < 0.14:                 # first read unmapped (mate mapped)
            recs.append(bamio.make_record(tid, mpos, 0x1 | 0x40 | 0x4 | (0x20 if mflag else 0), 0, qname, [], read_len, q1, tid, mpos, 0))
            recs.append(bamio.make_record(tid, mpos, 0x1 | 0x80 | 0x8 | mflag, mq, qname, random_cigar(rng, read_len), read_len, q2, tid, mpos, 0))
        elif kind < 0.20:                 # mate unmapped
            recs.append(bamio.make_record(tid, pos, 0x1 | 0x40 | 0x8 | sflag, mq, qname, random_cigar(rng, read_len), read_len, q1, tid, pos, 0))
            recs.append(bamio.make_record(tid, pos, 0x1 | 0x80 | 0x4, 0, qname, [], read_len, q2, tid, pos, 0))
        elif kind < 0.30:                 # chimeric pair
            mt = (tid + 1) % n_tid
            recs.append(bamio.make_record(tid, pos, 0x1 | 0x40 | sflag | (0x20 if mflag else 0), mq, qname, random_cigar(rng, read_len), read_len, q1, mt, mpos, 0))
            recs.append(bamio.make_record(mt, mpos, 0x1 | 0x80 | mflag | (0x20 if sflag else 0), mq, qname, random_cigar(rng, read_len), read_len, q2, tid, pos, 0))
        else:                             # proper pair
            recs.append(bamio.make_record(tid, pos, 0x1 | 0x2 | 0x40 | sflag | (0x20 if mflag else 0), mq, qname, random_cigar(rng, read_len), read_len, q1, tid, mpos, tl))
            recs.append(bamio.make_record(tid, mpos, 0x1 | 0x2 | 0x80 | mflag | (0x20 if sflag else 0), mq, qname, random_cigar(rng, read_len), read_len, q2, tid, pos, -tl))
            if rng.random() < 0.05:       # a secondary copy of the mate: only the first one in file order is written (:455-458)
                recs.append(recs[-1])
    rng.shuffle(recs)
    return header, recs
