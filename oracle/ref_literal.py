"""Pure-Python *literal* restatement of the umi-collapse-rs hot path.

TEST INFRASTRUCTURE ONLY (see oracle/umi_oracle.c header).  PARITY UNPINNED: the reference has
no tests or golden vectors and cannot be built here (no cargo); this file is a second,
independently written restatement that follows the Rust line by line (HashMap -> dict,
retain -> dict rebuild, recursion kept as recursion).  It is slow on purpose and is used to
(a) cross-check oracle/umi_oracle.c and (b) generate the fixtures under tests/golden/.

Canonicalisation of the reference's RandomState order: dictionaries are
filled in ascending UMI order with A < C < G < T < N, so the stable freq-descending sort breaks
ties in that order; survivors are reported in input order.

All citations are relative to /root/reference.
"""
from __future__ import annotations

import struct
import sys

ENCODING_DIST = 2          # src/utils/read.rs:13
ENCODING_LENGTH = 3        # src/utils/read.rs:14
ENCODING_MAP = {ord("A"): 0b000, ord("T"): 0b101, ord("C"): 0b110, ord("G"): 0b011, ord("N"): 0b100}  # read.rs:22-31
CHUNK_SIZE = 64            # src/utils/bitset.rs:6
_M64 = (1 << 64) - 1
_CANON = {ord("A"): 0, ord("C"): 1, ord("G"): 2, ord("T"): 3, ord("N"): 4}

ALGO_DIR, ALGO_ADJ_REF, ALGO_ADJ_UPSTREAM, ALGO_CC = 0, 1, 2, 3
MERGE_ANY, MERGE_AVGQUAL, MERGE_MAPQUAL = 0, 1, 2
I32_MAX = 2**31 - 1


def canon_key(umi: bytes):
    return tuple(_CANON[c] for c in umi)


def _f32(x: float) -> float:
    return struct.unpack("f", struct.pack("f", x))[0]


class BitSet:
    """src/utils/bitset.rs:9-14 (bits as unsigned 64-bit words; sign is irrelevant to every use)."""

    def __init__(self, length: int):
        cap = length // CHUNK_SIZE + (0 if length % CHUNK_SIZE == 0 else 1)   # bitset.rs:17
        self.bits = [0] * cap
        self.n_bits = None

    def set(self, idx: int, bit: bool):                                       # bitset.rs:51-60
        i, j = divmod(idx, CHUNK_SIZE)
        self.bits[i] = (self.bits[i] | (1 << j)) if bit else (self.bits[i] & ~(1 << j) & _M64)

    def set_n_bit(self, idx: int, bit: bool):                                 # bitset.rs:62-75
        if self.n_bits is None:
            self.n_bits = [0] * len(self.bits)
        i, j = divmod(idx, CHUNK_SIZE)
        self.n_bits[i] = (self.n_bits[i] | (1 << j)) if bit else (self.n_bits[i] & ~(1 << j) & _M64)

    def bit_count_xor(self, o: "BitSet") -> int:                              # bitset.rs:77-91
        res = 0
        for i in range(len(self.bits)):
            a = self.n_bits[i] if self.n_bits is not None else 0
            b = o.n_bits[i] if o.n_bits is not None else 0
            x = a ^ b
            res += bin(x | (self.bits[i] ^ o.bits[i])).count("1") - bin(x).count("1") // ENCODING_LENGTH
        return res

    def key(self):                                                            # Eq on bits only, bitset.rs:94-101
        return tuple(self.bits)


def to_bitset(s: bytes) -> BitSet:                                            # src/utils/mod.rs:63-83
    res = BitSet(len(s) * ENCODING_LENGTH)
    for i, c in enumerate(s):
        if c not in ENCODING_MAP:
            raise ValueError(f"Unknown character in UMI sequence: {c}")      # mod.rs:78
        enc = ENCODING_MAP[c]
        for b in range(ENCODING_LENGTH):                                      # char_set mod.rs:38-43
            res.set(i * ENCODING_LENGTH + b, (enc & (1 << b)) != 0)
        if c == ord("N"):
            for b in range(ENCODING_LENGTH):                                  # char_set_n_bit mod.rs:45-50
                res.set_n_bit(i * ENCODING_LENGTH + b, True)
    return res


def umi_dist(a: BitSet, b: BitSet) -> int:                                    # src/utils/mod.rs:24-26
    return a.bit_count_xor(b) // ENCODING_DIST


def dir_threshold(p: float, freq: int) -> int:                                # src/algo/directional.rs:38
    t = _f32(_f32(p) * _f32(float(freq + 1)))
    if t != t:
        return 0
    return max(-(2**31), min(I32_MAX, int(t)))


def avg_qual(qual: bytes) -> int:                                             # src/utils/read.rs:56-63
    s = 0.0
    for b in qual:
        s = _f32(s + float(b))
    if len(qual) == 0:
        return 0
    return int(_f32(s / _f32(float(len(qual)))))


class Naive:
    """src/data/naive.rs:14-49"""

    def __init__(self, umi_freq: dict):
        self.umi_freq = dict(umi_freq)
        self.dist_calls = 0

    def remove_near(self, umi, bitsets, k: int, max_freq: int):               # naive.rs:26-40
        res, kept = [], {}
        for o, f in self.umi_freq.items():
            self.dist_calls += 1
            dist = umi_dist(bitsets[umi], bitsets[o])
            if dist <= k and (dist == 0 or f <= max_freq):
                res.append(o)
            else:
                kept[o] = f
        self.umi_freq = kept
        return res

    def contains(self, umi) -> bool:                                          # naive.rs:42-44
        return umi in self.umi_freq


def _visit_and_remove(start, freqs, bitsets, data, k, p, algo, label, root):  # directional.rs:30-54
    thr = I32_MAX if algo == ALGO_CC else dir_threshold(p, freqs[start])
    near = data.remove_near(start, bitsets, k, thr)
    for v in near:
        label[v] = root
    for v in near:
        if v == start:
            continue
        _visit_and_remove(v, freqs, bitsets, data, k, p, algo, label, root)


def cluster_bucket(umis: list[bytes], freq: list[int], algo: int, k: int, p: float):
    """Algorithm::apply (src/algo/mod.rs:13-20).  Returns (keep, label, dist_calls) indexed like input."""
    sys.setrecursionlimit(max(10000, 4 * len(umis) + 100))
    idx = sorted(range(len(umis)), key=lambda i: canon_key(umis[i]))          # canonical HashMap order
    bitsets = {i: to_bitset(umis[i]) for i in idx}
    freqs = {i: freq[i] for i in idx}
    data = Naive({i: freq[i] for i in idx})                                   # directional.rs:64-65,74
    order = sorted(idx, key=lambda i: -freq[i])                               # stable, directional.rs:67-72
    keep = [0] * len(umis)
    label = [-1] * len(umis)
    for u in order:                                                           # directional.rs:78 / adjacency.rs:47
        if not data.contains(u):
            continue
        if algo in (ALGO_ADJ_REF, ALGO_ADJ_UPSTREAM):
            mf = 0 if algo == ALGO_ADJ_REF else I32_MAX                       # adjacency.rs:56
            for v in data.remove_near(u, bitsets, k, mf):
                label[v] = u
        else:
            _visit_and_remove(u, freqs, bitsets, data, k, p, algo, label, u)
        keep[u] = 1                                                           # directional.rs:86
        label[u] = u
    return keep, label, data.dist_calls


def dedup(tid, pos, rev, umis: list[bytes], score, algo: int, merge: int, k: int, p: float, tlen=None):
    """deduplicate_sam.rs:93-233 on SoA input.  Returns (kept read indices ascending, counters dict).
    tlen given = --paired: the bucket key is PairedAlignment (deduplicate_sam.rs:133-139, :545-565)."""
    align: dict = {}
    for i in range(len(umis)):                                                # HOT LOOP A, :93
        key = (bool(rev[i]), int(pos[i]), int(tid[i]))                        # Alignment :485-489
        if tlen is not None:
            key = key + (int(tlen[i]),)                                       # PairedAlignment :547-552
        umi_reads = align.setdefault(key, {})                                 # :148-150
        ukey = to_bitset(umis[i]).key()                                       # :158
        if ukey not in umi_reads:                                             # Vacant :161-163
            umi_reads[ukey] = [i, 1, int(score[i]), umis[i]]
        else:                                                                 # Occupied :164-175
            e = umi_reads[ukey]
            keep_existing = True if merge == MERGE_ANY else e[2] >= int(score[i])   # merge/mod.rs
            e[1] += 1
            if not keep_existing:
                e[0], e[2] = i, int(score[i])
    kept = []
    ctr = dict(total_reads=len(umis), n_buckets=len(align), total_umis=0, max_umis=0, n_kept=0,
               dist_calls=0, unordered_pairs=0)
    for key, umi_reads in align.items():                                      # HOT LOOP B, :207
        ents = sorted(umi_reads.values(), key=lambda e: canon_key(e[3]))
        keep, _label, dc = cluster_bucket([e[3] for e in ents], [e[1] for e in ents], algo, k, p)
        ctr["total_umis"] += len(ents)                                        # :217
        ctr["max_umis"] = max(ctr["max_umis"], len(ents))                     # :218
        ctr["unordered_pairs"] += len(ents) * (len(ents) - 1) // 2
        ctr["dist_calls"] += dc
        for e, kf in zip(ents, keep):
            if kf:
                kept.append(e[0])
                ctr["n_kept"] += 1                                            # :219
    kept.sort()
    return kept, ctr


def unclipped_pos(pos: int, is_reverse: bool, cigar: list[tuple[int, int]]) -> int:
    """src/utils/mod.rs:96-104 with rust-htslib 0.49.0 CigarStringView semantics (BAM spec; unpinned).
    cigar = [(op, len)], op codes M0 I1 D2 N3 S4 H5 P6 =7 X8."""
    if not is_reverse:
        i, hard, soft = 0, 0, 0
        if i < len(cigar) and cigar[i][0] == 5:
            hard = cigar[i][1]; i += 1
        if i < len(cigar) and cigar[i][0] == 4:
            soft = cigar[i][1]
        return pos - soft - hard
    end = pos + sum(l for op, l in cigar if op in (0, 2, 3, 7, 8))
    i, hard, soft = len(cigar) - 1, 0, 0
    if i >= 0 and cigar[i][0] == 5:
        hard = cigar[i][1]; i -= 1
    if i >= 0 and cigar[i][0] == 4:
        soft = cigar[i][1]
    return end - 1 + soft + hard


def paired_filter(flag: int, tid: int, mtid: int, remove_unpaired: bool, remove_chimeric: bool):
    """deduplicate_sam.rs:96-129 with args.paired set, statement by statement.  Returns (passes, counted) where
    counted is a dict of the reference counters this record increments (total_read_count, unmapped, unpaired, chimeric)."""
    c = dict(total=0, unmapped=0, unpaired=0, chimeric=0)
    is_paired, is_last, is_unmapped, mate_unmapped = bool(flag & 0x1), bool(flag & 0x80), bool(flag & 0x4), bool(flag & 0x8)
    if is_paired and is_last:                                   # :96-98
        return False, c
    c["total"] += 1                                             # :100
    if is_unmapped:                                             # :102-108
        c["unmapped"] += 1
        return False, c
    if not is_paired:                                           # :111-116
        c["unpaired"] += 1
        if remove_unpaired:
            return False, c
    if is_paired and mate_unmapped:                             # :118-121
        c["unmapped"] += 1
        return False, c
    if is_paired and tid != mtid:                               # :123-128
        c["chimeric"] += 1
        if remove_chimeric:
            return False, c
    return True, c
