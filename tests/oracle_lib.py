"""ctypes wrappers around oracle/liboracle.so (the CPU checker; tests only)."""
import ctypes as C
import os

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_lib = None

ALGO_DIR, ALGO_ADJ_REF, ALGO_ADJ_UPSTREAM, ALGO_CC = 0, 1, 2, 3
MERGE_ANY, MERGE_AVGQUAL, MERGE_MAPQUAL = 0, 1, 2


class Counters(C.Structure):
    _fields_ = [("total_reads", C.c_int64), ("n_buckets", C.c_int64), ("total_umis", C.c_int64), ("max_umis", C.c_int64),
                ("n_kept", C.c_int64), ("dist_calls", C.c_uint64), ("unordered_pairs", C.c_uint64)]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(os.path.join(REPO, "oracle", "liboracle.so"))
        _lib.oracle_dir_threshold.argtypes = [C.c_float, C.c_int32]
        _lib.oracle_dir_threshold.restype = C.c_int32
        _lib.oracle_avg_qual.argtypes = [C.c_void_p, C.c_int64]
        _lib.oracle_avg_qual.restype = C.c_int32
        _lib.oracle_unclipped_pos.argtypes = [C.c_int64, C.c_int, C.c_void_p, C.c_int]
        _lib.oracle_unclipped_pos.restype = C.c_int64
        _lib.oracle_bam_decode.argtypes = [C.c_char_p, C.c_uint64, C.c_int, C.c_uint8, C.c_int] + [C.c_void_p] * 6
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def umi_dist(a: bytes, b: bytes) -> int:
    return lib().oracle_umi_dist(a, b, len(a))


def cluster_bucket(umis: np.ndarray, freq, algo, k, p):
    umis = np.ascontiguousarray(umis, np.uint8)
    n, L = umis.shape
    f = np.ascontiguousarray(freq, np.int32)
    keep = np.zeros(n, np.uint8)
    label = np.zeros(n, np.int32)
    dc = C.c_uint64()
    rc = lib().oracle_cluster_bucket(C.c_int64(n), _p(umis), L, _p(f), algo, k, C.c_float(p), _p(keep), _p(label), C.byref(dc))
    if rc:
        raise RuntimeError(f"oracle_cluster_bucket rc={rc}")
    return keep, label, dc.value


def remove_near(umis: np.ndarray, freq, query: bytes, k, max_freq):
    umis = np.ascontiguousarray(umis, np.uint8)
    n, L = umis.shape
    f = np.ascontiguousarray(freq, np.int32)
    out = np.zeros(n, np.uint8)
    rc = lib().oracle_remove_near(C.c_int64(n), _p(umis), L, _p(f), query, k, C.c_int32(max_freq), _p(out))
    if rc:
        raise RuntimeError(f"oracle_remove_near rc={rc}")
    return out


def dedup(tid, pos, rev, umi, score, algo, merge, k, p, want_roots=False, max_bucket=0, tlen=None):
    tid = np.ascontiguousarray(tid, np.int32); pos = np.ascontiguousarray(pos, np.int64)
    rev = np.ascontiguousarray(rev, np.uint8); umi = np.ascontiguousarray(umi, np.uint8)
    score = None if score is None else np.ascontiguousarray(score, np.int32)
    n = tid.shape[0]
    L = umi.shape[1] if umi.ndim == 2 else 1
    kept = np.zeros(max(n, 1), np.int64)
    roots = np.zeros(max(n, 1), np.int64) if want_roots else None
    ctr = Counters()
    tlen = None if tlen is None else np.ascontiguousarray(tlen, np.int64)
    rc = lib().oracle_dedup_paired(C.c_int64(n), _p(tid), _p(pos), _p(rev), _p(tlen), _p(umi), L, _p(score), algo, merge, k,
                                   C.c_float(p), _p(kept), _p(roots), C.byref(ctr), C.c_int64(max_bucket))
    if rc:
        raise RuntimeError(f"oracle_dedup rc={rc}")
    return kept[: ctr.n_kept].copy(), (roots[:n].copy() if want_roots else None), ctr.as_dict()


def set_threads(n: int):
    """1 = the reference (one clustering thread); > 1 = the all-core variant (buckets handed out to n threads)."""
    lib().oracle_set_threads(int(n))


def avg_qual(q: np.ndarray) -> int:
    q = np.ascontiguousarray(q, np.uint8)
    return lib().oracle_avg_qual(_p(q), len(q))


def unclipped_pos(pos, rev, cigar):
    arr = np.array([(l << 4) | op for op, l in cigar], dtype=np.uint32)
    return lib().oracle_unclipped_pos(pos, int(rev), _p(arr), len(arr))


CLS_MATE, CLS_UNMAPPED, CLS_UNPAIRED, CLS_CHIMERIC = 1, 2, 4, 8


def bam_decode(rec: bytes, umi_len: int, sep: int, use_mapq: bool, paired=False, remove_unpaired=False, remove_chimeric=False):
    """oracle_bam_decode(_paired) on one raw record; returns dict or raises on the reference's panic conditions."""
    tid, score, cls = C.c_int32(), C.c_int32(), C.c_int32()
    pos, tlen = C.c_int64(), C.c_int64()
    rev, valid = C.c_uint8(), C.c_uint8()
    umi = C.create_string_buffer(umi_len)
    rc = lib().oracle_bam_decode_paired(rec, C.c_uint64(len(rec)), umi_len, C.c_uint8(sep), int(use_mapq), int(paired),
                                        int(remove_unpaired), int(remove_chimeric), C.byref(tid), C.byref(pos),
                                        C.byref(rev), umi, C.byref(score), C.byref(valid), C.byref(tlen), C.byref(cls))
    if rc:
        raise RuntimeError(f"oracle_bam_decode rc={rc}")
    return dict(valid=valid.value, tid=tid.value, pos=pos.value, rev=rev.value, umi=umi.raw, score=score.value,
                tlen=tlen.value, cls=cls.value)
