"""Host-side mirror of the reference's interfaces for the UMI clustering path, over the C ABI.

Names, argument meaning and error behaviour follow the reference (tkob-vh/umi-collapse-rs):
  Cli                        src/cli.rs:7-77
  Algorithm.apply            src/algo/mod.rs:13-20    -> Directional / Adjacency / ConnectedComponents
  DataStruct (Naive)         src/data/mod.rs:11-17, src/data/naive.rs:14-49
  Merge                      src/merge/mod.rs:6-51
  DeduplicateInterface       src/deduplicate_sam.rs:27-29
All compute happens in libumigpu.so (CUDA, sm_100a); this file only marshals pointers.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
from typing import Optional

import numpy as np

from . import _lib as L


@dataclasses.dataclass
class Cli:
    """src/cli.rs:7-77 — same field names and defaults."""
    mode: str = "bam"
    input: str = ""
    output: str = ""
    k: int = 1
    umi_length: int = 0
    percentage: float = 0.5
    num_threads: int = 1
    umi_separator: int = ord("_")
    algo_str: str = "dir"
    merge_str: Optional[str] = None
    data_str: str = "ngrambktree"
    two_pass: bool = False
    paired: bool = False
    remove_unpaired: bool = False
    remove_chimeric: bool = False
    keep_unmapped: bool = False
    track_clusters: bool = False


ALGO_BY_NAME = {"dir": L.ALGO_DIR, "adj": L.ALGO_ADJ, "cc": L.ALGO_CC, "adj-upstream": L.ALGO_ADJ_UPSTREAM}
MERGE_BY_NAME = {"any": L.MERGE_ANY, "avgqual": L.MERGE_AVGQUAL, "mapqual": L.MERGE_MAPQUAL}


def resolve_cli(args: Cli) -> tuple[int, int]:
    """main.rs:33-39 default merge and main.rs:52-92 dispatch.  `--data` is accepted and ignored exactly
    like the reference (main.rs:86-91 only prints it); unknown algo/merge raise like the panic there."""
    merge_str = args.merge_str or ("avgqual" if args.mode == "fastq" else "mapqual")
    if args.track_clusters and args.two_pass:
        raise ValueError("Cannot track clusters with the two pass algorithm!")        # main.rs:41-43
    if args.paired and args.keep_unmapped:
        raise ValueError("Cannot keep unmapped reads with paired-end reads!")         # main.rs:45-47
    if args.algo_str not in ALGO_BY_NAME or merge_str not in MERGE_BY_NAME:
        raise ValueError(f"Invalid algorithm combination: {args.algo_str} , {merge_str} and {args.data_str}")
    return ALGO_BY_NAME[args.algo_str], MERGE_BY_NAME[merge_str]


def pack_umis(umi_ascii: np.ndarray):
    """[n, L] uint8 ASCII -> (2-bit codes as uint32/uint64, N masks or None): the host-side half of the compact format
    (what the reference's to_bitset does per read, utils/mod.rs:63-83).  Raises on bytes outside ACGTN."""
    n, Lu = umi_ascii.shape
    lut = np.full(256, 255, np.uint8)
    for c, v in zip(b"ACGT", range(4)):
        lut[c] = v
    lut[ord("N")] = 4
    v = lut[umi_ascii]
    if (v == 255).any():
        raise ValueError("Unknown character in UMI sequence")
    isn = v == 4
    code = np.zeros(n, np.uint64)
    nm = np.zeros(n, np.uint32)
    for b in range(Lu):
        code = (code << np.uint64(2)) | np.where(isn[:, b], 0, v[:, b]).astype(np.uint64)
        nm |= isn[:, b].astype(np.uint32) << np.uint32(Lu - 1 - b)
    return (code.astype(np.uint32) if Lu <= 16 else code), (nm if isn.any() else None)


def _ptr(a):
    """Pointer of a numpy array (host) or torch tensor (host or device); None -> NULL."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(a.data_ptr())     # torch tensor


def _is_device(a) -> bool:
    return (not isinstance(a, np.ndarray)) and getattr(a, "is_cuda", False)


class Context:
    """Owns one umigpu_ctx (one CUDA device, one stream)."""

    def __init__(self, umi_len: int, k: int = 1, percentage: float = 0.5, algo: int = L.ALGO_DIR,
                 merge: int = L.MERGE_AVGQUAL, device: int = 0, flags: int = 0, stream: int = 0):
        self._lib = L.load()
        cfg = L.Config(k=k, percentage=percentage, algo=algo, merge=merge, umi_len=umi_len, device=device,
                       flags=flags, reserved=0, stream=C.c_void_p(stream) if stream else None)
        h = C.c_void_p()
        L.check(self._lib.umigpu_create(C.byref(cfg), C.byref(h)))
        self._h = h
        self.umi_len = umi_len
        self.flags = flags
        self._keepalive = []

    def close(self):
        if getattr(self, "_h", None):
            self._lib.umigpu_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def reset(self):
        self._keepalive.clear()
        L.check(self._lib.umigpu_reset(self._h), self._h)

    def push_reads(self, tid, pos, rev, umi, score=None, weight=None, first_read_index: int = 0, tlen=None, sync: bool = True):
        """umigpu_push_reads / _device: arrays of equal length n; umi is [n, umi_len] uint8 ASCII.
        tlen (int64, host arrays) selects umigpu_push_reads_paired: the template length joins the bucket key."""
        n = int(tid.shape[0])
        dev = _is_device(tid)
        if tlen is not None:
            assert not dev, "umigpu_push_reads_paired takes host arrays"
            arrs = [np.ascontiguousarray(a, dtype=dt) if a is not None else None
                    for a, dt in ((tid, "int32"), (pos, "int64"), (rev, "uint8"), (tlen, "int64"), (umi, "uint8"),
                                  (score, "int32"), (weight, "int32"))]
            self._keepalive.append(arrs)
            L.check(self._lib.umigpu_push_reads_paired(self._h, n, *[_ptr(a) for a in arrs], first_read_index), self._h)
            return
        arrs = []
        for a, dt in ((tid, "int32"), (pos, "int64"), (rev, "uint8"), (umi, "uint8"), (score, "int32"), (weight, "int32")):
            if a is None:
                arrs.append(None)
                continue
            if isinstance(a, np.ndarray):
                a = np.ascontiguousarray(a, dtype=dt)
            else:
                assert str(a.dtype).endswith(dt), (a.dtype, dt)
                a = a.contiguous()
            arrs.append(a)
        self._keepalive.append(arrs)     # async copies: keep the sources alive until fetch/reset
        if dev and sync:
            # the context works on its own (non-blocking) stream: the kernels that produced the tensors on torch's
            # stream must have finished before the library reads them (sync=False: the caller guarantees that)
            import torch
            torch.cuda.current_stream(tid.device).synchronize()
        fn = self._lib.umigpu_push_reads_device if dev else self._lib.umigpu_push_reads
        L.check(fn(self._h, n, *[_ptr(a) for a in arrs], first_read_index), self._h)

    def push_reads_packed(self, tid, pos32, rev, umi_2bit, n_mask=None, score8=None, first_read_index: int = 0):
        """umigpu_push_reads_packed: the compact host format (int32 positions, 2-bit UMIs as uint32/uint64, uint8 scores)."""
        n = int(tid.shape[0])
        udt = "uint32" if self.umi_len <= 16 else "uint64"
        arrs = [None if a is None else np.ascontiguousarray(a, dtype=dt)
                for a, dt in ((tid, "int32"), (pos32, "int32"), (rev, "uint8"), (umi_2bit, udt), (n_mask, "uint32"), (score8, "uint8"))]
        self._keepalive.append(arrs)
        L.check(self._lib.umigpu_push_reads_packed(self._h, n, *[_ptr(a) for a in arrs], first_read_index), self._h)

    def run(self):
        L.check(self._lib.umigpu_run(self._h), self._h)

    # ---- shard group (several devices, one dataset): include/umigpu.h "Several devices, ONE dataset" ----
    def xchg_create(self, rank: int, n_ranks: int, max_hot_uniques: int, max_hot_edges: int) -> bytes:
        """umigpu_xchg_create: this rank's exchange window; returns its 64-byte CUDA IPC handle."""
        h = (C.c_uint8 * 64)()
        L.check(self._lib.umigpu_xchg_create(self._h, rank, n_ranks, max_hot_uniques, max_hot_edges, h), self._h)
        return bytes(h)

    def xchg_attach_ipc(self, handles: list):
        buf = (C.c_uint8 * (64 * len(handles))).from_buffer_copy(b"".join(handles))
        L.check(self._lib.umigpu_xchg_attach_ipc(self._h, buf), self._h)

    def xchg_attach_local(self, group: list):
        arr = (C.c_void_p * len(group))(*[g._h for g in group])
        L.check(self._lib.umigpu_xchg_attach_local(self._h, arr), self._h)

    def run_sharded(self, hot=None, key_lo: int = -2 ** 63, key_hi: int = 2 ** 63 - 1):
        L.check(self._lib.umigpu_run_sharded(self._h, C.byref(hot) if hot is not None else None, key_lo, key_hi), self._h)

    def fetch(self, copy: bool = True):
        """copy=False returns views of the library-owned (pinned) result buffers: valid until the next
        reset / run / close of this context."""
        res = L.Result()
        L.check(self._lib.umigpu_fetch(self._h, C.byref(res)), self._h)
        self._keepalive.clear()
        kept = np.ctypeslib.as_array(res.kept_read_index, shape=(res.n_kept,)) if res.n_kept else np.zeros(0, np.uint64)
        roots = None
        if res.read_cluster_root:
            roots = np.ctypeslib.as_array(res.read_cluster_root, shape=(res.n_reads,))
        self.last_umi_rep = None
        if res.read_umi_rep:
            self.last_umi_rep = np.ctypeslib.as_array(res.read_umi_rep, shape=(res.n_reads,)).copy()
        if copy:
            kept = kept.copy()
            roots = None if roots is None else roots.copy()
        return kept, roots, res.counters.as_dict()

    def finish(self, copy: bool = True):
        self.run()
        return self.fetch(copy)

    def counters(self) -> dict:
        c = L.Counters()
        L.check(self._lib.umigpu_get_counters(self._h, C.byref(c)), self._h)
        return c.as_dict()

    def stage_ms(self) -> dict:
        out = {}
        for i, name in enumerate(L.STAGES):
            ms = C.c_float()
            L.check(self._lib.umigpu_stage_ms(self._h, i, C.byref(ms)), self._h)
            out[name] = float(ms.value)
        return out

    def launch_count(self, reset: bool = False) -> int:
        return int(self._lib.umigpu_launch_count(self._h, 1 if reset else 0))

    def cluster_bucket(self, umis: np.ndarray, freq: np.ndarray):
        umis = np.ascontiguousarray(umis, dtype=np.uint8)
        freq = np.ascontiguousarray(freq, dtype=np.int32)
        n = int(freq.shape[0])
        keep = np.zeros(n, np.uint8)
        label = np.zeros(n, np.int32)
        L.check(self._lib.umigpu_cluster_bucket(self._h, n, _ptr(umis), _ptr(freq), _ptr(keep), _ptr(label)), self._h)
        return keep, label

    def remove_near(self, umis: np.ndarray, freq: np.ndarray, query: bytes, k: int, max_freq: int) -> np.ndarray:
        umis = np.ascontiguousarray(umis, dtype=np.uint8)
        freq = np.ascontiguousarray(freq, dtype=np.int32)
        n = int(freq.shape[0])
        out = np.zeros(n, np.uint8)
        q = np.frombuffer(bytes(query), dtype=np.uint8).copy()
        L.check(self._lib.umigpu_remove_near(self._h, n, _ptr(umis), _ptr(freq), _ptr(q), k, max_freq, _ptr(out)), self._h)
        return out

    def neighbours(self, umis: np.ndarray, freq: np.ndarray, apply_rule: bool = True):
        umis = np.ascontiguousarray(umis, dtype=np.uint8)
        freq = np.ascontiguousarray(freq, dtype=np.int32)
        n = int(freq.shape[0])
        row_ptr = np.zeros(n + 1, np.uint64)
        cap = max(1024, 8 * n)
        while True:
            col = np.zeros(cap, np.uint32)
            ne = C.c_uint64()
            rc = self._lib.umigpu_neighbours(self._h, n, _ptr(umis), _ptr(freq), 1 if apply_rule else 0, _ptr(row_ptr), _ptr(col), cap, C.byref(ne))
            if rc == L.ERR_ARG and ne.value > cap:
                cap = int(ne.value)
                continue
            L.check(rc, self._h)
            return row_ptr, col[: ne.value]

    def avg_qual(self, qual: np.ndarray, offsets: np.ndarray) -> np.ndarray:
        qual = np.ascontiguousarray(qual, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = int(offsets.shape[0]) - 1
        out = np.zeros(n, np.int32)
        L.check(self._lib.umigpu_avg_qual(self._h, n, _ptr(qual), _ptr(offsets), _ptr(out)), self._h)
        return out

    def int_peak(self):
        a, b = C.c_double(), C.c_double()
        L.check(self._lib.umigpu_int_peak(self._h, C.byref(a), C.byref(b)), self._h)
        return float(a.value), float(b.value)


def dedup_sharded(umi_len: int, devices, tid, pos, rev, umi, score=None, k: int = 1, percentage: float = 0.5,
                  algo: int = L.ALGO_DIR, merge: int = L.MERGE_AVGQUAL, flags: int = 0):
    """umigpu_dedup_sharded: one call, one context + host thread per device, survivors merged in input order."""
    lib = L.load()
    cfg = L.Config(k=k, percentage=percentage, algo=algo, merge=merge, umi_len=umi_len, device=0, flags=flags, reserved=0, stream=None)
    tid = np.ascontiguousarray(tid, np.int32); pos = np.ascontiguousarray(pos, np.int64); rev = np.ascontiguousarray(rev, np.uint8)
    umi = np.ascontiguousarray(umi, np.uint8); score = None if score is None else np.ascontiguousarray(score, np.int32)
    dev = np.ascontiguousarray(devices, np.int32)
    kept_p, nk, ctr = C.POINTER(C.c_uint64)(), C.c_uint64(), L.Counters()
    L.check(lib.umigpu_dedup_sharded(C.byref(cfg), len(dev), _ptr(dev), tid.shape[0], _ptr(tid), _ptr(pos), _ptr(rev), _ptr(umi), _ptr(score),
                                     C.byref(kept_p), C.byref(nk), C.byref(ctr)))
    kept = np.ctypeslib.as_array(kept_p, shape=(nk.value,)).copy() if nk.value else np.zeros(0, np.uint64)
    lib.umigpu_free(kept_p)
    return kept, ctr.as_dict()


def shard_plan_sorted(tid, pos, rev, n_shards: int, hot_min_reads: int = 0):
    """umigpu_shard_plan_sorted: contiguous slices of a coordinate-sorted stream.  Returns (cuts uint64[n_shards + 1],
    cut_keys int64[n_shards + 1], hot (umigpu_hot), modelled cost per shard)."""
    lib = L.load()
    tid = np.ascontiguousarray(tid, np.int32); pos = np.ascontiguousarray(pos, np.int64); rev = np.ascontiguousarray(rev, np.uint8)
    cuts = np.zeros(n_shards + 1, np.uint64)
    keys = np.zeros(n_shards + 1, np.int64)
    cost = np.zeros(n_shards, np.float64)
    hot = L.Hot()
    L.check(lib.umigpu_shard_plan_sorted(tid.shape[0], _ptr(tid), _ptr(pos), _ptr(rev), n_shards, hot_min_reads, _ptr(cuts), _ptr(keys),
                                         C.byref(hot), _ptr(cost)))
    return cuts, keys, hot, cost


class Group:
    """umigpu_group_*: a persistent group of contexts in ONE process, one per device; dedup() shards one coordinate-sorted
    dataset over them (contiguous slices, hot bucket split over NVLink) and returns the merged kept list."""

    def __init__(self, umi_len: int, devices, k: int = 1, percentage: float = 0.5, algo: int = L.ALGO_DIR,
                 merge: int = L.MERGE_AVGQUAL, flags: int = 0):
        self._lib = L.load()
        cfg = L.Config(k=k, percentage=percentage, algo=algo, merge=merge, umi_len=umi_len, device=0, flags=flags, reserved=0, stream=None)
        dev = np.ascontiguousarray(devices, np.int32)
        self.n = len(dev)
        h = C.c_void_p()
        L.check(self._lib.umigpu_group_create(C.byref(cfg), self.n, _ptr(dev), C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.umigpu_group_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def dedup(self, tid, pos, rev, umi, score=None):
        tid = np.ascontiguousarray(tid, np.int32); pos = np.ascontiguousarray(pos, np.int64); rev = np.ascontiguousarray(rev, np.uint8)
        umi = np.ascontiguousarray(umi, np.uint8); score = None if score is None else np.ascontiguousarray(score, np.int32)
        kept_p, nk, ctr = C.POINTER(C.c_uint64)(), C.c_uint64(), L.Counters()
        ms = np.zeros(self.n, np.float32)
        L.check(self._lib.umigpu_group_dedup(self._h, tid.shape[0], _ptr(tid), _ptr(pos), _ptr(rev), _ptr(umi), _ptr(score),
                                             C.byref(kept_p), C.byref(nk), C.byref(ctr), _ptr(ms)))
        kept = np.ctypeslib.as_array(kept_p, shape=(nk.value,)).copy() if nk.value else np.zeros(0, np.uint64)
        self._lib.umigpu_free(kept_p)
        return kept, ctr.as_dict(), ms


def shard_plan(tid, pos, rev, n_shards: int):
    """umigpu_shard_plan: LPT over per-bucket cost; returns (shard_of_read int32[n], shard_cost uint64[n_shards])."""
    lib = L.load()
    tid = np.ascontiguousarray(tid, np.int32); pos = np.ascontiguousarray(pos, np.int64); rev = np.ascontiguousarray(rev, np.uint8)
    out = np.zeros(tid.shape[0], np.int32)
    cost = np.zeros(n_shards, np.uint64)
    L.check(lib.umigpu_shard_plan(tid.shape[0], _ptr(tid), _ptr(pos), _ptr(rev), n_shards, _ptr(out), _ptr(cost)))
    return out, cost


# ---------------------------------------------------------------------------------------------
# Reference-shaped objects
# ---------------------------------------------------------------------------------------------
class AnyMerge:                                   # merge/mod.rs:10-23
    code = L.MERGE_ANY

    def merge(self, a, b) -> bool:
        return True


class AvgQualMerge:                               # merge/mod.rs:25-37
    code = L.MERGE_AVGQUAL

    def merge(self, a, b) -> bool:
        return a.get_avg_qual() >= b.get_avg_qual()


class MapQualMerge:                               # merge/mod.rs:39-51
    code = L.MERGE_MAPQUAL

    def merge(self, a, b) -> bool:
        return a.get_map_qual() >= b.get_map_qual()


@dataclasses.dataclass
class ReadFreq:                                   # utils/read_freq.rs:4-7
    read: object
    freq: int


class _Algo:
    """Algorithm::apply(&mut self, reads: &HashMap<&BitSet,&ReadFreq<R>>, tracker, umi_length) -> Vec<&R>"""
    code = L.ALGO_DIR

    def __init__(self, args: Cli, device: int = 0):
        self.k, self.percentage, self.track_cluster = args.k, args.percentage, args.track_clusters
        self.device = device
        self._ctx = None

    def apply(self, reads: dict, tracker=None, umi_length: int = 0) -> list:
        if not reads:
            return []
        umis = list(reads.keys())
        umi_length = umi_length or len(umis[0])
        if self._ctx is None or self._ctx.umi_len != umi_length:
            self._ctx = Context(umi_length, self.k, self.percentage, self.code, L.MERGE_ANY, self.device)
        arr = np.frombuffer(b"".join(umis), dtype=np.uint8).reshape(len(umis), umi_length)
        freq = np.array([reads[u].freq for u in umis], dtype=np.int32)
        keep, label = self._ctx.cluster_bucket(arr, freq)
        if tracker is not None and self.track_cluster:
            tracker.extend((umis[i], umis[label[i]]) for i in range(len(umis)))
        # reference order: frequency descending (directional.rs:67-72), canonical tie-break
        kept = [i for i in range(len(umis)) if keep[i]]
        kept.sort(key=lambda i: (-int(freq[i]), _canon(umis[i])))
        return [reads[umis[i]].read for i in kept]


_CANON = bytes.maketrans(b"ACGTN", b"01234")


def _canon(u: bytes) -> bytes:
    return bytes(u).translate(_CANON)


class Directional(_Algo):                         # algo/directional.rs
    code = L.ALGO_DIR


class Adjacency(_Algo):                           # algo/adjacency.rs (as written)
    code = L.ALGO_ADJ


class AdjacencyUpstream(_Algo):                   # opt-in, not a reference behaviour
    code = L.ALGO_ADJ_UPSTREAM


class ConnectedComponents(_Algo):                 # --algo cc: help text only in the reference
    code = L.ALGO_CC


class Naive:
    """DataStruct over the GPU distance kernel: new / remove_near / contains (data/naive.rs:22-44).
    The membership set lives on the host exactly like the reference's HashMap; each remove_near is one
    kernel over the UMIs still present."""

    def __init__(self, umi_freq: dict, umi_length: int, max_edits: int, device: int = 0):
        self.umi_freq = dict(umi_freq)
        self.umi_length = umi_length
        self._ctx = Context(umi_length, max_edits, 0.5, L.ALGO_DIR, L.MERGE_ANY, device)

    @classmethod
    def new(cls, umi_freq: dict, umi_length: int, max_edits: int):
        return cls(umi_freq, umi_length, max_edits)

    def remove_near(self, umi: bytes, k: int, max_freq: int) -> set:
        if not self.umi_freq:
            return set()
        keys = list(self.umi_freq.keys())
        arr = np.frombuffer(b"".join(keys), dtype=np.uint8).reshape(len(keys), self.umi_length)
        freq = np.array([self.umi_freq[u] for u in keys], dtype=np.int32)
        flags = self._ctx.remove_near(arr, freq, umi, k, max_freq)
        res = {keys[i] for i in range(len(keys)) if flags[i]}
        for u in res:
            del self.umi_freq[u]
        return res

    def contains(self, umi: bytes) -> bool:
        return umi in self.umi_freq

    def stats(self) -> dict:
        return {}


class DeduplicateGPU:
    """The new `impl DeduplicateInterface` (deduplicate_sam.rs:27-29): all buckets in one batch.
    dedup_arrays is the SoA core; deduplicate_and_merge (BAM in / BAM out) lives in umigpu.bamio."""

    def __init__(self, args: Cli, device: int = 0, flags: int = 0, stream: int = 0):
        self.args = args
        self.algo, self.merge = resolve_cli(args)
        self.device, self.flags, self.stream = device, flags, stream
        self._ctx = None
        self.counters: dict = {}

    def context(self, umi_len: int) -> Context:
        if self._ctx is None or self._ctx.umi_len != umi_len:
            fl = self.flags | (L.FLAG_LABELS if self.args.track_clusters else 0)
            self._ctx = Context(umi_len, self.args.k, self.args.percentage, self.algo, self.merge, self.device, fl, self.stream)
        return self._ctx

    def dedup_arrays(self, tid, pos, rev, umi, score, chunk: int = 0):
        """Returns (kept read indices ascending, per-read cluster root or None, counters)."""
        ctx = self.context(int(umi.shape[1]))
        ctx.reset()
        n = int(tid.shape[0])
        if chunk and chunk < n:
            for s in range(0, n, chunk):
                e = min(n, s + chunk)
                ctx.push_reads(tid[s:e], pos[s:e], rev[s:e], umi[s:e], None if score is None else score[s:e], None, s)
        else:
            ctx.push_reads(tid, pos, rev, umi, score, None, 0)
        kept, roots, ctr = ctx.finish()
        self.counters = ctr
        return kept, roots, ctr
