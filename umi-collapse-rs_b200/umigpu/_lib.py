"""ctypes binding of libumigpu.so (the C ABI declared in include/umigpu.h).

The library is the product: if it is missing this module raises — there is no Python or CPU
fallback for any compute step."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("UMIGPU_LIB") or os.path.join(os.path.dirname(_HERE), "csrc", "libumigpu.so")   # UMIGPU_LIB: A/B builds

OK, ERR_ARG, ERR_CUDA, ERR_BAD_BASE, ERR_NOMEM, ERR_UNSUPPORTED, ERR_STATE = 0, -1, -2, -3, -4, -5, -6
ALGO_DIR, ALGO_ADJ, ALGO_ADJ_UPSTREAM, ALGO_CC = 0, 1, 2, 3
MERGE_ANY, MERGE_AVGQUAL, MERGE_MAPQUAL = 0, 1, 2
FLAG_LABELS, FLAG_NO_CULL, FLAG_KERNEL_DIRECT, FLAG_KERNEL_TILES, FLAG_NO_MULTI_INDEX = 1, 2, 4, 8, 16
FLAG_PAIRED, FLAG_REMOVE_UNPAIRED, FLAG_REMOVE_CHIMERIC = 32, 64, 128
STAGES = ["pack", "keys", "sort", "unique", "worklist", "neighbours", "cluster", "emit", "total", "hot_band"]

# every symbol include/umigpu.h declares (tests check that the built library exports all of them)
SYMBOLS = [
    "umigpu_version", "umigpu_last_error", "umigpu_create", "umigpu_destroy", "umigpu_reset",
    "umigpu_push_reads", "umigpu_push_reads_device", "umigpu_run", "umigpu_fetch", "umigpu_finish",
    "umigpu_get_counters", "umigpu_cluster_bucket", "umigpu_remove_near", "umigpu_neighbours",
    "umigpu_avg_qual", "umigpu_stage_ms", "umigpu_launch_count", "umigpu_result_free",
    "umigpu_shard_plan", "umigpu_int_peak", "umigpu_push_bam_records", "umigpu_bam_record_offsets",
    "umigpu_dedup_sharded", "umigpu_free", "umigpu_push_reads_paired", "umigpu_device_init", "umigpu_push_reads_packed",
    "umigpu_pos_key", "umigpu_shard_plan_sorted", "umigpu_xchg_create", "umigpu_xchg_attach_ipc", "umigpu_xchg_attach_local",
    "umigpu_run_sharded", "umigpu_group_create", "umigpu_group_destroy", "umigpu_group_context", "umigpu_group_dedup",
]


class Config(C.Structure):
    _fields_ = [("k", C.c_int32), ("percentage", C.c_float), ("algo", C.c_int32), ("merge", C.c_int32),
                ("umi_len", C.c_uint32), ("device", C.c_int32), ("flags", C.c_uint32), ("reserved", C.c_uint32),
                ("stream", C.c_void_p)]


class Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "total_reads", "n_buckets", "total_umis", "max_umis", "n_kept", "unordered_pairs",
        "pairs_evaluated", "n_edges", "n_tile_items", "n_tile_candidates", "n_sweeps", "n_block_pairs", "n_unmapped",
        "n_unpaired", "n_chimeric", "n_mates_skipped", "key_bits")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class Result(C.Structure):
    _fields_ = [("n_kept", C.c_uint64), ("kept_read_index", C.POINTER(C.c_uint64)), ("n_reads", C.c_uint64),
                ("read_cluster_root", C.POINTER(C.c_uint64)), ("counters", Counters), ("read_umi_rep", C.POINTER(C.c_uint64))]


class Hot(C.Structure):
    """umigpu_hot: the bucket whose neighbour search is split over the devices of a shard group."""
    _fields_ = [("present", C.c_int32), ("owner", C.c_int32), ("read_index", C.c_uint64), ("reads_est", C.c_uint64)]


class UmiGpuError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"umigpu error {code}: {msg}")
        self.code = code


_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc -gencode arch=compute_100a,code=sm_100a).  umigpu has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    p, u64, i32 = C.c_void_p, C.c_uint64, C.c_int32
    lib.umigpu_version.restype = C.c_char_p
    lib.umigpu_last_error.restype = C.c_char_p
    lib.umigpu_last_error.argtypes = [p]
    lib.umigpu_create.argtypes = [C.POINTER(Config), C.POINTER(p)]
    lib.umigpu_destroy.argtypes = [p]
    lib.umigpu_destroy.restype = None
    lib.umigpu_reset.argtypes = [p]
    for name in ("umigpu_push_reads", "umigpu_push_reads_device"):
        getattr(lib, name).argtypes = [p, u64, p, p, p, p, p, p, u64]
    lib.umigpu_device_init.argtypes = [i32]
    lib.umigpu_push_reads_packed.argtypes = [p, u64, p, p, p, p, p, p, u64]
    lib.umigpu_push_reads_paired.argtypes = [p, u64, p, p, p, p, p, p, p, u64]
    lib.umigpu_push_bam_records.argtypes = [p, u64, p, p, C.c_uint8, u64, C.POINTER(u64)]
    lib.umigpu_bam_record_offsets.argtypes = [p, u64, p, u64, C.POINTER(u64), C.POINTER(u64)]
    lib.umigpu_dedup_sharded.argtypes = [C.POINTER(Config), i32, p, u64, p, p, p, p, p, C.POINTER(C.POINTER(u64)), C.POINTER(u64), C.POINTER(Counters)]
    lib.umigpu_free.argtypes = [p]
    lib.umigpu_free.restype = None
    lib.umigpu_run.argtypes = [p]
    lib.umigpu_fetch.argtypes = [p, C.POINTER(Result)]
    lib.umigpu_finish.argtypes = [p, C.POINTER(Result)]
    lib.umigpu_get_counters.argtypes = [p, C.POINTER(Counters)]
    lib.umigpu_cluster_bucket.argtypes = [p, u64, p, p, p, p]
    lib.umigpu_remove_near.argtypes = [p, u64, p, p, p, i32, i32, p]
    lib.umigpu_neighbours.argtypes = [p, u64, p, p, i32, p, p, u64, C.POINTER(u64)]
    lib.umigpu_avg_qual.argtypes = [p, u64, p, p, p]
    lib.umigpu_stage_ms.argtypes = [p, C.c_int, C.POINTER(C.c_float)]
    lib.umigpu_launch_count.argtypes = [p, C.c_int]
    lib.umigpu_launch_count.restype = u64
    lib.umigpu_result_free.argtypes = [p]
    lib.umigpu_result_free.restype = None
    lib.umigpu_shard_plan.argtypes = [u64, p, p, p, i32, p, p]
    lib.umigpu_int_peak.argtypes = [p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.umigpu_pos_key.argtypes = [i32, C.c_int64]
    lib.umigpu_pos_key.restype = C.c_int64
    lib.umigpu_shard_plan_sorted.argtypes = [u64, p, p, p, i32, u64, p, p, C.POINTER(Hot), p]
    lib.umigpu_xchg_create.argtypes = [p, i32, i32, u64, u64, p]
    lib.umigpu_xchg_attach_ipc.argtypes = [p, p]
    lib.umigpu_xchg_attach_local.argtypes = [p, p]
    lib.umigpu_run_sharded.argtypes = [p, C.POINTER(Hot), C.c_int64, C.c_int64]
    lib.umigpu_group_create.argtypes = [C.POINTER(Config), i32, p, C.POINTER(p)]
    lib.umigpu_group_destroy.argtypes = [p]
    lib.umigpu_group_destroy.restype = None
    lib.umigpu_group_context.argtypes = [p, i32]
    lib.umigpu_group_context.restype = p
    lib.umigpu_group_dedup.argtypes = [p, u64, p, p, p, p, p, C.POINTER(C.POINTER(u64)), C.POINTER(u64), C.POINTER(Counters), p]
    _lib = lib
    return lib


def check(rc: int, ctx=None):
    if rc != 0:
        msg = load().umigpu_last_error(ctx)
        raise UmiGpuError(rc, msg.decode() if msg else "")
