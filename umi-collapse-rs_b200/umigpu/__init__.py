"""umigpu — B200-native UMI clustering hot path of umi-collapse-rs (host-side mirror over libumigpu.so)."""
from ._lib import (ALGO_ADJ, ALGO_ADJ_UPSTREAM, ALGO_CC, ALGO_DIR, FLAG_KERNEL_DIRECT, FLAG_KERNEL_TILES, FLAG_LABELS, FLAG_NO_CULL, FLAG_NO_MULTI_INDEX,
                   FLAG_PAIRED, FLAG_REMOVE_UNPAIRED, FLAG_REMOVE_CHIMERIC,
                   Hot, LIB_PATH, MERGE_ANY, MERGE_AVGQUAL, MERGE_MAPQUAL, STAGES, SYMBOLS, UmiGpuError, load)
from .api import (Adjacency, AdjacencyUpstream, AnyMerge, AvgQualMerge, Cli, ConnectedComponents, Context,
                  DeduplicateGPU, Directional, Group, MapQualMerge, Naive, ReadFreq, dedup_sharded, pack_umis, resolve_cli, shard_plan,
                  shard_plan_sorted)

__all__ = [n for n in dir() if not n.startswith("_")]
