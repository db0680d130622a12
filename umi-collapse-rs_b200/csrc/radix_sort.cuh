// radix_sort.cuh — K2 bucket_group: stable LSD radix sort of (key, read index) pairs, keys of one
// or two 64-bit words held as SoA.  Replaces the two nested std::HashMap probes per read of
// src/deduplicate_sam.rs:148-176: after the sort, reads of one (bucket, UMI) are adjacent and
// buckets are contiguous, so counting and merging become segmented scans.
//
// Per pass (<= 8 bits):  radix_hist -> scan of the digit-major histogram table -> radix_scatter.
// Ranking inside a tile is warp-synchronous: MATCH.ANY groups equal digits, the group leader bumps
// the warp's private counter, so there are no shared-memory atomics and the order is stable.
#pragma once
#include "common.cuh"
#include "scan.cuh"

#define RS_THREADS 512
#define RS_WARPS   (RS_THREADS / 32)
#define RS_ITEMS   16
#define RS_TILE    (RS_THREADS * RS_ITEMS)   // 8192 keys per CTA

struct KeyArr { u64 *w[2]; };   // w[0] = least significant word

// Ranks the tile's keys by digit.  dig[j] = digit of the thread's j-th key (0xffffffff = past the end).
// packed[j] = digit | (rank within (warp, digit) << 8).  On return whist[w][d] = number of keys with
// digit d in warp w's slice (before any __syncthreads).  The digits are loaded by the caller in one
// batch so that all RS_ITEMS global loads are in flight together; this loop only touches shared memory.
__device__ __forceinline__ void rs_rank_tile(const u32 *dig, u32 (*whist)[256], u32 *packed) {
    const u32 w = threadIdx.x >> 5, lane = lane_id();
    for (u32 i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&whist[0][0])[i] = 0;
    __syncthreads();
    const u32 lt = lanemask_lt();
#pragma unroll
    for (int j = 0; j < RS_ITEMS; j++) {
        const u32 d = dig[j];
        const bool active = d != 0xffffffffu;
        const u32 act = __ballot_sync(0xffffffffu, active);
        packed[j] = 0;
        if (active) {
            u32 peers = __match_any_sync(act, d);
            u32 leader = __ffs(peers) - 1;
            u32 old = 0;
            if (lane == leader) { old = whist[w][d]; whist[w][d] = old + __popc(peers); }
            old = __shfl_sync(peers, old, leader);
            packed[j] = d | ((old + __popc(peers & lt)) << 8);
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(RS_THREADS) radix_hist(const u64 *__restrict__ wsel, u64 n, int sh, u32 mask,
                                                         u32 *__restrict__ hist, u32 nblk) {
    __shared__ u32 whist[RS_WARPS][256];
    u32 dig[RS_ITEMS], packed[RS_ITEMS];
    const u64 tile_base = (u64)blockIdx.x * RS_TILE;
    const u32 w = threadIdx.x >> 5, lane = lane_id();
#pragma unroll
    for (int j = 0; j < RS_ITEMS; j++) {
        u64 i = tile_base + (u64)(w * RS_ITEMS + j) * 32 + lane;
        dig[j] = i < n ? ((u32)(wsel[i] >> sh) & mask) : 0xffffffffu;
    }
    rs_rank_tile(dig, whist, packed);
    __syncthreads();
    if (threadIdx.x < 256) {
        u32 s = 0;
#pragma unroll
        for (int ww = 0; ww < RS_WARPS; ww++) s += whist[ww][threadIdx.x];
        hist[(u64)threadIdx.x * nblk + blockIdx.x] = s;
    }
}

template <int NW>
__global__ void __launch_bounds__(RS_THREADS) radix_scatter(KeyArr in, const u32 *__restrict__ idx_in, KeyArr out,
                                                            u32 *__restrict__ idx_out, u64 n, int wsel, int sh, u32 mask,
                                                            const u32 *__restrict__ hist_scanned, u32 nblk, int iota) {
    __shared__ u32 whist[RS_WARPS][256];
    __shared__ u32 sbase[256];
    u32 dig[RS_ITEMS], packed[RS_ITEMS], vidx[RS_ITEMS];
    u64 key[NW][RS_ITEMS];
    const u64 tile_base = (u64)blockIdx.x * RS_TILE;
    const u32 w = threadIdx.x >> 5, lane = lane_id();
    // batch every global load of the tile up front (RS_ITEMS * (NW + 1) independent loads per thread)
#pragma unroll
    for (int j = 0; j < RS_ITEMS; j++) {
        u64 i = tile_base + (u64)(w * RS_ITEMS + j) * 32 + lane;
        bool a = i < n;
#pragma unroll
        for (int k = 0; k < NW; k++) key[k][j] = a ? in.w[k][i] : 0;
        vidx[j] = a ? (iota ? (u32)i : idx_in[i]) : 0;
    }
#pragma unroll
    for (int j = 0; j < RS_ITEMS; j++) {
        u64 i = tile_base + (u64)(w * RS_ITEMS + j) * 32 + lane;
        u64 kw = NW == 1 ? key[0][j] : (wsel == 0 ? key[0][j] : key[NW - 1][j]);
        dig[j] = i < n ? ((u32)(kw >> sh) & mask) : 0xffffffffu;
    }
    rs_rank_tile(dig, whist, packed);
    __syncthreads();
    if (threadIdx.x < 256) {
        u32 run = 0;
#pragma unroll
        for (int ww = 0; ww < RS_WARPS; ww++) { u32 t = whist[ww][threadIdx.x]; whist[ww][threadIdx.x] = run; run += t; }
        sbase[threadIdx.x] = hist_scanned[(u64)threadIdx.x * nblk + blockIdx.x];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < RS_ITEMS; j++) {
        if (dig[j] != 0xffffffffu) {
            u32 d = packed[j] & 0xff, r = packed[j] >> 8;
            u64 pos = (u64)sbase[d] + whist[w][d] + r;
#pragma unroll
            for (int k = 0; k < NW; k++) out.w[k][pos] = key[k][j];
            idx_out[pos] = vidx[j];
        }
    }
}

struct HistLoad  { const u32 *h; __device__ u32 operator()(u64 i) const { return h[i]; } };
struct HistStore { u32 *h; __device__ void operator()(u64 i, u32, u32 ex) const { h[i] = ex; } };

struct SortPass { int word, shift, bits; };

// Plans the passes that cover bit range [0, total_bits) of the key without straddling a word.
static inline std::vector<SortPass> rs_plan(int total_bits) {
    std::vector<SortPass> p;
    for (int word = 0; word < 2; word++) {
        int lo = word * 64, hi = total_bits < lo + 64 ? total_bits : lo + 64;
        if (hi <= lo) break;
        int nbits = hi - lo, npass = (nbits + 7) / 8;
        // spread the bits evenly over the passes (e.g. 53 bits -> 7 passes of 7/8 bits)
        int done = 0;
        for (int i = 0; i < npass; i++) {
            int b = (nbits - done + (npass - i) - 1) / (npass - i);
            p.push_back({word, done, b});
            done += b;
        }
    }
    return p;
}
