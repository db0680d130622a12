// radix_sort.cuh — K2 bucket_group: stable LSD radix sort of (key, read index) pairs, keys of one
// or two 64-bit words held as SoA.  Replaces the two nested std::HashMap probes per read of
// src/deduplicate_sam.rs:148-176: after the sort, reads of one (bucket, UMI) are adjacent and
// buckets are contiguous, so counting and merging become segmented scans.
//
// One-sweep design: a single pre-pass (radix_global_hist) reads the keys once and builds the 256-bin
// digit histogram of EVERY pass (digit counts do not depend on the order of the keys).  Each pass is
// then one kernel (radix_onesweep): a CTA takes the next tile (atomic ticket, so all predecessors are
// resident), ranks its keys by digit, publishes its per-digit counts and obtains its global offsets
// by decoupled look-back over the predecessors' published states — keys and indices are read once
// and written once per pass.
// Ranking inside a tile is warp-synchronous: per-bit ballots group equal digits, the group leader bumps
// the warp's private counter, so there are no shared-memory atomics and the order is stable.
#pragma once
#include "common.cuh"
#include "scan.cuh"

#ifndef RS_RB
#define RS_RB 8                  // digit width in bits (RS_RADIX bins)
#endif
#define RS_RADIX (1 << RS_RB)
#ifndef RS_THREADS
#define RS_THREADS 512         // A/B builds: -DRS_THREADS=256 (4 CTAs/SM)
#endif
#define RS_WARPS   (RS_THREADS / 32)
#define RS_MAX_PASSES 16
#define RS_LB 8             // look-back window (predecessor tiles examined per round)

struct KeyArr { u64 *w[2]; };   // w[0] = least significant word
struct SortPass { int word, shift, bits; };
struct SortPlan { int npass; SortPass p[RS_MAX_PASSES]; };

#define RS_FLAG_AGG    (1ull << 62)
#define RS_FLAG_PREFIX (2ull << 62)
#define RS_FLAG_MASK   (3ull << 62)

// ---- pre-pass: global digit histograms of all passes in one read of the keys -------------------
// Each thread walks 8 consecutive keys and aggregates runs of equal digits in a register before
// touching shared memory, so the highly repetitive high digits of coordinate-sorted input do not
// serialise on one shared-memory address.
#define GH_THREADS 256
#define GH_ITEMS   8
__global__ void __launch_bounds__(GH_THREADS) radix_global_hist(KeyArr in, u64 n, SortPlan plan, u32 *__restrict__ ghist /* [npass][RS_RADIX] */) {
    __shared__ u32 sh[RS_MAX_PASSES * RS_RADIX];
    for (u32 i = threadIdx.x; i < (u32)plan.npass * RS_RADIX; i += GH_THREADS) sh[i] = 0;
    __syncthreads();
    const u64 stride = (u64)gridDim.x * GH_THREADS * GH_ITEMS;
    for (u64 base = ((u64)blockIdx.x * GH_THREADS + threadIdx.x) * GH_ITEMS; base < n; base += stride) {
        u64 k0[GH_ITEMS], k1[GH_ITEMS];
#pragma unroll
        for (int j = 0; j < GH_ITEMS; j++) {
            u64 i = base + j;
            k0[j] = i < n ? in.w[0][i] : 0;
            k1[j] = (i < n && in.w[1]) ? in.w[1][i] : 0;
        }
        const int cnt = (int)min((u64)GH_ITEMS, n - base);
        for (int p = 0; p < plan.npass; p++) {
            const int sh_ = plan.p[p].shift; const u32 mask = (1u << plan.p[p].bits) - 1; const bool hi = plan.p[p].word != 0;
            u32 run_d = 0xffffffffu, run_c = 0;
#pragma unroll
            for (int j = 0; j < GH_ITEMS; j++) {
                if (j < cnt) {
                    u32 d = (u32)((hi ? k1[j] : k0[j]) >> sh_) & mask;
                    if (d != run_d) { if (run_c) atomicAdd(&sh[p * RS_RADIX + run_d], run_c); run_d = d; run_c = 0; }
                    run_c++;
                }
            }
            if (run_c) atomicAdd(&sh[p * RS_RADIX + run_d], run_c);
        }
    }
    __syncthreads();
    for (u32 i = threadIdx.x; i < (u32)plan.npass * RS_RADIX; i += GH_THREADS) if (sh[i]) atomicAdd(&ghist[i], sh[i]);
}

// exclusive scan of each pass's RS_RADIX bins (one CTA; tiny)
__global__ void __launch_bounds__(RS_RADIX) radix_digit_starts(u32 *ghist, int npass) {
    __shared__ u32 s[RS_RADIX];
    for (int p = 0; p < npass; p++) {
        u32 v = ghist[p * RS_RADIX + threadIdx.x];
        s[threadIdx.x] = v;
        __syncthreads();
        // Hillis-Steele inclusive scan
        for (int o = 1; o < RS_RADIX; o <<= 1) {
            u32 t = threadIdx.x >= (u32)o ? s[threadIdx.x - o] : 0;
            __syncthreads();
            s[threadIdx.x] += t;
            __syncthreads();
        }
        ghist[p * RS_RADIX + threadIdx.x] = s[threadIdx.x] - v;
        __syncthreads();
    }
}

// Ranks the thread's keys by digit.  packed[j] = digit | (rank within (warp, digit) << RS_RB), or
// 0xffffffff past the end.  On return whist[w][d] = number of keys with digit d in warp w's slice.
template <int ITEMS, bool FULL>
__device__ __forceinline__ void rs_rank_tile(u32 *packed /* in: digit or 0xffffffff */, u32 (*whist)[RS_RADIX]) {
    const u32 w = threadIdx.x >> 5, lane = lane_id();
    for (u32 i = threadIdx.x; i < RS_WARPS * RS_RADIX; i += RS_THREADS) (&whist[0][0])[i] = 0;
    __syncthreads();
    const u32 lt = lanemask_lt();
#pragma unroll
    for (int j = 0; j < ITEMS; j++) {
        const u32 d = packed[j];
        const bool active = FULL || d != 0xffffffffu;
        // lanes holding the same digit: eight ballots (one per digit bit) instead of MATCH.ANY, whose cost grows
        // with the number of distinct values in the warp (up to 32 for the random UMI digits).  Every lane takes
        // part in every collective with the full mask (no divergent collectives); lanes past the end are masked
        // out of the result by `act`.
        u32 peers = FULL ? 0xffffffffu : __ballot_sync(0xffffffffu, active);
#pragma unroll
        for (int b = 0; b < RS_RB; b++) {
            const bool bit = (d & (1u << b)) != 0u;
            const u32 m = __ballot_sync(0xffffffffu, bit);
            peers &= bit ? m : ~m;
        }
        const u32 leader = active ? (u32)__ffs(peers) - 1u : lane;
        u32 old = 0;
        if (active && lane == leader) { old = whist[w][d]; whist[w][d] = old + __popc(peers); }
        old = __shfl_sync(0xffffffffu, old, leader);
        if (active) packed[j] = d | ((old + __popc(peers & lt)) << RS_RB);
        __syncwarp();
    }
}

// One pass.  After ranking, the tile is permuted into digit order in shared memory, so that the
// global writes of a warp are runs of consecutive addresses (one run per digit) instead of 32
// unrelated 8-byte fragments.  err[0] is set if a look-back ever exceeds its spin budget (cannot
// happen with ticketed tiles; it turns a would-be hang into a reported error).
struct RsShared {
    u32 whist[RS_WARPS][RS_RADIX];
    u32 sbase[RS_RADIX], slocal[RS_RADIX];
    u32 sscan[RS_THREADS / 32 + 1];
    u32 s_tile;
};

// FULL = the tile has all TILE keys: no per-key bounds checks anywhere (all tiles but the last)
template <int NW, int ITEMS, bool FULL>
__device__ __forceinline__ void rs_onesweep_body(
    RsShared &S, u64 *skey, u32 *sidx, const u32 tile, const u32 tile_count,
    KeyArr in, const u32 *__restrict__ idx_in, KeyArr out, u32 *__restrict__ idx_out, u64 n, int wsel, int sh, u32 mask,
    const u32 *__restrict__ digit_start, unsigned long long *tile_state, u32 *err, int iota) {
    constexpr u32 TILE = RS_THREADS * ITEMS;
    u32 (*whist)[RS_RADIX] = S.whist;
    u32 *sbase = S.sbase, *slocal = S.slocal, *sscan = S.sscan;
    const u64 tile_base = (u64)tile * TILE;
    const u32 w = threadIdx.x >> 5, lane = lane_id();
    u32 packed[ITEMS], vidx[ITEMS];
    u64 key[NW][ITEMS];
    // every global load of the tile is issued before anything else (ITEMS * (NW + 1) loads in flight per thread)
#pragma unroll
    for (int j = 0; j < ITEMS; j++) {
        u64 i = tile_base + (u64)(w * ITEMS + j) * 32 + lane;
        bool a = FULL || i < n;
#pragma unroll
        for (int k = 0; k < NW; k++) key[k][j] = a ? in.w[k][i] : 0;
        vidx[j] = a ? (iota ? (u32)i : idx_in[i]) : 0;
    }
#pragma unroll
    for (int j = 0; j < ITEMS; j++) {
        u64 i = tile_base + (u64)(w * ITEMS + j) * 32 + lane;
        u64 kw = NW == 1 ? key[0][j] : (wsel == 0 ? key[0][j] : key[NW - 1][j]);
        packed[j] = (FULL || i < n) ? ((u32)(kw >> sh) & mask) : 0xffffffffu;
    }
    rs_rank_tile<ITEMS, FULL>(packed, whist);
    __syncthreads();
    u32 total = 0;
    if (threadIdx.x < RS_RADIX) {
        const u32 d = threadIdx.x;
#pragma unroll
        for (int ww = 0; ww < RS_WARPS; ww++) { u32 t = whist[ww][d]; whist[ww][d] = total; total += t; }
        // publish this tile's digit counts right away; the look-back itself is deferred until after the shared-memory
        // permutation so that the predecessors have had that long to publish theirs
        atomicExch(tile_state + (u64)tile * RS_RADIX + d, (tile == 0 ? RS_FLAG_PREFIX : RS_FLAG_AGG) | (unsigned long long)total);
    }
    // start of each digit inside the tile (exclusive scan of the tile's digit counts)
    u32 tot_all;
    const u32 lstart = block_exclusive_scan<u32, RS_THREADS>(total, sscan, &tot_all);
    if (threadIdx.x < RS_RADIX) slocal[threadIdx.x] = lstart;
    __syncthreads();
    // permute the tile into digit order in shared memory
#pragma unroll
    for (int j = 0; j < ITEMS; j++) {
        if (FULL || packed[j] != 0xffffffffu) {
            const u32 d = packed[j] & (RS_RADIX - 1), r = packed[j] >> RS_RB;
            const u32 lp = slocal[d] + whist[w][d] + r;
#pragma unroll
            for (int k = 0; k < NW; k++) skey[(size_t)k * TILE + lp] = key[k][j];
            sidx[lp] = vidx[j];
        }
    }
    if (threadIdx.x < RS_RADIX) {
        const u32 d = threadIdx.x;
        unsigned long long *mine = tile_state + (u64)tile * RS_RADIX + d;
        u64 excl = 0;
        if (tile != 0) {
            // decoupled look-back, RS_LB predecessors per round: the loads of a round are independent
            // (one latency per round instead of one per tile); a not-yet-published state ends the round
            u32 p = tile;          // predecessors [0, p) are still to be accounted for
            u32 spins = 0;
            bool done = false;
            while (!done && p > 0) {
                unsigned long long v[RS_LB];
#pragma unroll
                for (int i = 0; i < RS_LB; i++) {
                    const volatile unsigned long long *prev = tile_state + (u64)(p > (u32)i ? p - 1 - i : 0) * RS_RADIX + d;
                    v[i] = *prev;
                }
#pragma unroll
                for (int i = 0; i < RS_LB; i++) {
                    if (done || p == 0) break;
                    if ((v[i] & RS_FLAG_MASK) == 0) {
                        if (++spins > (1u << 24)) { err[0] = 1; done = true; }
                        if (i == 0) __nanosleep(40);
                        break;                                   // re-poll from this predecessor
                    }
                    excl += v[i] & ~RS_FLAG_MASK;
                    p--;
                    if ((v[i] & RS_FLAG_MASK) == RS_FLAG_PREFIX) done = true;
                }
            }
            atomicExch(mine, RS_FLAG_PREFIX | (unsigned long long)(excl + total));
        }
        sbase[d] = digit_start[d] + (u32)excl;
    }
    __syncthreads();
    // coalesced write-out: consecutive threads hold consecutive elements of (mostly) the same digit
#pragma unroll
    for (int j = 0; j < ITEMS; j++) {
        const u32 i = threadIdx.x + j * RS_THREADS;
        if (FULL || i < tile_count) {
            u64 kk[NW];
#pragma unroll
            for (int k = 0; k < NW; k++) kk[k] = skey[(size_t)k * TILE + i];
            const u64 kw = NW == 1 ? kk[0] : (wsel == 0 ? kk[0] : kk[NW - 1]);
            const u32 d = (u32)(kw >> sh) & mask;
            const u64 pos = (u64)sbase[d] + (i - slocal[d]);
#pragma unroll
            for (int k = 0; k < NW; k++) out.w[k][pos] = kk[k];
            idx_out[pos] = sidx[i];
        }
    }
}

template <int NW, int ITEMS>
__global__ void __launch_bounds__(RS_THREADS, (NW == 1 ? 1024 / RS_THREADS : 1)) radix_onesweep(
    KeyArr in, const u32 *__restrict__ idx_in, KeyArr out, u32 *__restrict__ idx_out, u64 n, int wsel, int sh, u32 mask,
    const u32 *__restrict__ digit_start, unsigned long long *tile_state /* [ntiles][RS_RADIX] */, u32 *ticket, u32 *err, int iota) {
    constexpr u32 TILE = RS_THREADS * ITEMS;
    extern __shared__ __align__(16) unsigned char rs_dyn[];          // u64 skey[NW][TILE]; u32 sidx[TILE]
    u64 *skey = reinterpret_cast<u64 *>(rs_dyn);
    u32 *sidx = reinterpret_cast<u32 *>(rs_dyn + (size_t)NW * TILE * 8);
    __shared__ RsShared S;
    if (threadIdx.x == 0) S.s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const u32 tile = S.s_tile;
    const u32 tile_count = (u32)min((u64)TILE, n - (u64)tile * TILE);
    if (tile_count == TILE)
        rs_onesweep_body<NW, ITEMS, true>(S, skey, sidx, tile, tile_count, in, idx_in, out, idx_out, n, wsel, sh, mask, digit_start, tile_state, err, iota);
    else
        rs_onesweep_body<NW, ITEMS, false>(S, skey, sidx, tile, tile_count, in, idx_in, out, idx_out, n, wsel, sh, mask, digit_start, tile_state, err, iota);
}

// Plans the passes that cover bit range [0, total_bits) of the key without straddling a word.
static inline SortPlan rs_plan(int total_bits) {
    SortPlan pl; pl.npass = 0;
    for (int word = 0; word < 2; word++) {
        int lo = word * 64, hi = total_bits < lo + 64 ? total_bits : lo + 64;
        if (hi <= lo) break;
        int nbits = hi - lo, npass = (nbits + RS_RB - 1) / RS_RB;
        // spread the bits evenly over the passes (e.g. 53 bits -> 7 passes of 7/8 bits)
        int done = 0;
        for (int i = 0; i < npass; i++) {
            int b = (nbits - done + (npass - i) - 1) / (npass - i);
            pl.p[pl.npass++] = {word, done, b};
            done += b;
        }
    }
    return pl;
}
