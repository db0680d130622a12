// seg_sort.cuh — K2 bucket_group for coordinate-sorted input (what a BAM is): a SEGMENTED sort.
//
// The sort key is [P | S] with P = (contig, position) and S = (strand, [tlen,] UMI).  When the reads arrive ordered by P —
// src/deduplicate_sam.rs:93 streams a coordinate-sorted BAM — the generic LSD radix sort (radix_sort.cuh) spends more than
// half of its passes re-establishing an order the input already has.  Here the runs of equal P ("segments") stay where they
// are and only their insides are ordered by S:
//   plan     seg_block_summary_kernel  one read of the keys: per block of 2048 positions the number of segment heads, the
//                                      first and the last head; also detects input that is NOT ordered by P (-> generic sort)
//            seg_plan_kernel           one CTA over the block table (n / 2048 entries): previous / next head per block, the
//                                      list of BIG segments (> 2048 reads: only the last head of a block can start one) and
//                                      their tile table
//   small    seg_window_sort_kernel    CTA b owns the segments that START in block b (at most 4095 reads): composite
//                                      (segment-in-window | S | position) sorted by an LSD radix sort entirely in shared
//                                      memory, ceil((bits(S) + bits(#segments)) / 8) passes, one global read + one write
//   big      seg_hist_kernel + seg_onesweep  one-sweep LSD passes over S only (4 passes for 25 bits instead of 7 for 53),
//                                      8-byte elements (S << idx_bits | read index) instead of key + index (24 B -> 16 B
//                                      per read per pass), all big segments batched in one launch per pass: a tile belongs
//                                      to one segment, its look-back stops at the segment's first tile
// Output is exactly what the generic sort produces (sorted one-word keys + read indices), so K3 does not change.
#pragma once
#include "common.cuh"
#include "pack.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"

#define SEG_BLK 2048u                 // block of positions; segments longer than this are "big"
#define SEG_WIN (2 * SEG_BLK)         // capacity of a window (segments starting in one block: < SEG_BLK + SEG_BLK reads)
#define SEG_THREADS RS_THREADS        // 512: shares rs_rank_tile / RsShared with the one-sweep kernel
#define SEG_WIN_ITEMS (SEG_WIN / SEG_THREADS)   // 8
#define SEG_POS_BITS 12               // position-in-window field of the composite
#define SEG_NONE 0xffffffffu
// Digit width of the segmented passes.  The passes are issue-bound on the ranking (one ballot per key bit), so what a wider
// digit buys is not fewer ballots but fewer passes' worth of everything else: 25 bits of S = 3 passes of 9,8,8 instead of 4,
// a window's 33 bits = 4 passes instead of 5 (round 1 measured a 9-bit pass 15 % dearer than an 8-bit one).
#define SEG_RB 9                      // widest digit; tables are laid out for it
#define SEG_RADIX (1 << SEG_RB)
// 9-bit digits only where they save a pass (25 bits: 3 instead of 4; 40 bits: 5 either way -> 8-bit digits)
static __host__ __device__ __forceinline__ int seg_pick_rb(int bits) { return (bits + 8) / 9 < (bits + 7) / 8 ? 9 : 8; }

// rs_rank_tile (radix_sort.cuh) with the digit width as a template parameter.  packed[j] = digit | (rank within (warp, digit)
// << RB), or 0xffffffff past the end.  On return whist[w][d] = number of keys with digit d in warp w's slice.
template <int ITEMS, bool FULL, int RB>
__device__ __forceinline__ void seg_rank_tile(u32 *packed, u32 (*whist)[1 << RB]) {
    const u32 w = threadIdx.x >> 5, lane = lane_id();
    for (u32 i = threadIdx.x; i < RS_WARPS * (1u << RB); i += SEG_THREADS) (&whist[0][0])[i] = 0;
    __syncthreads();
    const u32 lt = lanemask_lt();
#pragma unroll
    for (int j = 0; j < ITEMS; j++) {
        const u32 d = packed[j];
        const bool active = FULL || d != 0xffffffffu;
        u32 peers = FULL ? 0xffffffffu : __ballot_sync(0xffffffffu, active);
#pragma unroll
        for (int b = 0; b < RB; b++) {
            const bool bit = (d & (1u << b)) != 0u;
            const u32 m = __ballot_sync(0xffffffffu, bit);
            peers &= bit ? m : ~m;
        }
        const u32 leader = active ? (u32)__ffs(peers) - 1u : lane;
        u32 old = 0;
        if (active && lane == leader) { old = whist[w][d]; whist[w][d] = old + __popc(peers); }
        old = __shfl_sync(0xffffffffu, old, leader);
        if (active) packed[j] = d | ((old + __popc(peers & lt)) << RB);
        __syncwarp();
    }
}

struct SegBlock { u32 nheads, first, last, pad; };        // heads of segments that start in this block (positions), SEG_NONE = none
struct SegBig { u32 start, len, tile0, pad; };            // a big segment and the index of its first tile
struct SegPlanOut { u32 n_big, n_tiles, unsorted, pad; };

// CTA reduction of a block's table entry: sum, min (first), max (last; SEG_NONE = all ones must lose: map to 0 via +1 trick)
__device__ __forceinline__ void seg_block_reduce(u32 cnt, u32 first, u32 last, u32 bad, SegBlock *__restrict__ blk, SegPlanOut *plan) {
    __shared__ u32 s_cnt[8], s_first[8], s_last[8], s_bad;
    if (threadIdx.x == 0) s_bad = 0;
    u32 lastp = last == SEG_NONE ? 0u : last + 1u;           // 0 = none
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
        lastp = max(lastp, __shfl_xor_sync(0xffffffffu, lastp, o));
        bad |= __shfl_xor_sync(0xffffffffu, bad, o);
    }
    __syncthreads();
    if (lane_id() == 0) { s_cnt[threadIdx.x >> 5] = cnt; s_first[threadIdx.x >> 5] = first; s_last[threadIdx.x >> 5] = lastp; if (bad) s_bad = 1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 c = 0, f = SEG_NONE, l = 0;
        for (int w = 0; w < 8; w++) { c += s_cnt[w]; f = min(f, s_first[w]); l = max(l, s_last[w]); }
        blk[blockIdx.x] = SegBlock{c, f, l ? l - 1u : SEG_NONE, 0u};
        if (s_bad) plan->unsorted = 1;
    }
}

// ---- plan, step 1: block table ----
__global__ void __launch_bounds__(256) seg_block_summary_kernel(const u64 *__restrict__ key, u64 n, int sbits, SegBlock *__restrict__ blk,
                                                                SegPlanOut *plan) {
    const u64 base = (u64)blockIdx.x * SEG_BLK + (u64)threadIdx.x * 8;
    u64 k[8];
    u64 prev = 0;
    if (base < n) {
        if (base + 8 <= n) {
            const ulonglong2 *kv = reinterpret_cast<const ulonglong2 *>(key + base);
#pragma unroll
            for (int q = 0; q < 4; q++) { const ulonglong2 t = __ldg(kv + q); k[2 * q] = t.x; k[2 * q + 1] = t.y; }
        } else {
#pragma unroll
            for (int j = 0; j < 8; j++) k[j] = base + j < n ? key[base + j] : 0;
        }
    }
    prev = __shfl_up_sync(0xffffffffu, base < n ? k[7] : 0ull, 1);
    if (lane_id() == 0) prev = (base > 0 && base < n) ? key[base - 1] : 0;
    u32 cnt = 0, first = SEG_NONE, last = SEG_NONE, bad = 0;
    if (base < n) {
        u64 p = prev >> sbits;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const u64 i = base + j;
            if (i < n) {
                const u64 c = k[j] >> sbits;
                if (i == 0 || c != p) { cnt++; if (first == SEG_NONE) first = (u32)i; last = (u32)i; }
                if (i != 0 && c < p) bad = 1;
                p = c;
            }
        }
    }
    seg_block_reduce(cnt, first, last, bad, blk, plan);
}

// K1b and plan step 1 in ONE read of the inputs: the one-word keys of a block of 2048 reads are built (what
// build_keys_kernel<1, LIN> writes) and the block's table entry is filled from the keys while they are in registers, so the
// plan no longer re-reads 8 B per read.  Reads are taken striped (coalesced scalar loads: no alignment demands on the
// caller's arrays) and the keys staged in shared memory for the comparison with the predecessor.  (First form: predecessor
// by shuffle, lane 0 rebuilding its own from the inputs — eight divergent reload bubbles per warp made the kernel 0.33 ms
// slower on C5 than the two kernels it replaced; profiles/r3a_ab_C5.jsonl.)
template <bool LIN>
__device__ __forceinline__ u64 seg_key_of(u64 i, const i32 *__restrict__ tid, const i64 *__restrict__ pos, const u8 *__restrict__ rev,
                                          const i64 *__restrict__ tlen, const u64 *__restrict__ umi2, const u32 *__restrict__ nmask, const KeyLayout &lay) {
    u64 bucket;
    if (LIN) { const u32 t = (u32)(tid[i] - lay.tid_min); bucket = ((__ldg(lay.lin_off + t) + (u64)(pos[i] - __ldg(lay.lin_pmin + t))) << 1) | (rev[i] ? 1u : 0u); }
    else bucket = ((u64)(u32)(tid[i] - lay.tid_min) << (lay.pos_bits + 1)) | ((u64)(pos[i] - lay.pos_min) << 1) | (rev[i] ? 1u : 0u);
    if (lay.tlen_bits) bucket = (bucket << lay.tlen_bits) | (u64)(tlen[i] - lay.tlen_min);
    const u64 code = umi_sort_code(umi2[i], lay.has_n ? nmask[i] : 0u, lay.umi_len, lay.has_n);
    return (lay.umi_bits < 64 ? bucket << lay.umi_bits : 0) | code;
}

template <bool LIN>
__global__ void __launch_bounds__(256) build_keys_summary_kernel(u64 n, const i32 *__restrict__ tid, const i64 *__restrict__ pos, const u8 *__restrict__ rev,
                                                                 const i64 *__restrict__ tlen, const u64 *__restrict__ umi2, const u32 *__restrict__ nmask,
                                                                 KeyLayout lay, u64 *__restrict__ k0, int sbits, SegBlock *__restrict__ blk, SegPlanOut *plan) {
    // skey[1 + e] = key of the block's read e, skey[0] = key of the read before the block (the only one built twice)
    __shared__ u64 skey[SEG_BLK + 1];
    const u64 blk_base = (u64)blockIdx.x * SEG_BLK;
    if (threadIdx.x == 0) skey[0] = blk_base > 0 ? seg_key_of<LIN>(blk_base - 1, tid, pos, rev, tlen, umi2, nmask, lay) : 0ull;
#pragma unroll
    for (int j = 0; j < (int)(SEG_BLK / 256); j++) {
        const u32 e = (u32)j * 256 + threadIdx.x;
        const u64 i = blk_base + e;
        if (i < n) { const u64 key = seg_key_of<LIN>(i, tid, pos, rev, tlen, umi2, nmask, lay); k0[i] = key; skey[1 + e] = key; }
    }
    __syncthreads();
    u32 cnt = 0, first = SEG_NONE, last = SEG_NONE, bad = 0;
#pragma unroll
    for (int j = 0; j < (int)(SEG_BLK / 256); j++) {
        const u32 e = (u32)j * 256 + threadIdx.x;
        const u64 i = blk_base + e;
        if (i < n) {
            const u64 c = skey[1 + e] >> sbits, p = skey[e] >> sbits;
            if (i == 0 || c != p) { cnt++; if (first == SEG_NONE) first = (u32)i; last = (u32)i; }      // i ascends with j
            if (i != 0 && c < p) bad = 1;
        }
    }
    seg_block_reduce(cnt, first, last, bad, blk, plan);
}

// ---- plan, step 2 (one CTA): next head after every block, big segments, tile table ----
// next[b] = position of the first head in a block > b (n if none).  The last head of block b starts a big segment iff
// next[b] - last[b] > SEG_BLK.  big_of_blk[b] = 1 in that case (the window kernel then stops before it).
#define SEG_PLAN_PER 4          // block-table entries per thread and round
__global__ void __launch_bounds__(1024) seg_plan_kernel(const SegBlock *__restrict__ blk, u32 n_blk, u64 n, u32 tile, u32 *__restrict__ next_head,
                                                        u8 *__restrict__ big_of_blk, SegBig *__restrict__ big, SegPlanOut *plan) {
    __shared__ u32 sm[1024 / 32 + 1];
    __shared__ u32 s_carry;
    constexpr int PER = SEG_PLAN_PER;
    constexpr long long CHUNK = 1024LL * PER;
    // suffix minimum of `first`, walking the block table backwards: thread t of a round owns the PER entries below
    // hi - PER * t (descending), so "entries after mine" = lower threads + the rounds already done
    if (threadIdx.x == 0) s_carry = (u32)n;
    __syncthreads();
    for (long long hi = (long long)n_blk; hi > 0; hi -= CHUNK) {
        u32 v[PER];
        u32 agg = SEG_NONE;
#pragma unroll
        for (int q = 0; q < PER; q++) {
            const long long b = hi - 1 - (long long)threadIdx.x * PER - q;
            v[q] = (b >= 0 && blk[b].first != SEG_NONE) ? blk[b].first : SEG_NONE;
            agg = min(agg, v[q]);
        }
        u32 inc = agg;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_up_sync(0xffffffffu, inc, o); if (lane_id() >= (u32)o) inc = min(inc, t); }
        if (lane_id() == 31) sm[threadIdx.x >> 5] = inc;
        __syncthreads();
        if (threadIdx.x < 32) {
            u32 sv = sm[threadIdx.x];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_up_sync(0xffffffffu, sv, o); if (lane_id() >= (u32)o) sv = min(sv, t); }
            sm[threadIdx.x] = sv;                                      // inclusive over warps
        }
        __syncthreads();
        u32 ex = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane_id() == 0) ex = SEG_NONE;
        const u32 w = threadIdx.x >> 5;
        if (w > 0) ex = min(ex, sm[w - 1]);
        const u32 carry = s_carry;
        ex = min(ex, carry);                                           // heads in later rounds / n
        u32 run = ex;
#pragma unroll
        for (int q = 0; q < PER; q++) {
            const long long b = hi - 1 - (long long)threadIdx.x * PER - q;
            if (b >= 0) next_head[b] = run;
            run = min(run, v[q]);
        }
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = run;                        // minimum over everything from this round's lowest entry up
        __syncthreads();
    }
    // big segments in block order, with their first tile (exclusive prefix sum of tile counts)
    u32 carry_big = 0, carry_tiles = 0;
    for (u32 lo = 0; lo < n_blk; lo += (u32)CHUNK) {
        u32 start[PER], len[PER], nt[PER];
        u32 cnt_b = 0, cnt_t = 0;
#pragma unroll
        for (int q = 0; q < PER; q++) {
            const u32 b = lo + threadIdx.x * PER + q;
            start[q] = 0; len[q] = 0; nt[q] = 0;
            u32 is_big = 0;
            if (b < n_blk && blk[b].last != SEG_NONE) {
                start[q] = blk[b].last; len[q] = next_head[b] - start[q];
                is_big = len[q] > SEG_BLK ? 1u : 0u;
            }
            if (b < n_blk) big_of_blk[b] = (u8)is_big;
            if (is_big) { nt[q] = (len[q] + tile - 1) / tile; cnt_b++; cnt_t += nt[q]; } else len[q] = 0;
        }
        u32 tot_b, tot_t;
        u32 ex_b = block_exclusive_scan<u32, 1024>(cnt_b, sm, &tot_b);
        u32 ex_t = block_exclusive_scan<u32, 1024>(cnt_t, sm, &tot_t);
#pragma unroll
        for (int q = 0; q < PER; q++) {
            if (len[q]) { big[carry_big + ex_b] = SegBig{start[q], len[q], carry_tiles + ex_t, 0u}; ex_b++; ex_t += nt[q]; }
        }
        carry_big += tot_b; carry_tiles += tot_t;
    }
    if (threadIdx.x == 0) { plan->n_big = carry_big; plan->n_tiles = carry_tiles; }
}

// ---- small segments: one window per block, sorted in shared memory ----
struct SegWinShared {
    u32 whist[RS_WARPS][SEG_RADIX];
    u32 slocal[SEG_RADIX];
    u32 sscan[SEG_THREADS / 32 + 1];
};
static_assert(SEG_RADIX <= SEG_THREADS, "one thread per digit in the per-digit steps");

// one LSD pass over the elements held in registers (slot e = warp * ITEMS * 32 + j * 32 + lane); result back in x[], in slot order
template <int ITEMS, int RB>
__device__ __forceinline__ void seg_local_pass(u64 (&x)[ITEMS], u32 count, int shift, u32 mask, SegWinShared &S, u64 *sbuf) {
    constexpr u32 RADIX = 1u << RB;
    u32 (*whist)[RADIX] = reinterpret_cast<u32 (*)[RADIX]>(&S.whist[0][0]);
    const u32 w = threadIdx.x >> 5, lane = lane_id();
    u32 packed[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; j++) {
        const u32 e = (w * ITEMS + j) * 32 + lane;
        packed[j] = e < count ? ((u32)(x[j] >> shift) & mask) : 0xffffffffu;
    }
    seg_rank_tile<ITEMS, false, RB>(packed, whist);
    __syncthreads();
    u32 total = 0;
    if (threadIdx.x < RADIX) {
#pragma unroll
        for (int ww = 0; ww < RS_WARPS; ww++) { const u32 t = whist[ww][threadIdx.x]; whist[ww][threadIdx.x] = total; total += t; }
    }
    u32 tot_all;
    const u32 lstart = block_exclusive_scan<u32, SEG_THREADS>(total, S.sscan, &tot_all);
    if (threadIdx.x < RADIX) S.slocal[threadIdx.x] = lstart;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < ITEMS; j++) {
        if (packed[j] != 0xffffffffu) {
            const u32 d = packed[j] & (RADIX - 1), r = packed[j] >> RB;
            sbuf[S.slocal[d] + whist[w][d] + r] = x[j];
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < ITEMS; j++) {
        const u32 e = (w * ITEMS + j) * 32 + lane;
        x[j] = e < count ? sbuf[e] : 0;
    }
    __syncthreads();
}

// Windows are mostly far from full (a block whose tail lies inside a big segment, sparse loci): the element count picks how many
// slots a thread owns (1, 2, 4 or 8), so a window of 500 reads ranks 512 slots per pass, not 4096.
template <int ITEMS, int RB>
__device__ __forceinline__ void seg_window_body(SegWinShared &S, u64 *sbuf, const u64 *__restrict__ key_in, u32 lo, u32 count, int sbits, int segbits,
                                                u64 *__restrict__ key_out, u32 *__restrict__ idx_out) {
    const u64 smask = (1ull << sbits) - 1;
    const u32 w = threadIdx.x >> 5, lane = lane_id();
    u64 x[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; j++) {
        const u32 e = (w * ITEMS + j) * 32 + lane;
        x[j] = e < count ? sbuf[e] : 0;
    }
    __syncthreads();
    const int bits = sbits + segbits;
    // bits spread evenly over the passes (33 bits -> 9,8,8,8)
    const int npass = (bits + RB - 1) / RB;
    for (int p = 0, done = 0; p < npass; p++) {
        const int b = (bits - done + (npass - p) - 1) / (npass - p);
        seg_local_pass<ITEMS, RB>(x, count, SEG_POS_BITS + done, (1u << b) - 1u, S, sbuf);
        done += b;
    }
    // slot e now holds the element of sorted position lo + e; P comes from the (unsorted) input at the same position
#pragma unroll
    for (int j = 0; j < ITEMS; j++) {
        const u32 e = (w * ITEMS + j) * 32 + lane;
        if (e < count) {
            const u64 i = (u64)lo + e;
            key_out[i] = (key_in[i] & ~smask) | ((x[j] >> SEG_POS_BITS) & smask);
            idx_out[i] = lo + (u32)(x[j] & ((1u << SEG_POS_BITS) - 1u));
        }
    }
}

__global__ void __launch_bounds__(SEG_THREADS, 2) seg_window_sort_kernel(const u64 *__restrict__ key_in, u64 n, int sbits, const SegBlock *__restrict__ blk,
                                                                         const u32 *__restrict__ next_head, const u8 *__restrict__ big_of_blk,
                                                                         u64 *__restrict__ key_out, u32 *__restrict__ idx_out) {
    __shared__ SegWinShared S;
    extern __shared__ __align__(16) unsigned char seg_win_dyn[];         // u64 sbuf[SEG_WIN] (static + this exceeds 48 KB)
    u64 *sbuf = reinterpret_cast<u64 *>(seg_win_dyn);
    const SegBlock sb = blk[blockIdx.x];
    if (sb.nheads == 0) return;                                   // the block lies inside a segment that started earlier
    const u32 lo = sb.first;
    const u32 hi = big_of_blk[blockIdx.x] ? sb.last : next_head[blockIdx.x];
    const u32 count = hi - lo;                                    // < SEG_WIN
    if (count == 0) return;
    const u64 smask = (1ull << sbits) - 1;
    // blocked load (8 consecutive reads per thread) for the head scan
    const u32 e0 = threadIdx.x * SEG_WIN_ITEMS;
    u64 k[SEG_WIN_ITEMS];
#pragma unroll
    for (int j = 0; j < SEG_WIN_ITEMS; j++) k[j] = e0 + j < count ? key_in[(u64)lo + e0 + j] : 0;
    u64 prev = __shfl_up_sync(0xffffffffu, k[SEG_WIN_ITEMS - 1], 1);
    if (lane_id() == 0) prev = (e0 > 0 && e0 < count) ? key_in[(u64)lo + e0 - 1] : 0;
    u32 hm = 0;
    {
        u64 p = prev >> sbits;
#pragma unroll
        for (int j = 0; j < SEG_WIN_ITEMS; j++) {
            const u64 c = k[j] >> sbits;
            if (e0 + j < count && e0 + j > 0 && c != p) hm |= 1u << j;
            p = c;
        }
    }
    u32 nseg_m1;
    u32 seg = block_exclusive_scan<u32, SEG_THREADS>((u32)__popc(hm), S.sscan, &nseg_m1);      // segment-in-window of the element before e0
    // composite = (segment | S | position in window)
#pragma unroll
    for (int j = 0; j < SEG_WIN_ITEMS; j++) {
        seg += (hm >> j) & 1u;
        if (e0 + j < count) sbuf[e0 + j] = ((((u64)seg << sbits) | (k[j] & smask)) << SEG_POS_BITS) | (u64)(e0 + j);
    }
    __syncthreads();
    int segbits = 0;
    while ((nseg_m1 >> segbits) != 0) segbits++;
#define SEG_WIN_GO(IT) do { if (rb9) seg_window_body<IT, 9>(S, sbuf, key_in, lo, count, sbits, segbits, key_out, idx_out); \
                            else     seg_window_body<IT, 8>(S, sbuf, key_in, lo, count, sbits, segbits, key_out, idx_out); } while (0)
    const bool rb9 = seg_pick_rb(sbits + segbits) == 9;
    if (count <= SEG_THREADS)          SEG_WIN_GO(1);
    else if (count <= 2 * SEG_THREADS) SEG_WIN_GO(2);
    else if (count <= 4 * SEG_THREADS) SEG_WIN_GO(4);
    else                               SEG_WIN_GO(8);
#undef SEG_WIN_GO
}

// ---- big segments: batched one-sweep passes over S ----
#define SEG_ITEMS 12
#define SEG_TILE (SEG_THREADS * SEG_ITEMS)     // 6144
#define SEG_MAX_PASSES 6
// the bits of S spread evenly over the passes (25 bits -> 7,6,6,6: a pass of one leftover bit would cost a full pass)
struct SegPasses { int npass, rb; int shift[SEG_MAX_PASSES]; int bits[SEG_MAX_PASSES]; };
static inline SegPasses seg_passes(int sbits) {
    SegPasses sp; sp.rb = seg_pick_rb(sbits); sp.npass = (sbits + sp.rb - 1) / sp.rb;
    int done = 0;
    for (int i = 0; i < sp.npass; i++) { const int b = (sbits - done + (sp.npass - i) - 1) / (sp.npass - i); sp.shift[i] = done; sp.bits[i] = b; done += b; }
    return sp;
}

__device__ __forceinline__ u32 seg_find_big(const SegBig *__restrict__ big, u32 n_big, u32 tile) {
    u32 lo = 0, hi = n_big;                    // last segment with tile0 <= tile
    while (hi - lo > 1) { const u32 mid = (lo + hi) >> 1; if (big[mid].tile0 <= tile) lo = mid; else hi = mid; }
    return lo;
}

// digit histograms of every pass for every big segment: hist[seg][pass][digit]
__global__ void __launch_bounds__(SEG_THREADS) seg_hist_kernel(const u64 *__restrict__ key_in, int sbits, SegPasses sp, const SegBig *__restrict__ big,
                                                               u32 n_big, u32 n_tiles, u32 *__restrict__ hist) {
    const int npass = sp.npass;
    __shared__ u32 sh[SEG_MAX_PASSES * SEG_RADIX];
    __shared__ u32 s_seg;
    for (u32 tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        __syncthreads();
        if (threadIdx.x == 0) s_seg = seg_find_big(big, n_big, tile);
        for (u32 i = threadIdx.x; i < (u32)npass * SEG_RADIX; i += SEG_THREADS) sh[i] = 0;
        __syncthreads();
        const SegBig sg = big[s_seg];
        const u32 off = (tile - sg.tile0) * SEG_TILE, cnt = min((u32)SEG_TILE, sg.len - off);
        const u64 smask = (1ull << sbits) - 1;
        for (u32 e = threadIdx.x; e < cnt; e += SEG_THREADS) {
            const u64 s = key_in[(u64)sg.start + off + e] & smask;
            for (int p = 0; p < npass; p++) atomicAdd(&sh[p * SEG_RADIX + ((u32)(s >> sp.shift[p]) & ((1u << sp.bits[p]) - 1u))], 1u);
        }
        __syncthreads();
        for (u32 i = threadIdx.x; i < (u32)npass * SEG_RADIX; i += SEG_THREADS)
            if (sh[i]) atomicAdd(&hist[(u64)s_seg * npass * SEG_RADIX + i], sh[i]);
    }
}
// exclusive scan of each (segment, pass) histogram: one CTA of SEG_RADIX threads per row
__global__ void __launch_bounds__(SEG_RADIX) seg_digit_starts_kernel(u32 *hist, u32 n_rows) {
    __shared__ u32 sm[SEG_RADIX / 32 + 1];
    for (u32 row = blockIdx.x; row < n_rows; row += gridDim.x) {
        const u32 v = hist[(u64)row * SEG_RADIX + threadIdx.x];
        u32 tot;
        const u32 ex = block_exclusive_scan<u32, SEG_RADIX>(v, sm, &tot);
        hist[(u64)row * SEG_RADIX + threadIdx.x] = ex;
    }
}

// One pass.  FIRST: elements are built from key_in (S << ib | read index); LAST: writes sorted keys and indices.
// FULL = the tile has all TILE elements (every tile of a segment but its last): no bounds checks, no activity ballots.
struct SegPassArgs {
    const u64 *key_in; const u64 *w_in; u64 *w_out; u64 *key_out; u32 *idx_out;
    int sbits, ib, pass; SegPasses sp;
    const SegBig *big; u32 n_big, n_tiles; const u32 *hist; unsigned long long *tile_state; u32 *ticket, *err;
};
struct SegShared {
    u32 whist[RS_WARPS][SEG_RADIX];
    u32 sbase[SEG_RADIX], slocal[SEG_RADIX];
    u32 sscan[SEG_THREADS / 32 + 1];
    u32 s_tile;
};
template <int ITEMS, bool FULL, int RB>
__device__ __forceinline__ void seg_onesweep_body(const SegPassArgs &a, SegShared &S, u64 *selem, u32 tile, u32 seg_i, const SegBig sg, u32 off, u32 tile_count) {
    constexpr u32 TILE = SEG_THREADS * ITEMS;
    const u64 tile_base = (u64)sg.start + off;
    const int npass = a.sp.npass, pass = a.pass, ib = a.ib;
    const bool first = pass == 0, last = pass == npass - 1;
    const u64 smask = (1ull << a.sbits) - 1;
    const int sh = ib + a.sp.shift[pass];
    const u32 mask = (1u << a.sp.bits[pass]) - 1u;
    const u32 *digit_start = a.hist + ((u64)seg_i * npass + pass) * SEG_RADIX;
    constexpr u32 RADIX = 1u << RB;                           // digits of this build; table rows stay SEG_RADIX wide
    u32 (*whist)[RADIX] = reinterpret_cast<u32 (*)[RADIX]>(&S.whist[0][0]);
    u32 *sbase = S.sbase, *slocal = S.slocal, *sscan = S.sscan;
    const u32 w = threadIdx.x >> 5, lane = lane_id();
    u64 x[ITEMS];
    u32 packed[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; j++) {
        const u32 e = (w * ITEMS + j) * 32 + lane;
        const u64 i = tile_base + e;
        if (FULL || e < tile_count) x[j] = first ? (((a.key_in[i] & smask) << ib) | i) : a.w_in[i]; else x[j] = 0;
    }
#pragma unroll
    for (int j = 0; j < ITEMS; j++) {
        const u32 e = (w * ITEMS + j) * 32 + lane;
        packed[j] = (FULL || e < tile_count) ? ((u32)(x[j] >> sh) & mask) : 0xffffffffu;
    }
    seg_rank_tile<ITEMS, FULL, RB>(packed, whist);
    __syncthreads();
    u32 total = 0;
    if (threadIdx.x < RADIX) {
        const u32 d = threadIdx.x;
#pragma unroll
        for (int ww = 0; ww < RS_WARPS; ww++) { const u32 t = whist[ww][d]; whist[ww][d] = total; total += t; }
        atomicExch(a.tile_state + (u64)tile * SEG_RADIX + d, (tile == sg.tile0 ? RS_FLAG_PREFIX : RS_FLAG_AGG) | (unsigned long long)total);
    }
    u32 tot_all;
    const u32 lstart = block_exclusive_scan<u32, SEG_THREADS>(total, sscan, &tot_all);
    if (threadIdx.x < RADIX) slocal[threadIdx.x] = lstart;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < ITEMS; j++) {
        if (FULL || packed[j] != 0xffffffffu) {
            const u32 d = packed[j] & (RADIX - 1), r = packed[j] >> RB;
            selem[slocal[d] + whist[w][d] + r] = x[j];
        }
    }
    if (threadIdx.x < RADIX) {
        const u32 d = threadIdx.x;
        unsigned long long *mine = a.tile_state + (u64)tile * SEG_RADIX + d;
        u64 excl = 0;
        if (tile != sg.tile0) {
            u32 p = tile;                      // predecessors [tile0, p) of this segment are still to be accounted for
            u32 spins = 0;
            bool done = false;
            while (!done && p > sg.tile0) {
                unsigned long long v[RS_LB];
#pragma unroll
                for (int i = 0; i < RS_LB; i++) {
                    const volatile unsigned long long *prev = a.tile_state + (u64)(p > sg.tile0 + (u32)i ? p - 1 - i : sg.tile0) * SEG_RADIX + d;
                    v[i] = *prev;
                }
#pragma unroll
                for (int i = 0; i < RS_LB; i++) {
                    if (done || p == sg.tile0) break;
                    if ((v[i] & RS_FLAG_MASK) == 0) {
                        if (++spins > (1u << 24)) { a.err[0] = 1; done = true; }
                        if (i == 0) __nanosleep(40);
                        break;
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
    const u64 pseg = last ? (a.key_in[sg.start] & ~smask) : 0ull;      // every read of the segment has this (contig, position)
    const u64 imask = (1ull << ib) - 1;
#pragma unroll
    for (int j = 0; j < ITEMS; j++) {
        const u32 i = threadIdx.x + j * SEG_THREADS;
        if (FULL || i < tile_count) {
            const u64 v = selem[i];
            const u32 d = (u32)(v >> sh) & mask;
            const u64 pos = (u64)sg.start + sbase[d] + (i - slocal[d]);
            if (last) { a.key_out[pos] = pseg | (v >> ib); a.idx_out[pos] = (u32)(v & imask); }
            else a.w_out[pos] = v;
        }
    }
}

template <int ITEMS, int RB>
__global__ void __launch_bounds__(SEG_THREADS, 2) seg_onesweep(SegPassArgs a) {
    constexpr u32 TILE = SEG_THREADS * ITEMS;
    extern __shared__ __align__(16) unsigned char seg_dyn[];            // u64 selem[TILE]
    u64 *selem = reinterpret_cast<u64 *>(seg_dyn);
    __shared__ SegShared S;
    __shared__ u32 s_seg;
    if (threadIdx.x == 0) { const u32 t = atomicAdd(a.ticket, 1u); S.s_tile = t; s_seg = t < a.n_tiles ? seg_find_big(a.big, a.n_big, t) : 0u; }
    __syncthreads();
    const u32 tile = S.s_tile;
    if (tile >= a.n_tiles) return;
    const u32 seg_i = s_seg;
    const SegBig sg = a.big[seg_i];
    const u32 off = (tile - sg.tile0) * TILE, tile_count = min(TILE, sg.len - off);
    if (tile_count == TILE) seg_onesweep_body<ITEMS, true, RB>(a, S, selem, tile, seg_i, sg, off, tile_count);
    else                    seg_onesweep_body<ITEMS, false, RB>(a, S, selem, tile, seg_i, sg, off, tile_count);
}
