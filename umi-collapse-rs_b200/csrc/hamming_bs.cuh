// hamming_bs.cuh — K5 hamming_neighbours, bit-sliced one-hot kernel in its dense shared-memory TILE form.
// Used when the work is dense (culling disabled, or culling keeps > 30 % of the scheduled pair space) and with
// UMIGPU_FLAG_KERNEL_TILES; the sparse block-pair form (hamming_blocks.cuh) is the default after culling.
//
// Same contract as hamming_tiles_direct (hamming.cuh) but 32 column UMIs are compared per
// instruction.  For a tile of up to 2048 columns the CTA first builds, in shared memory, one-hot
// match words
//        eq[g][j][x]  (u32)   bit c = "column 32g+c has letter x at position j"      (warp ballots)
// Each thread then owns one row UMI at a time.  For position j it loads the word selected by its
// own letter a_j — the 32 columns that MATCH the row at j — and feeds a bit-sliced saturating
// mismatch counter:   m1 &= w;  m2 = w ? m2 : m1;  [m3 = w ? m3 : m2; ...]   (one LOP3 each).
// After umi_len positions, bit c of m_{k+1} says dist(row, column c) <= k.  That is k+1 LOP3 and one
// broadcast LDS per base per 32 pairs (0.75 ALU ops per pair at 12 nt, k = 1) instead of ~4 per
// pair for the direct XOR+popcount form; N needs no extra work (it is just a fifth letter).
// Words of four consecutive column groups are interleaved so one LDS.128 feeds four counters.
// Hits are rare (a handful per UMI) and go through the cold path record_hit().
#pragma once
#include "common.cuh"
#include "hamming.cuh"

#define BS_THREADS 256
#define BS_G4      (HT_COLS / 128)     // blocks of 4 column groups (128 columns) per tile

template <int LP, int K, bool HASN>
__global__ void __launch_bounds__(BS_THREADS) hamming_tiles_bs(
    const TileItem *__restrict__ items, u32 n_items, const uint2 *__restrict__ planes, const u32 *__restrict__ nplane,
    const u32 *__restrict__ bsum, int L, int cull, EdgeSink es, u32 *work_counter, unsigned long long *pairs_eval) {
    constexpr int XS = HASN ? 8 : 4;          // letter slots per position
    constexpr int NLET = HASN ? 5 : 4;
    constexpr int PASSES = HT_ROWS / BS_THREADS;
    extern __shared__ __align__(16) uint4 eq4[];   // [g4][LP][XS] x (4 groups)
    __shared__ __align__(16) u32 sset[BS_G4 * 8];  // per 128-column block: letter-set masks (bit j = letter x occurs at position j)
    __shared__ u32 s_item, s_need;
    const u32 lane = lane_id(), warp = threadIdx.x >> 5;
    const u32 lmask = L >= 32 ? 0xffffffffu : ((1u << L) - 1);
    u64 evaluated = 0;

    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) { s_item = atomicAdd(work_counter, 1u); s_need = 0; }
        __syncthreads();
        const u32 w_item = s_item;
        if (w_item >= n_items) break;
        const TileItem it = items[w_item];
        const u32 col_cnt = item_col_cnt(it), row_cnt = item_row_cnt(it);
        const bool diag = item_diag(it);
        const u32 g4_cnt = (col_cnt + 127) >> 7;
        const u32 all_g4 = (g4_cnt >= 32 ? 0xffffffffu : ((1u << g4_cnt) - 1));

        // ---- phase 0: letter sets of the column tile's 128-column blocks (precomputed per tile) ----
        if (cull) {
            if (threadIdx.x < g4_cnt * 8) sset[threadIdx.x] = bsum[(u64)it.col_blk0 * 8 + threadIdx.x];
            __syncthreads();
        }
        // ---- phase 1: which (32-row warp slice) x (128-column block) pairs can contain a neighbour at all ----
        // A warp's 32 rows are consecutive sorted UMIs, so their letter sets are small; a block pair with more
        // than K positions of disjoint letter sets cannot contain a pair within K.
        u32 mask[PASSES];
        u32 need = 0;
#pragma unroll
        for (int ps = 0; ps < PASSES; ps++) {
            mask[ps] = 0;
            const u32 rbase = ps * BS_THREADS;
            if (rbase + warp * 32 >= row_cnt) continue;        // warp-uniform
            if (!cull) { mask[ps] = all_g4; need |= all_g4; continue; }
            const u32 gi = rbase + threadIdx.x;
            const bool valid = gi < row_cnt;
            const uint2 rp = valid ? planes[it.row_start + gi] : make_uint2(0u, 0u);
            const u32 rn = (HASN && valid) ? nplane[it.row_start + gi] : 0u;
            u32 oh[5], rs[5] = {0u, 0u, 0u, 0u, 0u};
            onehot_planes(rp, rn, valid ? lmask : 0u, oh);
#pragma unroll
            for (int x = 0; x < NLET; x++) rs[x] = __reduce_or_sync(0xffffffffu, oh[x]);
            const u32 row_lo = rbase + warp * 32;
            u32 mk = 0;
            for (u32 g4 = 0; g4 < g4_cnt; g4++) {
                if (diag && g4 * 128 + 127 <= row_lo) continue;   // diagonal tile: only column > row is wanted
                const uint4 s = *reinterpret_cast<const uint4 *>(&sset[g4 * 8]);
                u32 t = (rs[0] & s.x) | (rs[1] & s.y) | (rs[2] & s.z) | (rs[3] & s.w);
                if (HASN) t |= rs[4] & sset[g4 * 8 + 4];
                if (__popc(~t & lmask) <= K) mk |= 1u << g4;
            }
            mask[ps] = mk; need |= mk;
        }
        if (lane == 0 && need) atomicOr(&s_need, need);
        __syncthreads();
        const u32 needed = s_need;
        if (needed == 0) continue;                              // nothing in this tile pair survives the exact cull

        // ---- phase 2: one-hot match words of the needed column blocks ----
        for (u32 g = warp; g < g4_cnt * 4; g += BS_THREADS / 32) {
            if (!((needed >> (g >> 2)) & 1u)) continue;         // warp-uniform
            u32 c = g * 32 + lane;
            bool valid = c < col_cnt;
            uint2 p = valid ? planes[it.col_start + c] : make_uint2(0u, 0u);
            u32 pn = (HASN && valid) ? nplane[it.col_start + c] : 0u;
            u32 *dst = reinterpret_cast<u32 *>(eq4) + ((size_t)(g >> 2) * LP * XS) * 4 + (g & 3);
#pragma unroll
            for (int j = 0; j < LP; j++) {
                u32 letter = ((p.y >> j) & 1u) * 2u + ((p.x >> j) & 1u);
                if (HASN && ((pn >> j) & 1u)) letter = 4u;
                if (!valid) letter = 15u;                 // padding column: matches no letter at any real position
                u32 v = 0;
#pragma unroll
                for (int x = 0; x < NLET; x++) {
                    u32 b = __ballot_sync(0xffffffffu, letter == (u32)x);
                    if (lane == (u32)x) v = b;
                }
                if (j >= L) v = 0xffffffffu;              // positions beyond umi_len always match
                if (lane < (u32)XS) dst[(j * XS + lane) * 4] = v;
            }
        }
        __syncthreads();

        // ---- phase 3: evaluate the surviving blocks; each thread owns one row UMI per pass ----
#pragma unroll
        for (int ps = 0; ps < PASSES; ps++) {
            u32 mk = mask[ps];
            if (mk == 0) continue;                              // warp-uniform
            const u32 gi = ps * BS_THREADS + threadIdx.x;
            const bool valid = gi < row_cnt;
            const u32 nvalid = __popc(__ballot_sync(0xffffffffu, valid));
            const uint2 rp = valid ? planes[it.row_start + gi] : make_uint2(0u, 0u);
            const u32 rn = (HASN && valid) ? nplane[it.row_start + gi] : 0u;
            u32 off[LP];                                   // byte offset of the row's letter slot at position j
#pragma unroll
            for (int j = 0; j < LP; j++) {
                u32 letter = ((rp.y >> j) & 1u) * 2u + ((rp.x >> j) & 1u);
                if (HASN && ((rn >> j) & 1u)) letter = 4u;
                off[j] = (u32)(j * XS + letter) * 16u;
            }
            while (mk) {
                const u32 g4 = __ffs(mk) - 1; mk &= mk - 1;
                evaluated += (u64)nvalid * min(128u, col_cnt - g4 * 128);
                if (!valid) continue;
                const char *base = reinterpret_cast<const char *>(eq4) + (size_t)g4 * (LP * XS * 16);
                uint4 m1 = make_uint4(~0u, ~0u, ~0u, ~0u), m2 = m1, m3 = m1, m4 = m1;
#pragma unroll
                for (int j = 0; j < LP; j++) {
                    const uint4 w = *reinterpret_cast<const uint4 *>(base + off[j]);
                    if (K >= 3) { m4.x = (w.x & m4.x) | (~w.x & m3.x); m4.y = (w.y & m4.y) | (~w.y & m3.y);
                                  m4.z = (w.z & m4.z) | (~w.z & m3.z); m4.w = (w.w & m4.w) | (~w.w & m3.w); }
                    if (K >= 2) { m3.x = (w.x & m3.x) | (~w.x & m2.x); m3.y = (w.y & m3.y) | (~w.y & m2.y);
                                  m3.z = (w.z & m3.z) | (~w.z & m2.z); m3.w = (w.w & m3.w) | (~w.w & m2.w); }
                    m2.x = (w.x & m2.x) | (~w.x & m1.x); m2.y = (w.y & m2.y) | (~w.y & m1.y);
                    m2.z = (w.z & m2.z) | (~w.z & m1.z); m2.w = (w.w & m2.w) | (~w.w & m1.w);
                    m1.x &= w.x; m1.y &= w.y; m1.z &= w.z; m1.w &= w.w;
                }
                const uint4 h = K == 1 ? m2 : (K == 2 ? m3 : m4);
                if (h.x | h.y | h.z | h.w) {
                    const u32 a = it.row_start + gi;
                    const u32 hv[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        u32 bits = hv[i];
                        while (bits) {
                            u32 b = __ffs(bits) - 1; bits &= bits - 1;
                            u32 c = (g4 * 4 + i) * 32 + b;
                            if (c < col_cnt) {
                                u32 bb = it.col_start + c;
                                if (!diag || a < bb) record_hit(es, a, bb);
                            }
                        }
                    }
                }
            }
        }
    }
    if (lane == 0 && evaluated) atomicAdd(pairs_eval, (unsigned long long)evaluated);
}

template <int LP, int K, bool HASN>
static int bs_launch_one(cudaStream_t stream, int num_sms, const TileItem *items, u32 n_items, const uint2 *planes,
                         const u32 *nplane, const u32 *bsum, int L, int cull, EdgeSink es, u32 *counter, unsigned long long *pairs_eval) {
    constexpr int XS = HASN ? 8 : 4;
    size_t smem = (size_t)BS_G4 * LP * XS * 16;
    auto kern = hamming_tiles_bs<LP, K, HASN>;
    if (smem + 1024 > 48 * 1024) {      // + the kernel's static shared memory
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    }
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, BS_THREADS, smem) != cudaSuccess || occ < 1) occ = 1;
    u32 grid = (u32)std::min<u64>((u64)n_items, (u64)num_sms * occ);
    if (cudaMemsetAsync(counter, 0, sizeof(u32), stream) != cudaSuccess) return -1;
    kern<<<grid, BS_THREADS, smem, stream>>>(items, n_items, planes, nplane, bsum, L, cull, es, counter, pairs_eval);
    return cudaPeekAtLastError() == cudaSuccess ? 0 : -1;
}

template <int K, bool HASN>
static int bs_launch_k(cudaStream_t stream, int num_sms, const TileItem *items, u32 n_items, const uint2 *planes,
                       const u32 *nplane, const u32 *bsum, int L, int cull, EdgeSink es, u32 *counter, unsigned long long *pairs_eval) {
    if (L <= 8)  return bs_launch_one<8, K, HASN>(stream, num_sms, items, n_items, planes, nplane, bsum, L, cull, es, counter, pairs_eval);
    if (L <= 12) return bs_launch_one<12, K, HASN>(stream, num_sms, items, n_items, planes, nplane, bsum, L, cull, es, counter, pairs_eval);
    if (L <= 16) return bs_launch_one<16, K, HASN>(stream, num_sms, items, n_items, planes, nplane, bsum, L, cull, es, counter, pairs_eval);
    if (L <= 24) return bs_launch_one<24, K, HASN>(stream, num_sms, items, n_items, planes, nplane, bsum, L, cull, es, counter, pairs_eval);
    if (HASN) return 1;   // N-containing batches are limited to 21 nt upstream of here
    return bs_launch_one<32, K, false>(stream, num_sms, items, n_items, planes, nplane, bsum, L, cull, es, counter, pairs_eval);
}

// returns 0 = launched, 1 = configuration not covered (caller uses the direct kernel), -1 = CUDA error
static int launch_neighbours_bitsliced(cudaStream_t stream, int num_sms, const TileItem *items, u32 n_items,
                                       const uint2 *planes, const u32 *nplane, const u32 *bsum, int L, int k, bool has_n, int cull, EdgeSink es,
                                       u32 *counter, unsigned long long *pairs_eval) {
    if (k < 1 || k > 3) return 1;
    if (!has_n) {
        if (k == 1) return bs_launch_k<1, false>(stream, num_sms, items, n_items, planes, nplane, bsum, L, cull, es, counter, pairs_eval);
        if (k == 2) return bs_launch_k<2, false>(stream, num_sms, items, n_items, planes, nplane, bsum, L, cull, es, counter, pairs_eval);
        return bs_launch_k<3, false>(stream, num_sms, items, n_items, planes, nplane, bsum, L, cull, es, counter, pairs_eval);
    }
    if (k == 1) return bs_launch_k<1, true>(stream, num_sms, items, n_items, planes, nplane, bsum, L, cull, es, counter, pairs_eval);
    if (k == 2) return bs_launch_k<2, true>(stream, num_sms, items, n_items, planes, nplane, bsum, L, cull, es, counter, pairs_eval);
    return bs_launch_k<3, true>(stream, num_sms, items, n_items, planes, nplane, bsum, L, cull, es, counter, pairs_eval);
}
