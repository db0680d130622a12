// hamming.cuh — K5 hamming_neighbours: all-pairs Hamming <= k over the unique UMIs of a bucket,
// tile by tile, emitting the directed edges that pass the reference's count rule.
//
// Replaces Naive::remove_near (src/data/naive.rs:26-40) + umi_dist (src/utils/mod.rs:24-26,
// src/utils/bitset.rs:77-91).  Naive scans every remaining UMI once per removed UMI; here every
// unordered pair of a bucket is evaluated exactly once and the result is kept as an edge list, so
// the clustering (K6) never needs a distance again.
//
// Distance: a UMI is three bit planes (one bit per base): p0/p1 = low/high bit of the base code,
// pn = "is N".  mismatch mask m = (a.p0^b.p0) | (a.p1^b.p1) | (a.pn^b.pn); dist = popc(m).  This is
// plain Hamming distance over {A,C,G,T,N} (N==N matches, N!=base mismatches), which is exactly what
// the reference's popcount formula computes (bitset.rs:77-91; checked in tests/test_oracle.py).
// Any umi_len <= 32 costs the same: 2 LOP3 per pair for m, +1 with N.
//
// Edge rule (directional.rs:38, naive.rs:31): u -> v  iff dist(u,v) <= k and freq[v] <= thr[u],
// thr[u] = trunc(p * (freq[u] + 1)) precomputed per UMI (INT_MAX for cc / upstream adjacency).
#pragma once
#include "common.cuh"
#include "group.cuh"

#define HT_THREADS 256
#define HT_RPT     (HT_ROWS / HT_THREADS)   // rows per thread = 8

struct EdgeSink {
    uint2 *edges;
    unsigned long long *count;
    u64 cap;
    const i32 *freq, *thr;
};

// cold path: a < b are unique ids with dist <= k
__device__ __noinline__ void record_hit(const EdgeSink &es, u32 a, u32 b) {
    i32 fa = es.freq[a], fb = es.freq[b];
    bool ab = fb <= es.thr[a], ba = fa <= es.thr[b];
    u32 cnt = (ab ? 1 : 0) + (ba ? 1 : 0);
    if (!cnt) return;
    u64 at = atomicAdd(es.count, (unsigned long long)cnt);
    if (ab) { if (at < es.cap) es.edges[at] = make_uint2(a, b); at++; }
    if (ba) { if (at < es.cap) es.edges[at] = make_uint2(b, a); }
}

template <int K>
__device__ __forceinline__ bool within_k(u32 m) {
    if (K == 1) return (m & (m - 1)) == 0;
    return __popc(m) <= K;
}

// Direct tile kernel: each thread keeps HT_RPT row UMIs in registers, the column tile sits in
// shared memory and is read with broadcast LDS.128 (two columns per load).
template <int K, bool HASN>
__global__ void __launch_bounds__(HT_THREADS) hamming_tiles_direct(
    const TileItem *__restrict__ items, u32 n_items, const uint2 *__restrict__ planes, const u32 *__restrict__ nplane,
    EdgeSink es, int kdyn) {
    __shared__ __align__(16) uint2 scol[HT_COLS];
    __shared__ u32 sncol[HASN ? HT_COLS : 1];
    for (u32 w = blockIdx.x; w < n_items; w += gridDim.x) {
        TileItem it = items[w];
        const u32 col_cnt = item_col_cnt(it), row_cnt = item_row_cnt(it);
        const bool diag = item_diag(it);
        const u32 cols_pad = (col_cnt + 3) & ~3u;
        __syncthreads();
        for (u32 c = threadIdx.x; c < cols_pad; c += HT_THREADS) {
            bool v = c < col_cnt;
            scol[c] = v ? planes[it.col_start + c] : make_uint2(0xffffffffu, 0xffffffffu);
            if (HASN) sncol[c] = v ? nplane[it.col_start + c] : 0u;
        }
        u32 r0[HT_RPT], r1[HT_RPT], rn[HT_RPT];
#pragma unroll
        for (int r = 0; r < HT_RPT; r++) {
            u32 gi = threadIdx.x + r * HT_THREADS;
            bool v = gi < row_cnt;
            uint2 p = v ? planes[it.row_start + gi] : make_uint2(0xffffffffu, 0u);
            r0[r] = p.x; r1[r] = p.y;
            rn[r] = (HASN && v) ? nplane[it.row_start + gi] : 0u;
        }
        __syncthreads();
        for (u32 c = 0; c < cols_pad; c += 2) {
            uint4 cc = *reinterpret_cast<const uint4 *>(&scol[c]);
            u32 n0 = 0, n1 = 0;
            if (HASN) { n0 = sncol[c]; n1 = sncol[c + 1]; }
#pragma unroll
            for (int r = 0; r < HT_RPT; r++) {
                u32 m0 = (r0[r] ^ cc.x) | (r1[r] ^ cc.y);
                u32 m1 = (r0[r] ^ cc.z) | (r1[r] ^ cc.w);
                if (HASN) { m0 |= rn[r] ^ n0; m1 |= rn[r] ^ n1; }
                bool h0 = K > 0 ? within_k<K>(m0) : (__popc(m0) <= kdyn);
                bool h1 = K > 0 ? within_k<K>(m1) : (__popc(m1) <= kdyn);
                if (h0 | h1) {
                    u32 gi = threadIdx.x + r * HT_THREADS;
                    if (gi < row_cnt) {
                        u32 a = it.row_start + gi;
                        if (h0 && c < col_cnt) { u32 b = it.col_start + c; if (!diag || a < b) record_hit(es, a, b); }
                        if (h1 && c + 1 < col_cnt) { u32 b = it.col_start + c + 1; if (!diag || a < b) record_hit(es, a, b); }
                    }
                }
            }
        }
    }
}

// Buckets with 2..32 unique UMIs (the vast majority of alignment positions): one warp per bucket, one UMI
// per lane, every other UMI of the bucket arrives by shuffle.  No shared memory, no tile item.
__global__ void __launch_bounds__(256) small_buckets_kernel(u32 n_buckets, const u32 *__restrict__ bstart,
                                                            const uint2 *__restrict__ planes, const u32 *__restrict__ nplane,
                                                            int k, EdgeSink es, unsigned long long *pairs_eval, u32 skip) {
    const u32 b = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = lane_id();
    u64 np = 0;
    if (b < n_buckets && b != skip) {
        const u32 s = bstart[b], nb = bstart[b + 1] - s;
        if (nb >= 2 && nb <= SMALL_BUCKET) {
            const bool v = lane < nb;
            const uint2 p = v ? planes[s + lane] : make_uint2(0u, 0u);
            const u32 pn = (v && nplane) ? nplane[s + lane] : 0u;
            u32 hm = 0;                                   // bit j: UMI j (> lane) is within k of this lane's UMI
            for (u32 j = 1; j < nb; j++) {
                const u32 q0 = __shfl_sync(0xffffffffu, p.x, j), q1 = __shfl_sync(0xffffffffu, p.y, j);
                const u32 qn = __shfl_sync(0xffffffffu, pn, j);
                const u32 m = (p.x ^ q0) | (p.y ^ q1) | (pn ^ qn);
                if (lane < j && __popc(m) <= k) hm |= 1u << j;
            }
            if (__any_sync(0xffffffffu, hm != 0)) {
                // the count rule (directional.rs:38) for every hit, then ONE reservation per bucket: a global atomic per
                // hit serialises on the single edge counter
                const i32 fa = v ? es.freq[s + lane] : 0, ta = v ? es.thr[s + lane] : 0;
                u32 abm = 0, bam = 0;
                for (u32 j = 1; j < nb; j++) {
                    const i32 fj = __shfl_sync(0xffffffffu, fa, j), tj = __shfl_sync(0xffffffffu, ta, j);
                    if ((hm >> j) & 1u) { if (fj <= ta) abm |= 1u << j; if (fa <= tj) bam |= 1u << j; }
                }
                const u32 cnt = __popc(abm) + __popc(bam);
                u32 inc = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { u32 t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= (u32)o) inc += t; }
                const u32 total = __shfl_sync(0xffffffffu, inc, 31);
                if (total) {
                    unsigned long long base = 0;
                    if (lane == 0) base = atomicAdd(es.count, (unsigned long long)total);
                    base = __shfl_sync(0xffffffffu, base, 0) + inc - cnt;
                    while (abm) { const u32 j = __ffs(abm) - 1; abm &= abm - 1; if (base < es.cap) es.edges[base] = make_uint2(s + lane, s + j); base++; }
                    while (bam) { const u32 j = __ffs(bam) - 1; bam &= bam - 1; if (base < es.cap) es.edges[base] = make_uint2(s + j, s + lane); base++; }
                }
            }
            np = (u64)nb * (nb - 1) / 2;
        }
    }
    // one atomic per warp would serialise on a single address: aggregate per CTA through shared memory
    __shared__ unsigned long long s_np;
    if (threadIdx.x == 0) s_np = 0;
    __syncthreads();
    if (lane == 0 && np) atomicAdd(&s_np, (unsigned long long)np);
    __syncthreads();
    if (threadIdx.x == 0 && s_np) atomicAdd(pairs_eval, s_np);
}

// Naive::remove_near for ONE query (DataStruct-shaped API, naive.rs:26-40):
// out[i] = dist <= k && (dist == 0 || freq[i] <= max_freq)
__global__ void __launch_bounds__(256) remove_near_kernel(u32 n, const u64 *__restrict__ umi2, const u32 *__restrict__ nmask,
                                                          u64 q2, u32 qn, const i32 *__restrict__ freq, int k, i32 max_freq,
                                                          u8 *__restrict__ out) {
    u32 i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    u64 x = umi2[i] ^ q2;
    u64 m = (x | (x >> 1)) & 0x5555555555555555ull;
    // spread the N masks to the even bit of each base
    u32 nx = nmask[i] ^ qn;
    u64 ns = 0;
    for (int b = 0; b < 32; b++) ns |= (u64)((nx >> b) & 1) << (2 * b);
    int dist = __popcll(m | ns);
    out[i] = (dist <= k && (dist == 0 || freq[i] <= max_freq)) ? 1 : 0;
}
