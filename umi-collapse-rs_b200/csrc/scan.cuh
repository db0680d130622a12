// scan.cuh — three-phase device-wide exclusive scan with fused producer / consumer functors.
//   phase 1  scan_tile_sums   : tile_sums[t] = sum_{i in tile t} f(i)
//   phase 2  scan_spine       : exclusive scan of tile_sums in place (one CTA), total -> *total
//   phase 3  scan_apply       : g(i, f(i), exclusive prefix of f at i)
// f is recomputed in phase 3 instead of being stored: the producers here are one or two loads
// and a compare, so recomputing costs less HBM traffic than a flag array would.
#pragma once
#include "common.cuh"
#include <type_traits>
#include <utility>

#define SCAN_THREADS 256
#define SCAN_ITEMS   8
#define SCAN_TILE    (SCAN_THREADS * SCAN_ITEMS)

template <class T>
__device__ __forceinline__ T warp_inclusive_scan(T v) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane_id() >= (u32)o) v += t;
    }
    return v;
}

// exclusive scan of one value per thread across the CTA; returns exclusive prefix, *total = CTA sum
template <class T, int THREADS>
__device__ __forceinline__ T block_exclusive_scan(T v, T *smem /* THREADS/32 + 1 */, T *total) {
    T inc = warp_inclusive_scan(v);
    u32 w = threadIdx.x >> 5;
    if (lane_id() == 31) smem[w] = inc;
    __syncthreads();
    if (w == 0) {
        T s = lane_id() < THREADS / 32 ? smem[lane_id()] : (T)0;
        T si = warp_inclusive_scan(s);
        if (lane_id() < THREADS / 32) smem[lane_id()] = si - s;
        if (lane_id() == THREADS / 32 - 1) smem[THREADS / 32] = si;
    }
    __syncthreads();
    T res = smem[w] + inc - v;
    *total = smem[THREADS / 32];
    __syncthreads();
    return res;
}

template <class T, class F>
__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_sums(F f, u64 n, T *tile_sums) {
    __shared__ T sm[SCAN_THREADS / 32 + 1];
    // only the tile's sum is needed, so the tile is read striped (coalesced) rather than blocked
    u64 base = (u64)blockIdx.x * SCAN_TILE + (u64)threadIdx.x;
    T s = 0;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; j++) {
        u64 i = base + (u64)j * SCAN_THREADS;
        if (i < n) s += f(i);
    }
    T total;
    block_exclusive_scan<T, SCAN_THREADS>(s, sm, &total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

template <class T>
__global__ void __launch_bounds__(1024) scan_spine(T *tile_sums, u64 ntiles, T *total_out) {
    __shared__ T sm[1024 / 32 + 1];
    T carry = 0;
    for (u64 base = 0; base < ntiles; base += 1024) {
        u64 i = base + threadIdx.x;
        T v = i < ntiles ? tile_sums[i] : (T)0;
        T total;
        T ex = block_exclusive_scan<T, 1024>(v, sm, &total);
        if (i < ntiles) tile_sums[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry;
}

// consumers may keep per-thread state across their SCAN_ITEMS consecutive elements and flush it in finish()
template <class G, class = void> struct scan_has_finish : std::false_type {};
template <class G> struct scan_has_finish<G, std::void_t<decltype(std::declval<G &>().finish())>> : std::true_type {};

// consumers with dependent gathers (index -> payload) can issue them for all of a thread's elements up front:
// prefetch(i, j) is called for every valid element before the scan, operator() then receives the slot j as well
template <class G, class = void> struct scan_has_prefetch : std::false_type {};
template <class G> struct scan_has_prefetch<G, std::void_t<decltype(std::declval<G &>().prefetch(u64(0), 0))>> : std::true_type {};

template <class T, class F, class G>
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply(F f, G g_in, u64 n, const T *tile_sums) {
    G g = g_in;
    __shared__ T sm[SCAN_THREADS / 32 + 1];
    u64 base = (u64)blockIdx.x * SCAN_TILE + (u64)threadIdx.x * SCAN_ITEMS;
    if constexpr (scan_has_prefetch<G>::value) {
#pragma unroll
        for (int j = 0; j < SCAN_ITEMS; j++) if (base + j < n) g.prefetch(base + j, j);
    }
    T v[SCAN_ITEMS];
    T s = 0;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; j++) {
        u64 i = base + j;
        v[j] = i < n ? f(i) : (T)0;
        s += v[j];
    }
    T total;
    T ex = block_exclusive_scan<T, SCAN_THREADS>(s, sm, &total) + tile_sums[blockIdx.x];
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; j++) {
        u64 i = base + j;
        if (i < n) {
            if constexpr (scan_has_prefetch<G>::value) g(i, v[j], ex, j); else g(i, v[j], ex);
        }
        ex += v[j];
    }
    if constexpr (scan_has_finish<G>::value) g.finish();
}
