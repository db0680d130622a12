// misc.cuh — score kernel (a4) and the integer-pipe microbenchmark used as roofline denominator.
#pragma once
#include "common.cuh"

// UcSAMRead::new, src/utils/read.rs:56-63: avg_qual = (sum_f32(qual) / seq_len as f32) as i32.
// The reference accumulates in f32 left to right; while the sum stays below 2^24 every partial sum
// is an exact integer, so an integer warp reduction gives the identical value.  Longer reads fall
// back to the same sequential f32 accumulation on one lane.  len == 0 -> NaN -> 0.
__global__ void __launch_bounds__(256) avg_qual_kernel(u64 n, const u8 *__restrict__ qual, const u64 *__restrict__ offsets,
                                                       i32 *__restrict__ out) {
    u64 read = ((u64)blockIdx.x * 256 + threadIdx.x) >> 5;
    if (read >= n) return;
    u64 b = offsets[read], e = offsets[read + 1], len = e - b;
    u32 lane = lane_id();
    float q;
    if (len <= 65536) {
        u32 s = 0;
        for (u64 i = b + lane; i < e; i += 32) s += qual[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        q = __fdiv_rn(__uint2float_rn(s), __ull2float_rn(len));
    } else {
        float acc = 0.0f;
        if (lane == 0) for (u64 i = b; i < e; i++) acc = __fadd_rn(acc, (float)qual[i]);
        acc = __shfl_sync(0xffffffffu, acc, 0);
        q = __fdiv_rn(acc, __ull2float_rn(len));
    }
    if (lane == 0) out[read] = (q != q) ? 0 : (q >= 2147483648.0f ? 0x7fffffff : __float2int_rz(q));
}

// Independent LOP3 / POPC streams: 8 accumulators per thread, no memory traffic.
// ops counted = iterations * 8 * threads (per instruction kind).
__global__ void __launch_bounds__(256) int_peak_lop3_kernel(u32 iters, u32 seed, u32 *out) {
    u32 a[8];
#pragma unroll
    for (int j = 0; j < 8; j++) a[j] = seed * (threadIdx.x + 1) + j * 0x9e3779b9u;
    u32 x = seed ^ 0x5bd1e995u ^ blockIdx.x;
    for (u32 i = 0; i < iters; i += 8) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int j = 0; j < 8; j++) asm volatile("lop3.b32 %0, %0, 0x7f3f1f0f, %1, 0x6a;" : "+r"(a[j]) : "r"(x));   // ONE LOP3 each: (a & imm) ^ x
        }
        x += 0x01000193u;
    }
    u32 r = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) r ^= a[j];
    if (r == 0x12345678u) out[0] = r;
}
__global__ void __launch_bounds__(256) int_peak_popc_kernel(u32 iters, u32 seed, u32 *out) {
    u32 a[8];
#pragma unroll
    for (int j = 0; j < 8; j++) a[j] = seed * (threadIdx.x + 1) + j * 0x9e3779b9u;
    for (u32 i = 0; i < iters; i += 4) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int j = 0; j < 8; j++) a[j] = __popc(a[j]) ^ 0x7f4a7c15u;   // POPC + LOP3
        }
    }
    u32 r = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) r ^= a[j];
    if (r == 0x12345678u) out[0] = r;
}
