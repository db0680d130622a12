// common.cuh — shared device/host helpers for libumigpu (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>

typedef uint8_t  u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int32_t  i32;
typedef int64_t  i64;

#define NUM_SMS_B200 148

static inline u64 ceil_div_u64(u64 a, u64 b) { return (a + b - 1) / b; }

// Growable device buffer: capacity only ever grows, so a context that processes batch after
// batch stops calling cudaMalloc after the first one.
struct DevBuf {
    void  *p = nullptr;
    size_t cap = 0;
    bool   borrowed = false;   // p belongs to the caller (zero-copy device push); cap stays 0 so any growth copies out of it
    void borrow(const void *ptr) { release(); p = const_cast<void *>(ptr); borrowed = true; }
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        size_t want = bytes + bytes / 8 + 256;
        if (borrowed) { p = nullptr; borrowed = false; }
        if (p) { cudaError_t e = cudaFree(p); p = nullptr; cap = 0; if (e != cudaSuccess) return e; }
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { p = nullptr; cap = 0; return e; }
        cap = want;
        return cudaSuccess;
    }
    // grow but keep the first `keep_bytes` bytes
    cudaError_t reserve_keep(size_t bytes, size_t keep_bytes, cudaStream_t s) {
        if (bytes <= cap) return cudaSuccess;
        size_t want = bytes + bytes / 2 + 256;
        void *q = nullptr;
        cudaError_t e = cudaMalloc(&q, want);
        if (e != cudaSuccess) return e;
        if (p && keep_bytes) {
            e = cudaMemcpyAsync(q, p, keep_bytes, cudaMemcpyDeviceToDevice, s);
            if (e != cudaSuccess) { cudaFree(q); return e; }
            e = cudaStreamSynchronize(s);
            if (e != cudaSuccess) { cudaFree(q); return e; }
        }
        if (p && !borrowed) cudaFree(p);
        p = q; cap = want; borrowed = false;
        return cudaSuccess;
    }
    void release() { if (p && !borrowed) cudaFree(p); p = nullptr; cap = 0; borrowed = false; }
    template <class T> T *as() const { return (T *)p; }
};

__device__ __forceinline__ u32 lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ u32 lanemask_lt() {
    u32 m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m;
}

// src/algo/directional.rs:38  threshold = (percentage * (freq + 1) as f32) as i32 — f32 multiply
// (no fma contraction), round-toward-zero saturating cast like Rust's `as i32`.
__device__ __forceinline__ i32 dir_threshold(float p, i32 freq) {
    float t = __fmul_rn(p, __int2float_rn(freq + 1));
    if (t != t) return 0;
    if (t >= 2147483648.0f) return 0x7fffffff;
    if (t <= -2147483648.0f) return (i32)0x80000000;
    return __float2int_rz(t);
}
