// pack.cuh — K1 umi_pack and K1b build_keys.
//   K1  replaces to_bitset (src/utils/mod.rs:63-83, table src/utils/read.rs:22-31): ASCII UMI ->
//       2 bits/base (A0 C1 G2 T3, base 0 most significant, so integer order = string order) plus a
//       1 bit/base N mask.  The reference's 3-bit equidistant code exists only to make
//       popcount(xor)/2 a Hamming distance; the bit-plane form used by K5 gives the same distance
//       (see hamming.cuh).  Also reduces min/max of tid and unclipped position for the key layout.
//   K1b replaces the Alignment key (src/deduplicate_sam.rs:485-514, built :131-146): packs
//       (tid, unclipped pos, strand, UMI code) into one 64/128-bit sort key using only as many
//       bits as the batch needs.
#pragma once
#include "common.cuh"

struct DevScalars {
    i32 tid_min, tid_max;
    i64 pos_min, pos_max;
    u32 any_n, bad_base;
    u32 n_unique, n_buckets;
    u32 max_umis, changed;
    u32 n_kept, n_items;
    u64 pairs, pairs_eval, edge_count;
    u64 scratch;      // work counter of the neighbour kernel
    u64 scratch2;     // pairs in scheduled tile items (direct kernel accounting)
    u32 n_cand, n_tiles;
    u32 sort_ticket, sort_err;
    u32 bam_err, bam_valid;
    u32 n_blocks, pad0;
    u64 n_block_pairs;
    u32 n_big, m_big;
    i64 tlen_min, tlen_max;          // paired mode: template length is part of the bucket key (deduplicate_sam.rs:545-552)
    u64 n_unpaired, n_chimeric, n_mates_skipped, n_bam_unmapped;   // running totals over the BAM pushes of this batch
    u32 n_lowered, unsorted;         // label lowerings of the last sweep; bucket part of the keys not non-decreasing in input order
    u32 frontier_cnt[2];             // frontier clustering: sizes of the two ping-pong frontiers
    u32 hot_bucket, hot_u0, hot_cnt, hot_pad;   // sharded run: the bucket whose neighbour search is split across devices
    i64 key_lo, key_hi;              // sharded run: smallest / largest (tid << 32 | biased pos) of the slice (range check of the cuts)
    u32 seg_n_big, seg_n_tiles, seg_unsorted, seg_pad;   // segmented sort plan (SegPlanOut, seg_sort.cuh)
    u32 n_big_all, m_big_all;        // buckets with more than MI_BIG unique UMIs and their unique UMIs (bucket_stats_kernel)
};

struct KeyLayout {
    int umi_len, umi_bits, pos_bits, tid_bits, tlen_bits, bucket_bits, total_bits, nw, has_n;
    i64 pos_min, tlen_min;
    i32 tid_min;
    // linear coordinate layout (many contigs): position field = lin_off[tid - tid_min] + (pos - lin_pmin[tid - tid_min]),
    // i.e. the contigs' OCCUPIED position ranges laid end to end; tid_bits = 0.  nullptr = the plain [tid | pos] layout.
    const u64 *lin_off; const i64 *lin_pmin;
};

// Four bases per step (SWAR).  utils/read.rs:22-31 alphabet; anything else panics in the reference (utils/mod.rs:78).
// Bits 1-2 of the ASCII code separate A,C,T,G (0,1,3,2 -> Gray step -> A0 C1 G2 T3); a byte is valid iff it equals the
// letter its own bits name (one PRMT rebuilds the four letters); one multiply gathers the four 2-bit codes.  Only words
// holding an N or a bad byte take the per-byte path.  `word` holds bases b .. b+nb-1 of an L-base UMI, first base in the
// low byte.
__device__ __forceinline__ void pack_word(u32 word, int nb, int L, int b, u64 &code, u32 &nm, u32 &flags) {
    if (nb < 4) { const u32 m = (1u << (8 * nb)) - 1u; word = (word & m) | (0x41414141u & ~m); }
    const u32 x = (word >> 1) & 0x03030303u;
    u32 v = x ^ ((x >> 1) & 0x01010101u);
    const u32 t = v | (v >> 4);
    const u32 rec = __byte_perm(0x54474341u, 0u, __byte_perm(t, 0u, 0x4420));
    const u32 diff = word ^ rec;
    if (diff) {
        for (int j = 0; j < nb; j++) {
            if ((diff >> (8 * j)) & 0xffu) {
                v &= ~(0xffu << (8 * j));
                if (((word >> (8 * j)) & 0xffu) == 'N') nm |= 1u << (L - 1 - (b + j)); else flags |= 2u;
            }
        }
    }
    const u32 p8 = (v * 0x40100401u) >> 24;          // v0<<6 | v1<<4 | v2<<2 | v3
    code = (code << (2 * nb)) | (u64)(p8 >> (2 * (4 - nb)));
}

// Four consecutive reads of a UMI length 4*LW per thread, everything as 16-byte vectors: LW loads of ASCII in, two
// stores of codes and one of N masks out.
template <int LW>
__device__ __forceinline__ void pack_quad(const u8 *__restrict__ ascii, u64 g, u64 *__restrict__ umi2, u32 *__restrict__ nmask, u32 &flags) {
    constexpr int L = 4 * LW;
    const uint4 *src = reinterpret_cast<const uint4 *>(ascii) + g * LW;
    u32 w[4 * LW];
#pragma unroll
    for (int q = 0; q < LW; q++) { const uint4 t = __ldg(src + q); w[4 * q] = t.x; w[4 * q + 1] = t.y; w[4 * q + 2] = t.z; w[4 * q + 3] = t.w; }
    u64 code[4]; u32 nm[4];
#pragma unroll
    for (int r = 0; r < 4; r++) {
        code[r] = 0; nm[r] = 0;
#pragma unroll
        for (int k = 0; k < LW; k++) pack_word(w[r * LW + k], 4, L, 4 * k, code[r], nm[r], flags);
    }
    ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(umi2 + 4 * g);
    dst[0] = make_ulonglong2(code[0], code[1]); dst[1] = make_ulonglong2(code[2], code[3]);
    *reinterpret_cast<uint4 *>(nmask + 4 * g) = make_uint4(nm[0], nm[1], nm[2], nm[3]);
    if (nm[0] | nm[1] | nm[2] | nm[3]) flags |= 1u;
}

#define PACK_THREADS 256

__global__ void __launch_bounds__(PACK_THREADS) umi_pack_kernel(
    const u8 *__restrict__ ascii, u64 n, int L, const i32 *__restrict__ tid, const i64 *__restrict__ pos,
    const i64 *__restrict__ tlen, u64 *__restrict__ umi2, u32 *__restrict__ nmask, DevScalars *sc) {
    __shared__ __align__(16) u8 sbuf[PACK_THREADS * 32 + 16];
    __shared__ i64 s_lmin[PACK_THREADS / 32], s_lmax[PACK_THREADS / 32];
    i64 lmin = 0x7fffffffffffffffLL, lmax = (i64)0x8000000000000000LL;
    __shared__ i32 s_tmin[PACK_THREADS / 32], s_tmax[PACK_THREADS / 32];
    __shared__ i64 s_pmin[PACK_THREADS / 32], s_pmax[PACK_THREADS / 32];
    __shared__ u32 s_flags[PACK_THREADS / 32];
    i32 tmin = 0x7fffffff, tmax = (i32)0x80000000;
    i64 pmin = 0x7fffffffffffffffLL, pmax = (i64)0x8000000000000000LL;
    u32 flags = 0;     // bit 0 = any N, bit 1 = bad base
    // persistent CTAs: a grid-stride loop over 256-read tiles, ONE range reduction per CTA at the end
    // L a multiple of 4 on a 4-byte aligned array: every read is whole words, loaded straight from global (a warp's loads
    // cover one contiguous span; L1 serves the repeats) — no staging, no barriers
    const bool direct = (L & 3) == 0 && (((unsigned long long)ascii) & 3ull) == 0;
    // 8-, 12- and 16-nt UMIs on 16-byte aligned arrays: four reads per thread, vector loads and stores throughout
    u64 n4 = 0;
    if ((L == 8 || L == 12 || L == 16) &&
        ((((unsigned long long)ascii) | ((unsigned long long)tid) | ((unsigned long long)pos) | ((unsigned long long)umi2) |
          ((unsigned long long)nmask) | ((unsigned long long)tlen)) & 15ull) == 0) {
        n4 = n & ~3ull;
        for (u64 g = (u64)blockIdx.x * PACK_THREADS + threadIdx.x; g < (n4 >> 2); g += (u64)gridDim.x * PACK_THREADS) {
            if (L == 12) pack_quad<3>(ascii, g, umi2, nmask, flags); else if (L == 8) pack_quad<2>(ascii, g, umi2, nmask, flags);
            else pack_quad<4>(ascii, g, umi2, nmask, flags);
            const int4 t4 = __ldg(reinterpret_cast<const int4 *>(tid) + g);
            tmin = min(tmin, min(min(t4.x, t4.y), min(t4.z, t4.w))); tmax = max(tmax, max(max(t4.x, t4.y), max(t4.z, t4.w)));
            const longlong2 p0 = __ldg(reinterpret_cast<const longlong2 *>(pos) + 2 * g), p1 = __ldg(reinterpret_cast<const longlong2 *>(pos) + 2 * g + 1);
            pmin = min(pmin, min(min((i64)p0.x, (i64)p0.y), min((i64)p1.x, (i64)p1.y))); pmax = max(pmax, max(max((i64)p0.x, (i64)p0.y), max((i64)p1.x, (i64)p1.y)));
            if (tlen) {
                const longlong2 l0 = __ldg(reinterpret_cast<const longlong2 *>(tlen) + 2 * g), l1 = __ldg(reinterpret_cast<const longlong2 *>(tlen) + 2 * g + 1);
                lmin = min(lmin, min(min((i64)l0.x, (i64)l0.y), min((i64)l1.x, (i64)l1.y))); lmax = max(lmax, max(max((i64)l0.x, (i64)l0.y), max((i64)l1.x, (i64)l1.y)));
            }
        }
    }
    for (u64 base = n4 + (u64)blockIdx.x * PACK_THREADS; base < n; base += (u64)gridDim.x * PACK_THREADS) {
        const u32 cnt = (u32)min((u64)PACK_THREADS, n - base);
        const u32 bytes = cnt * (u32)L;
        const u8 *src = ascii + base * (u64)L;
        if (!direct) __syncthreads();         // the previous tile's bytes have been consumed
        // stage the tile's contiguous ASCII span with coalesced (vector) loads
        if (direct) {
        } else if ((((unsigned long long)src) & 15ull) == 0) {
            const u32 nv = bytes >> 4;
            for (u32 v = threadIdx.x; v < nv; v += PACK_THREADS) reinterpret_cast<uint4 *>(sbuf)[v] = reinterpret_cast<const uint4 *>(src)[v];
            for (u32 b = (nv << 4) + threadIdx.x; b < bytes; b += PACK_THREADS) sbuf[b] = src[b];
        } else {
            for (u32 b = threadIdx.x; b < bytes; b += PACK_THREADS) sbuf[b] = src[b];
        }
        if (!direct) __syncthreads();
        if (threadIdx.x < cnt) {
            const u64 i = base + threadIdx.x;
            const u32 o = threadIdx.x * (u32)L, sh = 8u * (o & 3u);
            const u32 *sw = reinterpret_cast<const u32 *>(sbuf) + (o >> 2);
            const u32 *gw = reinterpret_cast<const u32 *>(src) + (o >> 2);
            u64 code = 0; u32 nm = 0;
            u32 lo = direct ? 0u : sw[0];
            for (int b = 0; b < L; b += 4) {
                u32 word;
                if (direct) word = __ldg(gw + (b >> 2));
                else { const u32 hi = sw[(b >> 2) + 1]; word = __funnelshift_r(lo, hi, sh); lo = hi; }
                pack_word(word, min(4, L - b), L, b, code, nm, flags);
            }
            umi2[i] = code; nmask[i] = nm; if (nm) flags |= 1u;
            const i32 t = tid[i]; const i64 p = pos[i];
            tmin = min(tmin, t); tmax = max(tmax, t); pmin = min(pmin, p); pmax = max(pmax, p);
            if (tlen) { const i64 l = tlen[i]; lmin = min(lmin, l); lmax = max(lmax, l); }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        tmin = min(tmin, __shfl_xor_sync(0xffffffffu, tmin, o));
        tmax = max(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
        pmin = min(pmin, __shfl_xor_sync(0xffffffffu, pmin, o));
        pmax = max(pmax, __shfl_xor_sync(0xffffffffu, pmax, o));
        lmin = min(lmin, __shfl_xor_sync(0xffffffffu, lmin, o));
        lmax = max(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
        flags |= __shfl_xor_sync(0xffffffffu, flags, o);
    }
    const u32 w = threadIdx.x >> 5;
    if (lane_id() == 0) { s_tmin[w] = tmin; s_tmax[w] = tmax; s_pmin[w] = pmin; s_pmax[w] = pmax; s_lmin[w] = lmin; s_lmax[w] = lmax; s_flags[w] = flags; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < PACK_THREADS / 32; k++) {
            tmin = min(tmin, s_tmin[k]); tmax = max(tmax, s_tmax[k]); pmin = min(pmin, s_pmin[k]); pmax = max(pmax, s_pmax[k]);
            lmin = min(lmin, s_lmin[k]); lmax = max(lmax, s_lmax[k]);
            flags |= s_flags[k];
        }
        if (tlen && lmin <= lmax) { atomicMin((long long *)&sc->tlen_min, (long long)lmin); atomicMax((long long *)&sc->tlen_max, (long long)lmax); }
        if (tmin <= tmax) {
            atomicMin(&sc->tid_min, tmin); atomicMax(&sc->tid_max, tmax);
            atomicMin((long long *)&sc->pos_min, (long long)pmin); atomicMax((long long *)&sc->pos_max, (long long)pmax);
        }
        if (flags & 1u) atomicOr(&sc->any_n, 1u);
        if (flags & 2u) atomicOr(&sc->bad_base, 1u);
    }
}

// UMI code used inside the sort key: 2 bits/base when the batch has no N, else 3 bits/base with
// N = 4, so that ascending code = ascending string with A < C < G < T < N (the canonical tie-break).
__device__ __forceinline__ u64 umi_sort_code(u64 umi2, u32 nm, int L, int has_n) {
    if (!has_n) return umi2;
    u64 code = 0;
    for (int b = 0; b < L; b++) {
        int sh = L - 1 - b;
        u64 c = (nm >> sh) & 1 ? 4 : (umi2 >> (2 * sh)) & 3;
        code = (code << 3) | c;
    }
    return code;
}

// the same code for UMIs whose 3-bit form exceeds one word (N present, 22..32 nt): low 64 bits returned, the rest in hi
__device__ __forceinline__ u64 umi_sort_code_wide(u64 umi2, u32 nm, int L, u64 &hi) {
    u64 lo = 0; hi = 0;
    for (int b = 0; b < L; b++) {
        int sh = L - 1 - b;
        u64 c = (nm >> sh) & 1 ? 4 : (umi2 >> (2 * sh)) & 3;
        hi = (hi << 3) | (lo >> 61);
        lo = (lo << 3) | c;
    }
    return lo;
}

template <int NW, bool LIN>
__global__ void __launch_bounds__(256) build_keys_kernel(
    u64 n, const i32 *__restrict__ tid, const i64 *__restrict__ pos, const u8 *__restrict__ rev,
    const i64 *__restrict__ tlen, const u64 *__restrict__ umi2, const u32 *__restrict__ nmask, KeyLayout lay, u64 *__restrict__ k0, u64 *__restrict__ k1) {
    u64 i = (u64)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    u64 bucket;
    if (LIN) { const u32 t = (u32)(tid[i] - lay.tid_min); bucket = ((__ldg(lay.lin_off + t) + (u64)(pos[i] - __ldg(lay.lin_pmin + t))) << 1) | (rev[i] ? 1u : 0u); }
    else bucket = ((u64)(u32)(tid[i] - lay.tid_min) << (lay.pos_bits + 1)) | ((u64)(pos[i] - lay.pos_min) << 1) | (rev[i] ? 1u : 0u);
    if (lay.tlen_bits) bucket = (bucket << lay.tlen_bits) | (u64)(tlen[i] - lay.tlen_min);    // PairedAlignment: + tlen
    int ub = lay.umi_bits;
    if (NW == 2 && ub > 64) {                 // N present and 3 bits per base exceed one word: [bucket | code] = k1:k0 with the code's top in k1
        u64 chi;
        k0[i] = umi_sort_code_wide(umi2[i], nmask[i], lay.umi_len, chi);
        k1[i] = (bucket << (ub - 64)) | chi;
        return;
    }
    u64 code = umi_sort_code(umi2[i], nmask[i], lay.umi_len, lay.has_n);
    u64 lo = (ub < 64 ? bucket << ub : 0) | code;
    k0[i] = lo;
    if (NW == 2) k1[i] = ub == 64 ? bucket : (ub == 0 ? 0 : bucket >> (64 - ub));
}

// Compact host format (umigpu_push_reads_packed): 32-bit positions, UMIs already 2 bit/base (u32 when umi_len <= 16, else
// u64), 8-bit scores.  Widened into the SoA the path works on; N positions get code 0 like in umi_pack_kernel; a code
// with bits above 2*umi_len, or an N mask with bits above umi_len, is an error (bad_base).
__global__ void __launch_bounds__(256) unpack_compact_kernel(u64 n, int L, const i32 *__restrict__ pos32, const void *__restrict__ umi_c,
                                                             const u32 *__restrict__ nmask_c, const u8 *__restrict__ score8,
                                                             i64 *__restrict__ pos, u64 *__restrict__ umi2, u32 *__restrict__ nmask,
                                                             i32 *__restrict__ score, DevScalars *sc) {
    const u64 i = (u64)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    pos[i] = (i64)pos32[i];
    u64 code = L <= 16 ? (u64)reinterpret_cast<const u32 *>(umi_c)[i] : reinterpret_cast<const u64 *>(umi_c)[i];
    const u32 nm = nmask_c ? nmask_c[i] : 0u;
    bool bad = (L < 32 && (code >> (2 * L)) != 0) || (L < 32 && (nm >> L) != 0);
    if (nm) {
        u64 spread = 0;
        for (int b = 0; b < L; b++) if ((nm >> b) & 1u) spread |= 3ull << (2 * b);
        code &= ~spread;
    }
    umi2[i] = code; nmask[i] = nm;
    if (score) score[i] = (i32)score8[i];
    if (bad) sc->bad_base = 1;
}

// Occupied position range of every contig (order-preserving unsigned bias so that the tables can be memset), for the
// linear coordinate layout.  Coordinate-sorted input: a warp usually holds one contig -> one pair of atomics per warp.
#define POS_BIAS 0x8000000000000000ull
__global__ void __launch_bounds__(256) tid_range_kernel(u64 n, const i32 *__restrict__ tid, const i64 *__restrict__ pos, i32 tid_min,
                                                        unsigned long long *vmin, unsigned long long *vmax) {
    const u64 stride = (u64)gridDim.x * 256, rounds = (n + stride - 1) / stride;
    for (u64 it = 0; it < rounds; it++) {
        const u64 i = it * stride + (u64)blockIdx.x * 256 + threadIdx.x;
        const bool a = i < n;
        const u32 t = a ? (u32)(tid[i] - tid_min) : 0xffffffffu;
        unsigned long long lo = a ? ((unsigned long long)pos[i] ^ POS_BIAS) : ~0ull, hi = a ? lo : 0ull;
        const u32 t0 = __shfl_sync(0xffffffffu, t, 0);
        if (__all_sync(0xffffffffu, t == t0 || !a)) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const unsigned long long l2 = __shfl_xor_sync(0xffffffffu, lo, o), h2 = __shfl_xor_sync(0xffffffffu, hi, o);
                lo = l2 < lo ? l2 : lo; hi = h2 > hi ? h2 : hi;
            }
            if (lane_id() == 0 && t0 != 0xffffffffu) { atomicMin(&vmin[t0], lo); atomicMax(&vmax[t0], hi); }
        } else if (a) { atomicMin(&vmin[t], lo); atomicMax(&vmax[t], hi); }
    }
}

// bit planes of a sort code: plane0 bit b = low bit of base b's code, plane1 = high bit, planeN = N flag.
// (base 0 lands on bit L-1; any fixed permutation of positions leaves the Hamming distance unchanged)
__device__ __forceinline__ void code_to_planes(u64 code, int L, int has_n, u32 &p0, u32 &p1, u32 &pn) {
    p0 = p1 = pn = 0;
    if (!has_n) {
        for (int b = 0; b < L; b++) { u32 c = (u32)(code >> (2 * b)) & 3; p0 |= (c & 1) << b; p1 |= (c >> 1) << b; }
    } else {
        for (int b = 0; b < L; b++) { u32 c = (u32)(code >> (3 * b)) & 7; p0 |= (c & 1) << b; p1 |= ((c >> 1) & 1) << b; pn |= (c >> 2) << b; }
    }
}
