// group.cuh — K3 umi_count_merge, bucket segmentation and the tile work list.
//   K3 replaces the per-read map update of src/deduplicate_sam.rs:160-176 with the Merge policies of
//   src/merge/mod.rs:18-51: for every (bucket, UMI) segment of the sorted reads, freq = number of
//   reads (or sum of weights) and representative = first read in input order attaining the maximum
//   score (`keep_existing = a.score >= b.score` only ever replaces on a strictly larger score).
#pragma once
#include "common.cuh"
#include "pack.cuh"
#include "scan.cuh"

struct SortedKeys {
    const u64 *k0, *k1;   // k1 == nullptr for one-word keys
    int umi_bits;
    __device__ __forceinline__ bool same_key(u64 i, u64 j) const {
        return k0[i] == k0[j] && (k1 == nullptr || k1[i] == k1[j]);
    }
    __device__ __forceinline__ bool same_bucket(u64 i, u64 j) const {
        if (umi_bits >= 64) return (k1[i] >> (umi_bits - 64)) == (k1[j] >> (umi_bits - 64));
        if ((k0[i] >> umi_bits) != (k0[j] >> umi_bits)) return false;
        return k1 == nullptr || k1[i] == k1[j];
    }
};

struct HeadFlag {
    SortedKeys sk;
    __device__ u32 operator()(u64 i) const { return (i == 0 || !sk.same_key(i, i - 1)) ? 1u : 0u; }
};

struct UniqueEmit {
    SortedKeys sk;
    u64 n;
    const u32 *idx;       // sorted position -> read index
    const i32 *score;     // per read, may be null (MERGE_ANY)
    const i32 *weight;    // per read, may be null
    int L, has_n;
    u32 *useg;            // [U+1] first sorted position of each unique
    uint2 *planes; u32 *nplane;
    u64 *ucode;           // the unique's sort code (2 or 3 bits per base): letter j = (code >> bpb*j) & mask
    u8 *bhead;            // unique starts a new bucket
    unsigned long long *rep;   // packed (score biased << 32 | ~read idx), max wins
    i32 *wsum;            // weighted freq (only with weights)
    u32 *read_uid;        // optional: read index -> unique id
    // a thread's SCAN_ITEMS elements are consecutive sorted reads: runs of one unique are folded in registers and
    // cost ONE atomic per run instead of one per read
    u32 pend_uid; unsigned long long pend_val; i32 pend_w;
    __device__ __forceinline__ void flush() {
        if (pend_uid != 0xffffffffu) {
            atomicMax(&rep[pend_uid], pend_val);
            if (weight) atomicAdd(&wsum[pend_uid], pend_w);
        }
    }
    __device__ void finish() { flush(); }
    // the idx -> score / weight gathers of a thread's SCAN_ITEMS reads are independent: issue them all before the scan
    u32 pf_r[SCAN_ITEMS]; i32 pf_s[SCAN_ITEMS], pf_w[SCAN_ITEMS];
    __device__ __forceinline__ void prefetch(u64 i, int j) {
        const u32 r = __ldg(idx + i);
        pf_r[j] = r;
        pf_s[j] = score ? __ldg(score + r) : 0;
        pf_w[j] = weight ? __ldg(weight + r) : 0;
    }
    __device__ __forceinline__ void operator()(u64 i, u32 flag, u32 ex, int j) {
        u32 uid = ex + flag - 1;
        u32 r = pf_r[j];
        if (flag) {
            useg[uid] = (u32)i;
            // the bit planes are derived from ucode in unique_finalize_kernel (one thread per unique): here the
            // conversion would run divergently, once per element slot of every warp that holds a head
            ucode[uid] = sk.umi_bits >= 64 ? sk.k0[i] : (sk.k0[i] & ((1ull << sk.umi_bits) - 1));
            bhead[uid] = (i == 0 || !sk.same_bucket(i, i - 1)) ? 1 : 0;
        }
        if (i == n - 1) useg[uid + 1] = (u32)n;
        u32 s = score ? (u32)pf_s[j] ^ 0x80000000u : 0u;
        const unsigned long long pk = ((unsigned long long)s << 32) | (u32)~r;
        const i32 wv = pf_w[j];
        if (uid == pend_uid) { pend_val = pk > pend_val ? pk : pend_val; pend_w += wv; }
        else { flush(); pend_uid = uid; pend_val = pk; pend_w = wv; }
        if (read_uid) read_uid[r] = uid;
    }
};

// Phase 3 of the merge scan for one-word keys: what scan_apply<HeadFlag, UniqueEmit> does, with the thread's eight keys
// and indices loaded as 16-byte vectors (6 loads instead of 24 strided ones) and the predecessor key taken from the
// neighbouring lane.  Same outputs, same atomics.
__global__ void __launch_bounds__(SCAN_THREADS) unique_apply1_kernel(UniqueEmit g_in, u64 n, const u32 *__restrict__ tile_sums) {
    static_assert(SCAN_ITEMS == 8, "vector loads below assume 8 elements per thread");
    UniqueEmit g = g_in;
    __shared__ u32 sm[SCAN_THREADS / 32 + 1];
    const u64 base = (u64)blockIdx.x * SCAN_TILE + (u64)threadIdx.x * SCAN_ITEMS;
    const u64 *__restrict__ k0 = g.sk.k0;
    u64 key[SCAN_ITEMS]; u32 ridx[SCAN_ITEMS];
    if (base + SCAN_ITEMS <= n) {
        const ulonglong2 *kv = reinterpret_cast<const ulonglong2 *>(k0 + base);
        const uint4 *iv = reinterpret_cast<const uint4 *>(g.idx + base);
#pragma unroll
        for (int q = 0; q < 4; q++) { const ulonglong2 t = __ldg(kv + q); key[2 * q] = t.x; key[2 * q + 1] = t.y; }
#pragma unroll
        for (int q = 0; q < 2; q++) { const uint4 t = __ldg(iv + q); ridx[4 * q] = t.x; ridx[4 * q + 1] = t.y; ridx[4 * q + 2] = t.z; ridx[4 * q + 3] = t.w; }
    } else {
#pragma unroll
        for (int j = 0; j < SCAN_ITEMS; j++) { const u64 i = base + j; key[j] = i < n ? k0[i] : 0; ridx[j] = i < n ? g.idx[i] : 0; }
    }
    // predecessor of the thread's first key: the previous lane's last key (lane 0 reads it)
    u64 prev = __shfl_up_sync(0xffffffffu, key[SCAN_ITEMS - 1], 1);
    if (lane_id() == 0) prev = (base > 0 && base <= n) ? k0[base - 1] : 0;
    // the gathers are independent of the scan: issue them now
    i32 sc_[SCAN_ITEMS];
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; j++) sc_[j] = (base + j < n && g.score) ? __ldg(g.score + ridx[j]) : 0;
    u32 fmask = 0, bmask = 0;                                // bit j: element j starts a unique / a bucket
    {
        u64 p = prev;
#pragma unroll
        for (int j = 0; j < SCAN_ITEMS; j++) {
            const u64 i = base + j;
            if (i < n && (i == 0 || key[j] != p)) fmask |= 1u << j;
            if (i == 0 || (key[j] >> g.sk.umi_bits) != (p >> g.sk.umi_bits)) bmask |= 1u << j;
            p = key[j];
        }
    }
    u32 total;
    u32 ex = block_exclusive_scan<u32, SCAN_THREADS>((u32)__popc(fmask), sm, &total) + tile_sums[blockIdx.x];
    const u64 cmask = (1ull << g.sk.umi_bits) - 1;          // one-word keys: umi_bits < 64
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; j++) {
        const u64 i = base + j;
        const u32 fl = (fmask >> j) & 1u;
        if (i < n) {
            const u32 uid = ex + fl - 1;
            const u32 r = ridx[j];
            if (fl) { g.useg[uid] = (u32)i; g.ucode[uid] = key[j] & cmask; g.bhead[uid] = (u8)((bmask >> j) & 1u); }
            if (i == n - 1) g.useg[uid + 1] = (u32)n;
            const u32 sv = g.score ? (u32)sc_[j] ^ 0x80000000u : 0u;
            const unsigned long long pk = ((unsigned long long)sv << 32) | (u32)~r;
            const i32 wv = g.weight ? __ldg(g.weight + r) : 0;          // weighted pushes are the rare API path
            if (uid == g.pend_uid) { g.pend_val = pk > g.pend_val ? pk : g.pend_val; g.pend_w += wv; }
            else { g.flush(); g.pend_uid = uid; g.pend_val = pk; g.pend_w = wv; }
            if (g.read_uid) g.read_uid[r] = uid;
        }
        ex += fl;
    }
    g.flush();
}

// The merge scan in ONE pass over the sorted keys (one-word keys, no weights): tiles take tickets, publish their number of
// unique heads and look back over their predecessors' (decoupled look-back, a warp reads 32 predecessors at a time), so the
// separate head-count pass — a second 8 B/read sweep — is gone.  rep[] needs no memset either: a tile zeroes the slots of the
// uniques that START in it before its own atomics touch them.  Reads that continue the last unique of an EARLIER tile (the
// ones before the tile's first head) never touch rep[] here — that slot belongs to a tile that may not have zeroed it yet —
// they are folded into one (unique id, value) pair per tile which unique_carry_kernel applies afterwards.
#define UQ_FLAG_AGG    (1ull << 62)
#define UQ_FLAG_PREFIX (2ull << 62)
#define UQ_FLAG_MASK   (3ull << 62)
__global__ void __launch_bounds__(SCAN_THREADS, 4) unique_onepass_kernel(UniqueEmit g_in, u64 n, u32 ntiles, unsigned long long *tile_state, u32 *ticket,
                                                                      u32 *__restrict__ carry_uid, unsigned long long *__restrict__ carry_val,
                                                                      u32 *n_unique_out, u32 *err) {
    static_assert(SCAN_ITEMS == 8, "vector loads below assume 8 elements per thread");
    UniqueEmit g = g_in;
    __shared__ u32 sm[SCAN_THREADS / 32 + 1];
    __shared__ u32 s_tile, s_excl;
    __shared__ unsigned long long s_carry;
    if (threadIdx.x == 0) { s_tile = atomicAdd(ticket, 1u); s_carry = 0ull; }
    __syncthreads();
    const u32 tile = s_tile;
    if (tile >= ntiles) return;
    const u64 base = (u64)tile * SCAN_TILE + (u64)threadIdx.x * SCAN_ITEMS;
    const u64 *__restrict__ k0 = g.sk.k0;
    u64 key[SCAN_ITEMS]; u32 ridx[SCAN_ITEMS];
    if (base + SCAN_ITEMS <= n) {
        const ulonglong2 *kv = reinterpret_cast<const ulonglong2 *>(k0 + base);
        const uint4 *iv = reinterpret_cast<const uint4 *>(g.idx + base);
#pragma unroll
        for (int q = 0; q < 4; q++) { const ulonglong2 t = __ldg(kv + q); key[2 * q] = t.x; key[2 * q + 1] = t.y; }
#pragma unroll
        for (int q = 0; q < 2; q++) { const uint4 t = __ldg(iv + q); ridx[4 * q] = t.x; ridx[4 * q + 1] = t.y; ridx[4 * q + 2] = t.z; ridx[4 * q + 3] = t.w; }
    } else {
#pragma unroll
        for (int j = 0; j < SCAN_ITEMS; j++) { const u64 i = base + j; key[j] = i < n ? k0[i] : 0; ridx[j] = i < n ? g.idx[i] : 0; }
    }
    u64 prev = __shfl_up_sync(0xffffffffu, key[SCAN_ITEMS - 1], 1);
    if (lane_id() == 0) prev = (base > 0 && base <= n) ? k0[base - 1] : 0;
    i32 sc_[SCAN_ITEMS];
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; j++) sc_[j] = (base + j < n && g.score) ? __ldg(g.score + ridx[j]) : 0;
    u32 fmask = 0, bmask = 0;                                // bit j: element j starts a unique / a bucket
    {
        u64 p = prev;
#pragma unroll
        for (int j = 0; j < SCAN_ITEMS; j++) {
            const u64 i = base + j;
            if (i < n && (i == 0 || key[j] != p)) fmask |= 1u << j;
            if (i == 0 || (key[j] >> g.sk.umi_bits) != (p >> g.sk.umi_bits)) bmask |= 1u << j;
            p = key[j];
        }
    }
    u32 total;
    u32 ex = block_exclusive_scan<u32, SCAN_THREADS>((u32)__popc(fmask), sm, &total);
    // ---- look-back (warp 0): exclusive prefix of `total` over the tiles before this one ----
    if (threadIdx.x < 32) {
        const u32 lane = lane_id();
        u32 excl = 0;
        if (tile != 0) {
            if (lane == 0) atomicExch(tile_state + tile, UQ_FLAG_AGG | (unsigned long long)total);
            long long p = (long long)tile - 1;              // highest predecessor not yet accounted for
            u32 spins = 0;
            for (;;) {
                const long long t = p - (long long)lane;
                unsigned long long v = UQ_FLAG_PREFIX;      // before tile 0: prefix 0
                if (t >= 0) v = *reinterpret_cast<volatile unsigned long long *>(tile_state + t);
                const u32 m_empty = __ballot_sync(0xffffffffu, (v & UQ_FLAG_MASK) == 0ull);
                const u32 m_pref = __ballot_sync(0xffffffffu, (v & UQ_FLAG_MASK) == UQ_FLAG_PREFIX);
                const u32 fe = m_empty ? (u32)__ffs(m_empty) - 1u : 32u, fp = m_pref ? (u32)__ffs(m_pref) - 1u : 32u;
                const u32 upto = min(fe, fp + 1u);          // lanes [0, upto) hold published values up to (and including) a prefix
                excl += __reduce_add_sync(0xffffffffu, lane < upto ? (u32)(v & ~UQ_FLAG_MASK) : 0u);
                if (fp < fe) break;
                p -= (long long)upto;
                if (upto == 0u) { if (++spins > (1u << 22)) { if (lane == 0) err[0] = 1; break; } __nanosleep(40); }      // a would-be hang becomes a reported error
            }
        }
        if (lane == 0) {
            atomicExch(tile_state + tile, UQ_FLAG_PREFIX | (unsigned long long)(excl + total));
            s_excl = excl;
            if (tile == ntiles - 1) *n_unique_out = excl + total;
        }
    }
    __syncthreads();
    const u32 excl = s_excl;
    for (u32 u = threadIdx.x; u < total; u += SCAN_THREADS) g.rep[excl + u] = 0ull;
    __syncthreads();
    ex += excl;
    const u64 cmask = (1ull << g.sk.umi_bits) - 1;          // one-word keys: umi_bits < 64
    u32 pend_uid = 0xffffffffu; unsigned long long pend_val = 0ull;
#define UQ_FLUSH() do { if (pend_uid != 0xffffffffu) { if (pend_uid + 1u == excl) atomicMax(&s_carry, pend_val); else atomicMax(&g.rep[pend_uid], pend_val); } } while (0)
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; j++) {
        const u64 i = base + j;
        const u32 fl = (fmask >> j) & 1u;
        if (i < n) {
            const u32 uid = ex + fl - 1;
            const u32 r = ridx[j];
            if (fl) { g.useg[uid] = (u32)i; g.ucode[uid] = key[j] & cmask; g.bhead[uid] = (u8)((bmask >> j) & 1u); }
            if (i == n - 1) g.useg[uid + 1] = (u32)n;
            const u32 sv = g.score ? (u32)sc_[j] ^ 0x80000000u : 0u;
            const unsigned long long pk = ((unsigned long long)sv << 32) | (u32)~r;
            if (uid == pend_uid) pend_val = pk > pend_val ? pk : pend_val;
            else { UQ_FLUSH(); pend_uid = uid; pend_val = pk; }
            if (g.read_uid) g.read_uid[r] = uid;
        }
        ex += fl;
    }
    UQ_FLUSH();
#undef UQ_FLUSH
    __syncthreads();
    if (threadIdx.x == 0) { carry_uid[tile] = excl - 1u; carry_val[tile] = s_carry; }
}

// the contributions a tile held back for the unique that started before it (0 = none: a packed value is never 0)
__global__ void __launch_bounds__(256) unique_carry_kernel(u32 ntiles, const u32 *__restrict__ carry_uid, const unsigned long long *__restrict__ carry_val,
                                                           unsigned long long *__restrict__ rep) {
    const u32 t = blockIdx.x * 256 + threadIdx.x;
    if (t >= ntiles) return;
    const unsigned long long v = carry_val[t];
    if (v) atomicMax(&rep[carry_uid[t]], v);
}

// per unique: freq, directional threshold (directional.rs:38), representative read, initial label.
// label = (~freq << 32 | unique id): ascending label = the reference's visit order (freq descending,
// directional.rs:67-72) with the canonical tie-break (UMI ascending = unique id ascending in a bucket).
__global__ void __launch_bounds__(256) unique_finalize_kernel(
    u32 n_unique, const u32 *__restrict__ useg, const unsigned long long *__restrict__ rep, const i32 *__restrict__ wsum,
    float percentage, int algo_inf_thr, i32 *__restrict__ freq, i32 *__restrict__ thr, u32 *__restrict__ rep_idx,
    unsigned long long *__restrict__ label, const u64 *__restrict__ ucode, int L, int has_n, uint2 *__restrict__ planes,
    u32 *__restrict__ nplane, const u64 *__restrict__ umi2 = nullptr, const u32 *__restrict__ nmask = nullptr) {
    u32 u = blockIdx.x * 256 + threadIdx.x;
    if (u >= n_unique) return;
    if (umi2) {
        // wide codes (N present, more than 21 nt: 3 bits per base do not fit ucode): the planes come from the representative
        // read's 2-bit code and N mask — same position convention (base b at plane bit L-1-b)
        const u32 r = ~(u32)rep[u];
        const u64 c = umi2[r];
        u32 p0 = 0, p1 = 0;
        for (int b = 0; b < L; b++) { p0 |= (u32)((c >> (2 * b)) & 1ull) << b; p1 |= (u32)((c >> (2 * b + 1)) & 1ull) << b; }
        const u32 pn = nmask[r];
        planes[u] = make_uint2(p0 & ~pn, p1 & ~pn);
        nplane[u] = pn;
    } else {
        u32 p0, p1, pn;
        code_to_planes(ucode[u], L, has_n, p0, p1, pn);
        planes[u] = make_uint2(p0, p1);
        if (has_n) nplane[u] = pn;
    }
    i32 f = wsum ? wsum[u] : (i32)(useg[u + 1] - useg[u]);
    freq[u] = f;
    thr[u] = algo_inf_thr ? 0x7fffffff : dir_threshold(percentage, f);
    rep_idx[u] = ~(u32)rep[u];
    label[u] = ((unsigned long long)(u32)~(u32)f << 32) | u;
}

struct BucketHead { const u8 *bhead; __device__ u32 operator()(u64 u) const { return bhead[u]; } };
struct BucketEmit {
    u32 *bstart; u64 n_unique; u32 *ubkt;      // ubkt[u] = bucket index of unique u
    __device__ void operator()(u64 u, u32 flag, u32 ex) const {
        if (flag) bstart[ex] = (u32)u;
        if (u == n_unique - 1) bstart[ex + flag] = (u32)n_unique;
        ubkt[u] = ex + flag - 1;
    }
};

// ---- multi-index (pigeonhole) candidate generation for big buckets ---------------------------------------
// A pair within Hamming distance k agrees exactly on at least one of k+1 disjoint parts of the UMI.  In the main
// order (sorted by the whole code) pairs that agree on the top part are close together, but pairs whose mismatch is
// IN the top part are far apart and cost most of the block pairs that survive letter-set culling.  So a big bucket
// is processed in k+1 passes: pass q uses an order in which part q is the most significant, only considers blocks
// whose part-q letter sets intersect at every position (= value ranges overlap: near-diagonal), and reports a pair
// iff q is the FIRST part on which the two UMIs agree — every pair within k is reported exactly once.
#define MI_MAX_PARTS 4
#define MI_BIG 4096u        // buckets with more unique UMIs than this use the multi-index passes
struct MiParams {
    int part;                       // -1: off;  q: big buckets are filtered on part q
    u32 big;                        // "big" threshold on the bucket's unique count
    u32 pmask[MI_MAX_PARTS];        // plane-position mask of each part
    unsigned long long cmask[MI_MAX_PARTS];   // code-bit mask of each part
};
__device__ __forceinline__ bool mi_accept(const MiParams &mi, unsigned long long ca, unsigned long long cb) {
    const unsigned long long x = ca ^ cb;
    if (x & mi.cmask[mi.part]) return false;                 // must agree on this pass's part
    for (int q = 0; q < mi.part; q++) if (!(x & mi.cmask[q])) return false;   // ... and on no earlier one
    return true;
}
struct BucketIsBig { const u32 *bstart; u32 big; u32 skip; __device__ u32 operator()(u64 b) const { return (b != skip && bstart[b + 1] - bstart[b] > big) ? 1u : 0u; } };
struct BucketBigEmit {
    u32 *brank, *big_bid; u64 n_buckets;
    __device__ void operator()(u64 b, u32 flag, u32 ex) const { brank[b] = flag ? ex : 0xffffffffu; if (flag) big_bid[ex] = (u32)b; }
};
struct BigSize { const u32 *bstart, *big_bid; __device__ u32 operator()(u64 r) const { u32 b = big_bid[r]; return bstart[b + 1] - bstart[b]; } };
struct BigStartEmit {
    u32 *bstart_big; u64 nbig;
    __device__ void operator()(u64 r, u32 v, u32 ex) const { bstart_big[r] = ex; if (r == nbig - 1) bstart_big[r + 1] = ex + v; }
};
// uniques of big buckets, compacted in main order; key of pass q = (big-bucket rank, value of part q)
__global__ void __launch_bounds__(256) mi_keys_kernel(u32 n_unique, const u32 *__restrict__ ubkt, const u32 *__restrict__ brank,
                                                      const u32 *__restrict__ bstart, const u32 *__restrict__ bstart_big,
                                                      const u64 *__restrict__ ucode, int shift, unsigned long long vmask, int pbits,
                                                      u32 *__restrict__ big_uid, u64 *__restrict__ keys) {
    u32 u = blockIdx.x * 256 + threadIdx.x;
    if (u >= n_unique) return;
    u32 b = ubkt[u], r = brank[b];
    if (r == 0xffffffffu) return;
    u32 m = bstart_big[r] + (u - bstart[b]);
    big_uid[m] = u;
    keys[m] = ((u64)r << pbits) | ((ucode[u] >> shift) & vmask);
}
__global__ void __launch_bounds__(256) mi_gather_kernel(u32 m_total, const u32 *__restrict__ perm, const u32 *__restrict__ big_uid,
                                                        const uint2 *__restrict__ planes, const u32 *__restrict__ nplane,
                                                        const u64 *__restrict__ ucode, uint2 *__restrict__ planes_q, u32 *__restrict__ nplane_q,
                                                        u64 *__restrict__ ucode_q, u32 *__restrict__ uid_q) {
    u32 m = blockIdx.x * 256 + threadIdx.x;
    if (m >= m_total) return;
    u32 u = big_uid[perm[m]];
    planes_q[m] = planes[u]; ucode_q[m] = ucode[u]; uid_q[m] = u;
    if (nplane) nplane_q[m] = nplane[u];
}

// Tile geometry of the neighbour search (rows x cols of unique UMIs of one bucket).
#define HT_ROWS 2048
#define SMALL_BUCKET 32   // buckets up to this many unique UMIs are handled one per warp
#define HT_COLS 2048

// cnts = row_cnt | col_cnt << 12 | diag << 31 (counts <= 2048); col_blk0 = global id of the column tile's first
// 128-block (the row tile's is col_blk0 - (col_start - row_start) / 128: same bucket, 128-aligned)
// multi-index pass q orders a big bucket by its part-q value first, so two blocks (or tiles) can only hold a pair that
// agrees on part q if their part-q value RANGES overlap: with rows before columns in the order, iff last(rows) >=
// first(cols).  The summaries keep the (top 32 bits of the) part value of their first and last UMI in words 5 and 6.
__device__ __forceinline__ u32 mi_part_value32(const MiParams &mi, unsigned long long code) {
    const unsigned long long m = mi.cmask[mi.part];
    const int sh = __ffsll((long long)m) - 1, bits = __popcll(m);
    unsigned long long v = (code & m) >> sh;
    if (bits > 32) v >>= (bits - 32);
    return (u32)v;
}
struct TileItem { u32 row_start, col_start, cnts, col_blk0; };
__device__ __forceinline__ u32 item_row_cnt(const TileItem &it) { return it.cnts & 0xfffu; }
__device__ __forceinline__ u32 item_col_cnt(const TileItem &it) { return (it.cnts >> 12) & 0xfffu; }
__device__ __forceinline__ bool item_diag(const TileItem &it) { return it.cnts >> 31; }
__device__ __forceinline__ bool item_filtered(const TileItem &it) { return (it.cnts >> 30) & 1u; }

__device__ __forceinline__ u32 bucket_tiles(u32 nb) { return (nb + HT_ROWS - 1) / HT_ROWS; }

// number of tile pairs (ti <= tj) of a bucket; buckets with one UMI need no comparison at all
// `skip` = a bucket left out of this device's search (sharded run: the hot bucket is searched by the whole group)
struct BucketItems {
    const u32 *bstart; u32 skip;
    __device__ u32 operator()(u64 b) const {
        u32 nb = bstart[b + 1] - bstart[b];
        if (nb <= SMALL_BUCKET || b == skip) return 0;          // 0/1 UMIs: nothing to compare; 2..32: small_buckets_kernel
        u32 t = bucket_tiles(nb);
        return t * (t + 1) / 2;
    }
};
struct BucketItemsEmit {
    u32 *item_off; u64 n_buckets;
    __device__ void operator()(u64 b, u32 v, u32 ex) const {
        item_off[b] = ex;
        if (b == n_buckets - 1) item_off[b + 1] = ex + v;
    }
};

// Grid-stride over the buckets; n_ptr (optional) = device-resident bucket count, so that the launch need not wait for it.
// Also counts the big buckets (multi-index passes) and their unique UMIs: one read-back serves the work-list sizing.
__global__ void __launch_bounds__(256) bucket_stats_kernel(u32 n_buckets, const u32 *__restrict__ n_ptr, const u32 *__restrict__ bstart, DevScalars *sc) {
    if (n_ptr) n_buckets = *n_ptr;
    u64 pairs = 0;
    u32 mx = 0, nbig = 0, mbig = 0;
    for (u64 b = (u64)blockIdx.x * 256 + threadIdx.x; b < n_buckets; b += (u64)gridDim.x * 256) {
        const u32 nb = bstart[b + 1] - bstart[b];
        pairs += (u64)nb * (nb ? nb - 1 : 0) / 2;
        mx = max(mx, nb);
        if (nb > MI_BIG) { nbig++; mbig += nb; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        pairs += __shfl_xor_sync(0xffffffffu, pairs, o);
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        nbig += __shfl_xor_sync(0xffffffffu, nbig, o);
        mbig += __shfl_xor_sync(0xffffffffu, mbig, o);
    }
    if (lane_id() == 0 && mx) {
        atomicAdd((unsigned long long *)&sc->pairs, (unsigned long long)pairs);
        atomicMax(&sc->max_umis, mx);
        if (nbig) { atomicAdd(&sc->n_big_all, nbig); atomicAdd(&sc->m_big_all, mbig); }
    }
}

// ---- exact tile culling ------------------------------------------------------------------------
// Unique UMIs of a bucket are sorted, so a tile of 2048 consecutive UMIs shares a long prefix and uses
// few letters at the next positions.  Per tile we keep, for each letter x, a mask S_x whose bit j says
// "some UMI of the tile has letter x at position j".  For two tiles the number of positions whose
// letter sets are disjoint is a lower bound of the Hamming distance of every cross pair; when it
// exceeds k the tile pair cannot contain a neighbour and is never scheduled.  (Exact: no pair within
// k is ever dropped; tests compare culled and unculled runs with the oracle.)
#define TS_WORDS 8      // letter-set words per tile (A C G T N + padding)

struct BucketTiles {
    const u32 *bstart; u32 skip;
    __device__ u32 operator()(u64 b) const { u32 nb = bstart[b + 1] - bstart[b]; return (nb <= SMALL_BUCKET || b == skip) ? 0 : bucket_tiles(nb); }
};
struct BucketTilesEmit {
    u32 *tile_off; u64 n_buckets;
    __device__ void operator()(u64 b, u32 v, u32 ex) const {
        tile_off[b] = ex;
        if (b == n_buckets - 1) tile_off[b + 1] = ex + v;
    }
};

// 128-UMI blocks of the buckets that go through the tile / block path (global block id = blk_off[b] + local block)
struct BucketBlocks {
    const u32 *bstart; u32 skip;
    __device__ u32 operator()(u64 b) const { u32 nb = bstart[b + 1] - bstart[b]; return (nb <= SMALL_BUCKET || b == skip) ? 0 : (nb + 127) / 128; }
};

__device__ __forceinline__ void onehot_planes(uint2 p, u32 pn, u32 lmask, u32 *oh /*5*/) {
    u32 base = lmask & ~pn;
    oh[0] = ~p.y & ~p.x & base; oh[1] = ~p.y & p.x & base; oh[2] = p.y & ~p.x & base; oh[3] = p.y & p.x & base; oh[4] = pn & lmask;
}

// one warp per tile: letter sets of each 128-UMI block of the tile (bsum) and of the whole tile (tsum)
#define BLOCKS_PER_TILE (HT_COLS / 128)
__global__ void __launch_bounds__(256) tile_summary_kernel(u32 n_tiles, u32 n_buckets, const u32 *__restrict__ tile_off,
                                                           const u32 *__restrict__ bstart, const uint2 *__restrict__ planes,
                                                           const u32 *__restrict__ nplane, int L, u32 *__restrict__ tsum,
                                                           const u32 *__restrict__ blk_off, u32 *__restrict__ bsum,
                                                           u32 *__restrict__ blk_first, u32 *__restrict__ blk_cnt,
                                                           const u64 *__restrict__ ucode, MiParams mi) {
    u32 t = (blockIdx.x * 256 + threadIdx.x) >> 5;
    if (t >= n_tiles) return;
    u32 lo = 0, hi = n_buckets;
    while (hi - lo > 1) { u32 mid = (lo + hi) >> 1; if (tile_off[mid] <= t) lo = mid; else hi = mid; }
    u32 b = lo, ti = t - tile_off[b], s = bstart[b], nb = bstart[b + 1] - s;
    u32 first = s + ti * HT_ROWS, cnt = min((u32)HT_ROWS, nb - ti * HT_ROWS);
    u32 lmask = L >= 32 ? 0xffffffffu : ((1u << L) - 1);
    u32 tot[5] = {0, 0, 0, 0, 0};
    const bool filt = mi.part >= 0 && nb > mi.big;
    const u32 gb0 = blk_off[b] + ti * BLOCKS_PER_TILE;       // global id of the tile's first 128-block
    for (u32 blk = 0; blk * 128 < cnt; blk++) {
        u32 acc[5] = {0, 0, 0, 0, 0};
        for (u32 i = blk * 128 + lane_id(); i < min(cnt, (blk + 1) * 128); i += 32) {
            u32 oh[5];
            onehot_planes(planes[first + i], nplane ? nplane[first + i] : 0u, lmask, oh);
#pragma unroll
            for (int x = 0; x < 5; x++) acc[x] |= oh[x];
        }
#pragma unroll
        for (int x = 0; x < 5; x++) { acc[x] = __reduce_or_sync(0xffffffffu, acc[x]); tot[x] |= acc[x]; }
        if (lane_id() < TS_WORDS) {
            u32 v = 0;
#pragma unroll
            for (int x = 0; x < 5; x++) if (lane_id() == (u32)x) v = acc[x];
            if (filt && lane_id() == 5) v = mi_part_value32(mi, ucode[first + blk * 128]);
            if (filt && lane_id() == 6) v = mi_part_value32(mi, ucode[first + min(cnt, (blk + 1) * 128) - 1]);
            bsum[((u64)gb0 + blk) * TS_WORDS + lane_id()] = v;
        }
        if (lane_id() == 0) { blk_first[gb0 + blk] = first + blk * 128; blk_cnt[gb0 + blk] = min(128u, cnt - blk * 128); }
    }
    if (lane_id() < TS_WORDS) {
        u32 v = 0;
#pragma unroll
        for (int x = 0; x < 5; x++) if (lane_id() == (u32)x) v = tot[x];
        if (filt && lane_id() == 5) v = mi_part_value32(mi, ucode[first]);
        if (filt && lane_id() == 6) v = mi_part_value32(mi, ucode[first + cnt - 1]);
        tsum[(u64)t * TS_WORDS + lane_id()] = v;
    }
}

__device__ __forceinline__ u32 disjoint_positions(const u32 *a, const u32 *b, u32 lmask) {
    u32 t = (a[0] & b[0]) | (a[1] & b[1]) | (a[2] & b[2]) | (a[3] & b[3]) | (a[4] & b[4]);
    return __popc(~t & lmask);
}

// one thread per candidate tile pair: bucket by binary search over item_off, (ti, tj) by triangular
// decode; survivors of the cull test are appended (order is irrelevant: the edge SET is what matters)
__global__ void __launch_bounds__(256) build_items_kernel(u32 n_cand, u32 n_buckets, const u32 *__restrict__ item_off,
                                                          const u32 *__restrict__ bstart, const u32 *__restrict__ tile_off,
                                                          const u32 *__restrict__ blk_off, const u32 *__restrict__ tsum, int L, int k, int cull,
                                                          MiParams mi, TileItem *__restrict__ items, DevScalars *sc, u32 band, u32 n_bands) {
    u32 w = blockIdx.x * 256 + threadIdx.x;
    u64 npairs = 0;
    bool live = false;
    TileItem it;
    if (w < n_cand) {
        u32 lo = 0, hi = n_buckets;            // last b with item_off[b] <= w
        while (hi - lo > 1) { u32 mid = (lo + hi) >> 1; if (item_off[mid] <= w) lo = mid; else hi = mid; }
        u32 b = lo, local = w - item_off[b];
        u32 s = bstart[b], nb = bstart[b + 1] - s, t = bucket_tiles(nb);
        // row-major upper triangle: row ti holds (t - ti) items
        u32 ti = 0;
        {
            double tt = 2.0 * t + 1.0;
            double r = (tt - sqrt(tt * tt - 8.0 * (double)local)) * 0.5;
            ti = (u32)r; if (ti >= t) ti = t - 1;
            // first item index of row ti = ti*t - ti*(ti-1)/2
            while (ti > 0 && (u64)ti * t - (u64)ti * (ti - 1) / 2 > local) ti--;
            while ((u64)(ti + 1) * t - (u64)(ti + 1) * ti / 2 <= local) ti++;
        }
        u32 tj = ti + (local - (u32)((u64)ti * t - (u64)ti * (ti - 1) / 2));
        it.row_start = s + ti * HT_ROWS;
        it.col_start = s + tj * HT_COLS;
        u32 rc = min((u32)HT_ROWS, nb - ti * HT_ROWS);
        u32 cc = min((u32)HT_COLS, nb - tj * HT_COLS);
        const bool filt = mi.part >= 0 && nb > mi.big;         // multi-index pass: big buckets only see near-diagonal blocks
        it.cnts = rc | (cc << 12) | (filt ? 0x40000000u : 0u) | (ti == tj ? 0x80000000u : 0u);
        it.col_blk0 = blk_off[b] + tj * BLOCKS_PER_TILE;
        // sharded hot bucket: the row tiles are dealt out to the devices of the group round robin (near-diagonal survivors
        // make every row tile about equally expensive)
        live = n_bands <= 1 || (ti % n_bands) == band;
        if (!live) {
        } else if (cull && ti != tj) {
            u32 lmask = L >= 32 ? 0xffffffffu : ((1u << L) - 1);
            const u32 *a = tsum + (u64)(tile_off[b] + ti) * TS_WORDS, *c = tsum + (u64)(tile_off[b] + tj) * TS_WORDS;
            live = disjoint_positions(a, c, lmask) <= (u32)k;
            if (live && filt) live = disjoint_positions(a, c, mi.pmask[mi.part]) == 0;
            if (live) npairs = (u64)rc * cc;          // the pair space the density heuristic compares with: before the range test
            if (live && filt) live = a[6] >= c[5];    // part-q value ranges overlap
        } else npairs = (u64)rc * cc;
    }
    // warp-aggregated append
    u32 m = __ballot_sync(0xffffffffu, live);
    u32 base = 0;
    if (lane_id() == 0 && m) base = atomicAdd(&sc->n_items, (u32)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (live) items[base + __popc(m & lanemask_lt())] = it;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) npairs += __shfl_xor_sync(0xffffffffu, npairs, o);
    if (lane_id() == 0 && npairs) atomicAdd((unsigned long long *)&sc->scratch2, (unsigned long long)npairs);
}
