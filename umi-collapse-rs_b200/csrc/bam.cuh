// bam.cuh — host feed on the device: raw BAM alignment records
// (the bytes of a BGZF-inflated BAM stream) -> the SoA the hot path consumes.  Replaces, per record,
//   get_unclipped_pos        src/utils/mod.rs:96-104   (rust-htslib CigarStringView semantics, BAM spec)
//   UcSAMRead::get_umi       src/utils/read.rs:96-111  (first separator byte, then umi_length bytes)
//   to_bitset                src/utils/mod.rs:63-83    (packed straight to 2 bit/base + N mask)
//   UcSAMRead::new avg_qual  src/utils/read.rs:56-63   /  get_map_qual :77-79
//   the unmapped filter      src/deduplicate_sam.rs:102-108
#pragma once
#include "common.cuh"

struct BamDecodeOut {
    i32 *tid; i64 *pos; u8 *rev; u64 *umi2; u32 *nmask; i32 *score; u8 *valid; i64 *tlen;
};

// error bits accumulated in DevScalars.bam_err
#define BAM_ERR_NO_SEP   1u    // reference: panic!("failed to get the umi"), utils/read.rs:109
#define BAM_ERR_SHORT    2u    // reference: slice index panic, utils/read.rs:101
#define BAM_ERR_BAD_BASE 4u    // reference: panic!("Unknown character in UMI sequence"), utils/mod.rs:78
#define BAM_ERR_TRUNC    8u    // record runs past its block

__device__ __forceinline__ u32 ld_u16(const u8 *p) { return (u32)p[0] | ((u32)p[1] << 8); }
__device__ __forceinline__ u32 ld_u32(const u8 *p) { return (u32)p[0] | ((u32)p[1] << 8) | ((u32)p[2] << 16) | ((u32)p[3] << 24); }

// per-record classes counted by the paired-end filters
#define BAM_CLS_MATE     1u    // last-in-template record skipped before total_read_count, deduplicate_sam.rs:96-98
#define BAM_CLS_UNMAPPED 2u    // unmapped, or (paired mode) mate unmapped: :102-108, :118-121
#define BAM_CLS_UNPAIRED 4u    // :111-116 (counted whether or not it is removed)
#define BAM_CLS_CHIMERIC 8u    // :123-128 (counted whether or not it is removed)

// decodes record i; returns its BAM_CLS_* bits
__device__ __forceinline__ u32 bam_decode_one(u64 i, const u8 *__restrict__ buf, const u64 *__restrict__ offsets,
                                              int umi_len, u32 sep, int use_mapq, int paired, int remove_unpaired,
                                              int remove_chimeric, const BamDecodeOut &o, u32 *err) {
    u32 cls = 0;
    const u8 *r = buf + offsets[i];
    const u64 rec_len = offsets[i + 1] - offsets[i];
    u32 e = 0;
    const u32 block_size = ld_u32(r);
    if (rec_len < 36 || (u64)block_size + 4 > rec_len) { o.valid[i] = 0; atomicOr(err, BAM_ERR_TRUNC); return 0; }
    const i32 ref_id = (i32)ld_u32(r + 4);
    const i32 pos = (i32)ld_u32(r + 8);
    const u32 l_read_name = r[12], mapq = r[13];
    const u32 n_cigar = ld_u16(r + 16), flag = ld_u16(r + 18);
    const u32 l_seq = ld_u32(r + 20);
    const u8 *qname = r + 36;
    const u8 *cigar = qname + l_read_name;
    const u8 *qual = cigar + 4 * (u64)n_cigar + ((l_seq + 1) >> 1);
    if ((u64)(qual - r) + l_seq > rec_len) { o.valid[i] = 0; atomicOr(err, BAM_ERR_TRUNC); return 0; }
    const bool unmapped = flag & 0x4, reverse = flag & 0x10;
    bool valid = !unmapped;                               // deduplicate_sam.rs:102-108
    if (unmapped) cls = BAM_CLS_UNMAPPED;
    if (paired) {                                        // deduplicate_sam.rs:96-129, in the reference's order
        const bool is_paired = flag & 0x1, last = flag & 0x80, mate_unmapped = flag & 0x8;
        const i32 mtid = (i32)ld_u32(r + 24);
        if (is_paired && last) { o.valid[i] = 0; return BAM_CLS_MATE; }
        if (valid && !is_paired) { cls |= BAM_CLS_UNPAIRED; if (remove_unpaired) valid = false; }
        if (valid && is_paired && mate_unmapped) { cls |= BAM_CLS_UNMAPPED; valid = false; }
        if (valid && is_paired && ref_id != mtid) { cls |= BAM_CLS_CHIMERIC; if (remove_chimeric) valid = false; }
    }
    o.valid[i] = valid ? 1 : 0;
    if (!valid) return cls;
    if (paired) o.tlen[i] = (i64)(i32)ld_u32(r + 32);    // record.insert_size()

    // ---- a1: unclipped position ----
    i64 up;
    if (!reverse) {
        i64 soft = 0, hard = 0; u32 c = 0;
        if (c < n_cigar) { u32 v = ld_u32(cigar); if ((v & 0xf) == 5) { hard = v >> 4; c = 1; } }
        if (c < n_cigar) { u32 v = ld_u32(cigar + 4 * c); if ((v & 0xf) == 4) soft = v >> 4; }
        up = (i64)pos - soft - hard;
    } else {
        i64 end = pos, soft = 0, hard = 0;
        for (u32 c = 0; c < n_cigar; c++) {
            u32 v = ld_u32(cigar + 4 * c), op = v & 0xf;
            if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) end += v >> 4;
        }
        i32 c = (i32)n_cigar - 1;
        if (c >= 0) { u32 v = ld_u32(cigar + 4 * c); if ((v & 0xf) == 5) { hard = v >> 4; c--; } }
        if (c >= 0) { u32 v = ld_u32(cigar + 4 * c); if ((v & 0xf) == 4) soft = v >> 4; }
        up = end - 1 + soft + hard;
    }
    // ---- a2 + a3: UMI after the first separator of the read name (the NUL terminator is not part of it) ----
    const u32 name_len = l_read_name ? l_read_name - 1 : 0;
    u32 p = 0;
    while (p < name_len && qname[p] != sep) p++;
    u64 code = 0; u32 nm = 0;
    if (p >= name_len) e |= BAM_ERR_NO_SEP;
    else if (p + 1 + (u32)umi_len > name_len) e |= BAM_ERR_SHORT;
    else {
        const u8 *s = qname + p + 1;
        for (int b = 0; b < umi_len; b++) {
            u32 c = s[b], v;
            if (c == 'A') v = 0; else if (c == 'C') v = 1; else if (c == 'G') v = 2; else if (c == 'T') v = 3;
            else if (c == 'N') { v = 0; nm |= 1u << (umi_len - 1 - b); }
            else { v = 0; e |= BAM_ERR_BAD_BASE; }
            code = (code << 2) | v;
        }
    }
    // ---- a4: score ----
    i32 score;
    if (use_mapq) score = (i32)mapq;
    else if (l_seq <= 65536) {
        u32 s = 0;
        for (u32 b = 0; b < l_seq; b++) s += qual[b];
        float q = __fdiv_rn(__uint2float_rn(s), __uint2float_rn(l_seq));
        score = (q != q) ? 0 : __float2int_rz(q);
    } else {
        float acc = 0.0f;
        for (u32 b = 0; b < l_seq; b++) acc = __fadd_rn(acc, (float)qual[b]);
        float q = __fdiv_rn(acc, __uint2float_rn(l_seq));
        score = (q != q) ? 0 : (q >= 2147483648.0f ? 0x7fffffff : __float2int_rz(q));
    }
    o.tid[i] = ref_id; o.pos[i] = up; o.rev[i] = reverse ? 1 : 0; o.umi2[i] = code; o.nmask[i] = nm; o.score[i] = score;
    if (e) atomicOr(err, e);
    return cls;
}

// one thread per record; the class counters are warp-aggregated (one atomic per warp and class)
__global__ void __launch_bounds__(128) bam_decode_kernel(u64 n, const u8 *__restrict__ buf, const u64 *__restrict__ offsets,
                                                         int umi_len, u32 sep, int use_mapq, int paired, int remove_unpaired,
                                                         int remove_chimeric, BamDecodeOut o, u32 *err, DevScalars *sc) {
    u64 i = (u64)blockIdx.x * 128 + threadIdx.x;
    u32 cls = 0;
    if (i < n) cls = bam_decode_one(i, buf, offsets, umi_len, sep, use_mapq, paired, remove_unpaired, remove_chimeric, o, err);
    const u32 m_mate = __ballot_sync(0xffffffffu, cls & BAM_CLS_MATE), m_unm = __ballot_sync(0xffffffffu, cls & BAM_CLS_UNMAPPED);
    const u32 m_unp = __ballot_sync(0xffffffffu, cls & BAM_CLS_UNPAIRED), m_chi = __ballot_sync(0xffffffffu, cls & BAM_CLS_CHIMERIC);
    if (lane_id() == 0) {
        if (m_mate) atomicAdd((unsigned long long *)&sc->n_mates_skipped, (unsigned long long)__popc(m_mate));
        if (m_unm) atomicAdd((unsigned long long *)&sc->n_bam_unmapped, (unsigned long long)__popc(m_unm));
        if (m_unp) atomicAdd((unsigned long long *)&sc->n_unpaired, (unsigned long long)__popc(m_unp));
        if (m_chi) atomicAdd((unsigned long long *)&sc->n_chimeric, (unsigned long long)__popc(m_chi));
    }
}

// stream compaction of the decoded records that pass the filter into the context's read arrays
struct BamValid { const u8 *valid; __device__ u32 operator()(u64 i) const { return valid[i]; } };
struct BamCompact {
    BamDecodeOut src, dst;      // dst pointers already offset to the append position
    u32 *orig;                  // dst: record number inside this push
    u32 *n_valid; u64 n;
    __device__ void operator()(u64 i, u32 flag, u32 ex) const {
        if (flag) {
            dst.tid[ex] = src.tid[i]; dst.pos[ex] = src.pos[i]; dst.rev[ex] = src.rev[i];
            dst.umi2[ex] = src.umi2[i]; dst.nmask[ex] = src.nmask[i]; dst.score[ex] = src.score[i];
            if (dst.tlen) dst.tlen[ex] = src.tlen[i];
            orig[ex] = (u32)i;
        }
        if (i == n - 1) *n_valid = ex + flag;
    }
};

// min/max of tid/pos and the N flag over freshly appended reads (what umi_pack_kernel does for ASCII pushes)
__global__ void __launch_bounds__(256) range_reduce_kernel(u64 n, const i32 *__restrict__ tid, const i64 *__restrict__ pos,
                                                           const i64 *__restrict__ tlen, const u32 *__restrict__ nmask, DevScalars *sc) {
    u64 i = (u64)blockIdx.x * 256 + threadIdx.x;
    i32 tmin = 0x7fffffff, tmax = (i32)0x80000000;
    i64 pmin = 0x7fffffffffffffffLL, pmax = (i64)0x8000000000000000LL;
    i64 lmin = 0x7fffffffffffffffLL, lmax = (i64)0x8000000000000000LL;
    u32 anyn = 0;
    if (i < n) { tmin = tmax = tid[i]; pmin = pmax = pos[i]; anyn = nmask[i] != 0; if (tlen) lmin = lmax = tlen[i]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        tmin = min(tmin, __shfl_xor_sync(0xffffffffu, tmin, o)); tmax = max(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
        pmin = min(pmin, __shfl_xor_sync(0xffffffffu, pmin, o)); pmax = max(pmax, __shfl_xor_sync(0xffffffffu, pmax, o));
        lmin = min(lmin, __shfl_xor_sync(0xffffffffu, lmin, o)); lmax = max(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
        anyn |= __shfl_xor_sync(0xffffffffu, anyn, o);
    }
    if (lane_id() == 0) {
        volatile DevScalars *vs = sc;
        if (tlen && lmin <= lmax) { atomicMin((long long *)&sc->tlen_min, (long long)lmin); atomicMax((long long *)&sc->tlen_max, (long long)lmax); }
        if (tmin < vs->tid_min) atomicMin(&sc->tid_min, tmin);
        if (tmax > vs->tid_max) atomicMax(&sc->tid_max, tmax);
        if (pmin < vs->pos_min) atomicMin((long long *)&sc->pos_min, (long long)pmin);
        if (pmax > vs->pos_max) atomicMax((long long *)&sc->pos_max, (long long)pmax);
        if (anyn && !vs->any_n) atomicOr(&sc->any_n, 1u);
    }
}
