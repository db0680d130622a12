// hamming_blocks.cuh — K5 hamming_neighbours, sparse block-pair form (production kernel).
//
// Same bit-sliced one-hot arithmetic as hamming_tiles_bs (hamming_bs.cuh), organised for the regime that
// exact culling leaves behind: of the N^2/2 pairs of a hot locus only ~1 % of the (128 x 128) blocks can
// contain a neighbour, so the work is a LIST of block pairs rather than dense tiles.
//   onehot_build_kernel   once per run: one-hot match words of every 128-UMI block, in HBM (768 B per block at
//                         12 nt): eq[block][j][x][4 groups], bit c of group g = "UMI 32g+c has letter x at position j"
//   expand_blocks_kernel  one warp per surviving tile pair: tests its 16 x 16 block pairs with the per-block
//                         letter sets (disjoint-positions bound) and appends the survivors
//   hamming_blocks        one warp per block pair: the warp stages the column block's words in its own slice of shared
//                         memory (768 B at 12 nt, __syncwarp only), then runs four slices of 32 consecutive rows; a slice
//                         whose letter sets are disjoint from the column block's in more than k positions is skipped;
//                         otherwise each lane runs its row UMI against the 128 columns: one LDS.128 (4 distinct 16-byte rows
//                         per warp: conflict-free) and 2(k+1) LOP3 per position.
#pragma once
#include "common.cuh"
#include "hamming.cuh"

#define BLK_COLS 128

// ---- one-hot words of every 128-block (built once per run) ----
template <bool HASN>
__global__ void __launch_bounds__(256) onehot_build_kernel(u32 n_groups /* 4 per block */, const u32 *__restrict__ blk_first,
                                                           const u32 *__restrict__ blk_cnt, const uint2 *__restrict__ planes,
                                                           const u32 *__restrict__ nplane, int L, int LP, u32 *__restrict__ eq,
                                                           const u8 *__restrict__ need /* nullable: only blocks that are a column of some pair */) {
    constexpr int XS = HASN ? 8 : 4;
    constexpr int NLET = HASN ? 5 : 4;
    // one warp per 128-UMI block: lane x collects the four 32-UMI groups' words of letter x and stores them as one
    // 16-byte vector, so every position is written as XS consecutive uint4
    const u32 gb = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = lane_id();
    if (gb >= (n_groups >> 2)) return;
    if (need && !need[gb]) return;
    const u32 cnt = blk_cnt[gb], first = blk_first[gb];
    uint2 p[4]; u32 pn[4];
#pragma unroll
    for (int g = 0; g < 4; g++) {
        const u32 c = g * 32 + lane;
        p[g] = c < cnt ? planes[first + c] : make_uint2(0u, 0u);
        pn[g] = (HASN && c < cnt) ? nplane[first + c] : 0u;
    }
    uint4 *dst = reinterpret_cast<uint4 *>(eq) + (u64)gb * LP * XS;
    u32 vmask[4];
#pragma unroll
    for (int g = 0; g < 4; g++) vmask[g] = __ballot_sync(0xffffffffu, (u32)(g * 32) + lane < cnt);   // padding columns match no letter
    for (int j = 0; j < LP; j++) {
        u32 v[4];
#pragma unroll
        for (int g = 0; g < 4; g++) {
            // two (three with N) ballots transpose bit j of the 32 UMIs' planes; lane x then forms letter x's match word
            const u32 b0 = __ballot_sync(0xffffffffu, (p[g].x >> j) & 1u), b1 = __ballot_sync(0xffffffffu, (p[g].y >> j) & 1u);
            const u32 bn = HASN ? __ballot_sync(0xffffffffu, (pn[g] >> j) & 1u) : 0u;
            u32 w = 0;
            if (lane < 4) w = ((lane & 2u) ? b1 : ~b1) & ((lane & 1u) ? b0 : ~b0) & ~bn;
            else if (HASN && lane == 4) w = bn;
            v[g] = j >= L ? 0xffffffffu : (w & vmask[g]);
        }
        if (lane < (u32)XS) dst[j * XS + lane] = make_uint4(v[0], v[1], v[2], v[3]);
    }
}

// ---- surviving tile pairs -> surviving (128 x 128) block pairs ----
// fill == 0: only counts (out_count += survivors); fill == 1: appends uint2(row block, col block) while there is room (cap)
__global__ void __launch_bounds__(256) expand_blocks_kernel(const TileItem *__restrict__ items, u32 n_items, const u32 *__restrict__ bsum,
                                                            int L, int k, int cull, MiParams mi, int fill, uint2 *__restrict__ pairs,
                                                            unsigned long long *out_count, u8 *__restrict__ need = nullptr,
                                                            unsigned long long cap = ~0ull /* fill: slots in pairs[]; the count runs on past it */,
                                                            const u32 *__restrict__ n_items_ptr = nullptr /* device-resident item count: no read-back before the launch */) {
    if (n_items_ptr) n_items = *n_items_ptr;
    const u32 lane = lane_id(), n_warps = (gridDim.x * 256) >> 5;
    for (u32 w = (blockIdx.x * 256 + threadIdx.x) >> 5; w < n_items; w += n_warps) {
        const TileItem it = items[w];
        const u32 nrb = (item_row_cnt(it) + 127) >> 7, ncb = (item_col_cnt(it) + 127) >> 7;
        const bool diag = item_diag(it), filt = item_filtered(it);
        const u32 row_blk0 = it.col_blk0 - ((it.col_start - it.row_start) >> 7);
        const u32 lmask = L >= 32 ? 0xffffffffu : ((1u << L) - 1);
        const u32 r = lane & 15, c0 = (lane >> 4) * 8;
        u32 rs[5] = {0, 0, 0, 0, 0}, r_last = 0;
        if (r < nrb) {
#pragma unroll
            for (int x = 0; x < 5; x++) rs[x] = bsum[(u64)(row_blk0 + r) * 8 + x];
            r_last = bsum[(u64)(row_blk0 + r) * 8 + 6];
        }
        u32 live = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const u32 c = c0 + i;
            bool ok = r < nrb && c < ncb && (!diag || r <= c);
            if (ok && cull && !(diag && r == c)) {
                u32 cs[5];
#pragma unroll
                for (int x = 0; x < 5; x++) cs[x] = bsum[(u64)(it.col_blk0 + c) * 8 + x];
                ok = disjoint_positions(rs, cs, lmask) <= (u32)k;
                if (ok && filt) ok = disjoint_positions(rs, cs, mi.pmask[mi.part]) == 0;     // part-q letter sets must intersect
                if (ok && filt) ok = r_last >= bsum[(u64)(it.col_blk0 + c) * 8 + 5];         // part-q value ranges must overlap
            }
            if (ok) live |= 1u << i;
        }
        u32 cnt = __popc(live);
        // warp exclusive prefix of cnt
        u32 inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { u32 t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= (u32)o) inc += t; }
        const u32 total = __shfl_sync(0xffffffffu, inc, 31);
        if (total == 0) continue;
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(out_count, (unsigned long long)total);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (fill) {
            u64 o = base + inc - cnt;
#pragma unroll
            for (int i = 0; i < 8; i++) if (live & (1u << i)) {
                if (o < cap) pairs[o] = make_uint2((row_blk0 + r) | (filt ? 0x80000000u : 0u), it.col_blk0 + c0 + i);
                o++;
                if (need) need[it.col_blk0 + c0 + i] = 1;
            }
        }
    }
}
// ---- warp-buffered edge output ----
// Every hit used to cost one atomicAdd on the single global edge counter; in a saturated UMI space (the fastq
// single-bucket config: ~170 M edges) that one address serialises the whole kernel.  Each warp now stages its
// edges in a private shared-memory buffer (positions by ballot/prefix, no atomics) and reserves global space
// with ONE atomicAdd per flush of up to WB_CAP edges.
#define WB_CAP 128
struct WarpEdgeBuf {
    uint2 *buf;      // this warp's WB_CAP slots in shared memory
    u32 fill;        // warp-uniform
};
__device__ __forceinline__ void wb_flush(WarpEdgeBuf &wb, const EdgeSink &es) {
    if (wb.fill == 0) return;
    __syncwarp();
    unsigned long long base = 0;
    if (lane_id() == 0) base = atomicAdd(es.count, (unsigned long long)wb.fill);
    base = __shfl_sync(0xffffffffu, base, 0);
    for (u32 i = lane_id(); i < wb.fill; i += 32) if (base + i < es.cap) es.edges[base + i] = wb.buf[i];
    __syncwarp();
    wb.fill = 0;
}
// called by all 32 lanes; ab / ba = this lane emits edge (a -> b) / (b -> a)
__device__ __forceinline__ void wb_emit(WarpEdgeBuf &wb, const EdgeSink &es, bool ab, bool ba, u32 a, u32 b) {
    const u32 cnt = (ab ? 1u : 0u) + (ba ? 1u : 0u);
    u32 inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { u32 t = __shfl_up_sync(0xffffffffu, inc, o); if (lane_id() >= (u32)o) inc += t; }
    const u32 total = __shfl_sync(0xffffffffu, inc, 31);
    if (total == 0) return;
    if (wb.fill + total > WB_CAP) wb_flush(wb, es);
    u32 pos = wb.fill + inc - cnt;
    if (ab) wb.buf[pos++] = make_uint2(a, b);
    if (ba) wb.buf[pos] = make_uint2(b, a);
    wb.fill += total;
}

// ---- bulk asynchronous copy (TMA engine, non-tensor form) + mbarrier: the BULK variant of the evaluation kernel prefetches
// the NEXT block pair's column words (768 B at 12 nt, contiguous, 16-byte aligned) while the current pair is evaluated ----
__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, u32 bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, u32 phase) {
    asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}"
                 ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}

// ---- evaluation: one warp per block pair ----
// BULK = false: the warp copies the column block's words with two 16-byte loads per lane, then evaluates (the load latency is
// exposed once per pair unless another warp covers it).  BULK = true: two buffers per warp; the copy of pair i+1 is issued as
// ONE cp.async.bulk by lane 0 before pair i is evaluated and completes on an mbarrier.  A/B in profiles/ ("TMA").
template <int LP, int K, bool HASN, bool BULK>
__global__ void __launch_bounds__(256) hamming_blocks(const uint2 *__restrict__ pairs, u64 n_pairs, const u32 *__restrict__ blk_first,
                                                      const u32 *__restrict__ blk_cnt, const u32 *__restrict__ bsum,
                                                      const uint2 *__restrict__ planes, const u32 *__restrict__ nplane,
                                                      const u64 *__restrict__ ucode, const uint4 *__restrict__ eq, int L, int cull, EdgeSink es,
                                                      MiParams mi, const u32 *__restrict__ uidmap, unsigned long long *pairs_eval) {
    constexpr int XS = HASN ? 8 : 4;
    constexpr int NLET = HASN ? 5 : 4;
    const u32 lane = lane_id();
    const u32 lmask = L >= 32 ? 0xffffffffu : ((1u << L) - 1);
    const u64 nwarps = (u64)gridDim.x * (256 / 32);
    __shared__ uint2 s_edges[(256 / 32) * WB_CAP];
    constexpr int NBUF = BULK ? 2 : 1;
    constexpr u32 COL_BYTES = (u32)(LP * XS * 16);
    __shared__ __align__(128) uint4 s_eq[256 / 32][NBUF][LP * XS];
    __shared__ __align__(8) unsigned long long s_bar[256 / 32][2];
    const u32 wid = threadIdx.x >> 5;
    WarpEdgeBuf wb{s_edges + wid * WB_CAP, 0u};
    u64 evaluated = 0;
    u64 w = ((u64)blockIdx.x * 256 + threadIdx.x) >> 5;
    if (BULK) {
        if (lane == 0) { mbar_init(&s_bar[wid][0], 1); mbar_init(&s_bar[wid][1], 1); mbar_fence_init(); }
        __syncwarp();
        if (lane == 0 && w < n_pairs) {
            mbar_expect_tx(&s_bar[wid][0], COL_BYTES);
            bulk_g2s(s_eq[wid][0], eq + (u64)pairs[w].y * (LP * XS), COL_BYTES, &s_bar[wid][0]);
        }
    }
    for (u32 it = 0; w < n_pairs; w += nwarps, it++) {
        uint2 pr = pairs[w];
        const bool filt = pr.x >> 31;                 // multi-index pass: report a pair only in the pass of its first equal part
        pr.x &= 0x7fffffffu;
        const u32 rfirst = blk_first[pr.x], rcnt = blk_cnt[pr.x], cfirst = blk_first[pr.y], ccnt = blk_cnt[pr.y];
        const bool same = pr.x == pr.y;
        u32 cs[5] = {0, 0, 0, 0, 0};
        if (cull) {
#pragma unroll
            for (int x = 0; x < NLET; x++) cs[x] = bsum[(u64)pr.y * 8 + x];
        }
        const uint4 *sw;
        if (BULK) {
            const u32 cur = it & 1u;
            // the other buffer's readers finished at the end of the previous iteration (__syncwarp below): refill it now
            const u64 wn = w + nwarps;
            __syncwarp();
            if (lane == 0 && wn < n_pairs) {
                mbar_expect_tx(&s_bar[wid][cur ^ 1u], COL_BYTES);
                bulk_g2s(s_eq[wid][cur ^ 1u], eq + (u64)pairs[wn].y * (LP * XS), COL_BYTES, &s_bar[wid][cur ^ 1u]);
            }
            mbar_wait(&s_bar[wid][cur], (it >> 1) & 1u);
            sw = s_eq[wid][cur];
        } else {
            uint4 *swr = s_eq[wid][0];
            const uint4 *base = eq + (u64)pr.y * (LP * XS);
            __syncwarp();                                  // the previous pair's slices are done with the buffer
#pragma unroll
            for (int i = 0; i < (LP * XS + 31) / 32; i++) { const u32 t = i * 32 + lane; if (t < (u32)(LP * XS)) swr[t] = __ldg(base + t); }
            __syncwarp();
            sw = swr;
        }
        for (u32 s = 0; s * 32 < rcnt; s++) {
            const u32 r = s * 32 + lane;
            const bool valid = r < rcnt;
            const uint2 rp = valid ? planes[rfirst + r] : make_uint2(0u, 0u);
            const u32 rn = (HASN && valid) ? nplane[rfirst + r] : 0u;
            const u64 rc = valid ? ucode[rfirst + r] : 0ull;
            if (cull && !same) {
                // the 32 rows of a slice are consecutive sorted UMIs: few letters per position
                u32 oh[5], t = 0;
                onehot_planes(rp, rn, valid ? lmask : 0u, oh);
#pragma unroll
                for (int x = 0; x < NLET; x++) t |= __reduce_or_sync(0xffffffffu, oh[x]) & cs[x];
                if (__popc(~t & lmask) > K) continue;          // > K positions with disjoint letter sets
            }
            evaluated += (u64)__popc(__ballot_sync(0xffffffffu, valid)) * ccnt;
            uint4 h = make_uint4(0u, 0u, 0u, 0u);
            if (valid) {
                u32 off[LP];                                   // uint4 index of the row's letter slot at position j
                // the letter at position j straight from the interleaved sort code: one shift + one mask
                constexpr int BPB = HASN ? 3 : 2;
#pragma unroll
                for (int j = 0; j < LP; j++) off[j] = (u32)(j * XS) + ((u32)(rc >> (BPB * j)) & (HASN ? 7u : 3u));
                uint4 m1 = make_uint4(~0u, ~0u, ~0u, ~0u), m2 = m1, m3 = m1, m4 = m1;
#pragma unroll
                for (int j = 0; j < LP; j++) {
                    const uint4 wv = sw[off[j]];
                    if (K >= 3) { m4.x = (wv.x & m4.x) | (~wv.x & m3.x); m4.y = (wv.y & m4.y) | (~wv.y & m3.y);
                                  m4.z = (wv.z & m4.z) | (~wv.z & m3.z); m4.w = (wv.w & m4.w) | (~wv.w & m3.w); }
                    if (K >= 2) { m3.x = (wv.x & m3.x) | (~wv.x & m2.x); m3.y = (wv.y & m3.y) | (~wv.y & m2.y);
                                  m3.z = (wv.z & m3.z) | (~wv.z & m2.z); m3.w = (wv.w & m3.w) | (~wv.w & m2.w); }
                    m2.x = (wv.x & m2.x) | (~wv.x & m1.x); m2.y = (wv.y & m2.y) | (~wv.y & m1.y);
                    m2.z = (wv.z & m2.z) | (~wv.z & m1.z); m2.w = (wv.w & m2.w) | (~wv.w & m1.w);
                    m1.x &= wv.x; m1.y &= wv.y; m1.z &= wv.z; m1.w &= wv.w;
                }
                h = K == 1 ? m2 : (K == 2 ? m3 : m4);
            }
            // ---- hits (rare) ----
            // round 1 spent 34 % of this kernel's instructions here (ncu source view, profiles/r2m_*): every row of a diagonal
            // block "hits" itself and its lower triangle, each of the four 32-column groups ran its own warp-uniform loop, and
            // every round paid two five-step shuffle scans.  Now: on the diagonal the self / lower-triangle bits are masked out
            // before anything loops, one loop serves all four groups (rounds = the busiest lane's hit count), and an edge's
            // slot in the warp buffer comes from two ballots.
            if (same) {
                // keep columns c > r: group i keeps bits above (r - 32 i)
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int d = (int)r - 32 * i;                    // r = row index within the block (= column index of itself)
                    const u32 keep = d < 0 ? 0xffffffffu : (d >= 31 ? 0u : (0xffffffffu << (d + 1)));
                    if (i == 0) h.x &= keep; else if (i == 1) h.y &= keep; else if (i == 2) h.z &= keep; else h.w &= keep;
                }
            }
            if (__any_sync(0xffffffffu, (h.x | h.y | h.z | h.w) != 0u)) {
                const u32 a = rfirst + r;                               // index in this pass's order
                const u32 ao = (valid && uidmap) ? uidmap[a] : a;       // unique id (edges, freq, thr use the main order)
                i32 fa = 0, ta = 0;
                if (valid && (h.x | h.y | h.z | h.w)) { fa = es.freq[ao]; ta = es.thr[ao]; }
                while (__any_sync(0xffffffffu, (h.x | h.y | h.z | h.w) != 0u)) {
                    bool ok = (h.x | h.y | h.z | h.w) != 0u;
                    // next hit of this lane: lowest bit of the first non-empty group
                    const u32 g = h.x ? 0u : (h.y ? 1u : (h.z ? 2u : 3u));
                    const u32 word = h.x ? h.x : (h.y ? h.y : (h.z ? h.z : h.w));
                    const u32 bit = ok ? (u32)__ffs(word) - 1u : 0u;
                    const u32 cleared = word & (word - 1u);
                    if (g == 0u) h.x = cleared; else if (g == 1u) h.y = cleared; else if (g == 2u) h.z = cleared; else h.w = cleared;
                    const u32 c = g * 32u + bit, bb = cfirst + c;
                    ok = ok && c < ccnt;
                    if (ok && filt) ok = mi_accept(mi, rc, ucode[bb]);
                    bool ab = false, ba = false;
                    u32 bo = bb;
                    if (ok) { if (uidmap) bo = uidmap[bb]; ab = es.freq[bo] <= ta; ba = fa <= es.thr[bo]; }
                    const u32 mab = __ballot_sync(0xffffffffu, ab), mba = __ballot_sync(0xffffffffu, ba);
                    const u32 total = (u32)__popc(mab) + (u32)__popc(mba);
                    if (total) {
                        if (wb.fill + total > WB_CAP) wb_flush(wb, es);
                        const u32 lt = lanemask_lt();
                        if (ab) wb.buf[wb.fill + __popc(mab & lt)] = make_uint2(ao, bo);
                        if (ba) wb.buf[wb.fill + __popc(mab) + __popc(mba & lt)] = make_uint2(bo, ao);
                        wb.fill += total;
                    }
                }
            }
        }
    }
    wb_flush(wb, es);
    if (lane == 0 && evaluated) atomicAdd(pairs_eval, (unsigned long long)evaluated);
}

// (A row-grouped form — one warp per row block streaming its list of column blocks, rows set up once — was built and measured in
// round 2: same instruction count as this kernel (set-up was not where the instructions went: 34 % were hit handling), 114
// registers, 2-3x slower.  Numbers in profiles/r2k_ab_rows_vs_pairs_*.jsonl and r2m_ncu_full_hamming_rows_C2.txt; the code is in
// the history at the commit "Row-grouped neighbour evaluation (hamming_rows)".)

// UMIGPU_K5_BULK=0/1 selects the staging variant at run time (A/B); the default is the measured winner ("TMA").
static inline bool blk_use_bulk() {
    const char *e = getenv("UMIGPU_K5_BULK");
    return e ? atoi(e) != 0 : false;
}
template <int LP, int K, bool HASN>
static int blk_launch_one(cudaStream_t stream, int num_sms, const uint2 *pairs, u64 n_pairs, const u32 *blk_first, const u32 *blk_cnt,
                          const u32 *bsum, const uint2 *planes, const u32 *nplane, const u64 *ucode, const uint4 *eq, int L, int cull, EdgeSink es,
                          MiParams mi, const u32 *uidmap, unsigned long long *pairs_eval) {
    // the bulk variant keeps two column buffers per warp: instantiated where they stay small (no N plane)
    constexpr bool CAN_BULK = !HASN;
    const bool bulk = CAN_BULK && blk_use_bulk();
    auto kern = (bulk && CAN_BULK) ? hamming_blocks<LP, K, HASN, CAN_BULK> : hamming_blocks<LP, K, HASN, false>;
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, 0) != cudaSuccess || occ < 1) occ = 1;
    u32 grid = (u32)std::min<u64>((n_pairs + 7) / 8, (u64)num_sms * occ * 4);
    if (grid == 0) return 0;
    kern<<<grid, 256, 0, stream>>>(pairs, n_pairs, blk_first, blk_cnt, bsum, planes, nplane, ucode, eq, L, cull, es, mi, uidmap, pairs_eval);
    return cudaPeekAtLastError() == cudaSuccess ? 0 : -1;
}

static inline int blk_lp(int L) { return L <= 8 ? 8 : L <= 12 ? 12 : L <= 16 ? 16 : L <= 24 ? 24 : 32; }

template <int K, bool HASN>
static int blk_launch_k(cudaStream_t stream, int num_sms, const uint2 *pairs, u64 n_pairs, const u32 *blk_first, const u32 *blk_cnt,
                        const u32 *bsum, const uint2 *planes, const u32 *nplane, const u64 *ucode, const uint4 *eq, int L, int cull, EdgeSink es,
                        MiParams mi, const u32 *uidmap, unsigned long long *pairs_eval) {
#define BLK_ARGS stream, num_sms, pairs, n_pairs, blk_first, blk_cnt, bsum, planes, nplane, ucode, eq, L, cull, es, mi, uidmap, pairs_eval
    switch (blk_lp(L)) {
    case 8:  return blk_launch_one<8, K, HASN>(BLK_ARGS);
    case 12: return blk_launch_one<12, K, HASN>(BLK_ARGS);
    case 16: return blk_launch_one<16, K, HASN>(BLK_ARGS);
    case 24: return blk_launch_one<24, K, HASN>(BLK_ARGS);
    default: if (HASN) return 1; return blk_launch_one<32, K, false>(BLK_ARGS);
    }
#undef BLK_ARGS
}

// returns 0 = launched, 1 = configuration not covered (k outside 1..3), -1 = CUDA error
static int launch_neighbours_blocks(cudaStream_t stream, int num_sms, const uint2 *pairs, u64 n_pairs, const u32 *blk_first,
                                    const u32 *blk_cnt, const u32 *bsum, const uint2 *planes, const u32 *nplane, const u64 *ucode,
                                    const uint4 *eq, int L, int k, bool has_n, int cull, EdgeSink es, MiParams mi, const u32 *uidmap,
                                    unsigned long long *pairs_eval) {
#define BLK_ARGS stream, num_sms, pairs, n_pairs, blk_first, blk_cnt, bsum, planes, nplane, ucode, eq, L, cull, es, mi, uidmap, pairs_eval
    if (k < 1 || k > 3) return 1;
    if (!has_n) {
        if (k == 1) return blk_launch_k<1, false>(BLK_ARGS);
        if (k == 2) return blk_launch_k<2, false>(BLK_ARGS);
        return blk_launch_k<3, false>(BLK_ARGS);
    }
    if (k == 1) return blk_launch_k<1, true>(BLK_ARGS);
    if (k == 2) return blk_launch_k<2, true>(BLK_ARGS);
    return blk_launch_k<3, true>(BLK_ARGS);
#undef BLK_ARGS
}
