// umigpu.cu — libumigpu.so: C ABI (include/umigpu.h) and the host-side pipeline that strings the
// sm_100a kernels together.  No CPU fallback: every compute step below is a kernel launch.
#include "../../include/umigpu.h"

#include "pack.cuh"
#include <algorithm>
#include <atomic>
#include <chrono>
#include <memory>
#include <condition_variable>
#include <mutex>
#include <queue>
#include <thread>
#include <stdarg.h>
#include <unordered_map>

#include "bam.cuh"
#include "cluster.cuh"
#include "common.cuh"
#include "group.cuh"
#include "hamming.cuh"
#include "hamming_bs.cuh"
#include "hamming_blocks.cuh"
#include "misc.cuh"
#include "pack.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"
#include "seg_sort.cuh"

static thread_local std::string g_last_error;

struct Chunk { u64 start, n, first_index; };

struct umigpu_ctx {
    umigpu_config cfg;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t side = nullptr;          // second stream: work that is independent of the main neighbour pass overlaps it
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    std::string err;
    u64 launches = 0;
    int num_sms = NUM_SMS_B200;

    // accumulated reads (SoA in HBM)
    u64 n_reads = 0;
    std::vector<Chunk> chunks;
    int have_score = -1, have_weight = -1;
    DevBuf d_tid, d_pos, d_rev, d_umi2, d_nmask, d_score, d_weight, d_ascii, d_tlen, d_btlen;
    int have_tlen = -1;                   // paired mode (template length in the bucket key): -1 undecided, 0/1 fixed by the first push
    DevBuf d_sc;
    DevScalars *h_sc = nullptr;   // pinned
    DevScalars *h_sc_init = nullptr;   // pinned, constant: the image umigpu_reset uploads (no host synchronisation needed)

    DevBuf d_key[2][2], d_idx[2], d_hist, d_tiles;
    DevBuf d_useg, d_rep, d_planes, d_nplane, d_bhead, d_wsum, d_read_uid, d_freq, d_thr, d_repidx, d_label, d_prio;
    DevBuf d_bstart, d_itemoff, d_items, d_edges, d_keep, d_state, d_blocked, d_bitmap, d_kept, d_roots, d_tileoff, d_tsum, d_tilestate, d_bsum, d_blkoff, d_blkfirst, d_blkcnt, d_eq, d_pairs, d_comp, d_ucode, d_ubkt, d_brank, d_bigbid, d_bstartbig, d_biguid, d_miplanes, d_minplane, d_miucode, d_miuid, d_cedges, d_tidtab, d_lintab;

    // results
    bool ran = false;
    umigpu_counters ctr;
    u32 n_unique = 0, n_buckets = 0;
    bool used_direct = false;
    u32 n_blocks = 0;
    u64 direct_pairs = 0;
    u64 n_edges = 0;
    KeyLayout lay;
    u64 *h_kept = nullptr; size_t h_kept_cap = 0;       // pinned; what umigpu_result.kept_read_index points to
    u64 *h_roots = nullptr; size_t h_roots_cap = 0;     // pinned
    u64 *h_umirep = nullptr;                            // pinned, same capacity as h_roots
    DevBuf d_umirep;
    DevBuf d_chunks;
    u64 *h_chunks = nullptr; size_t h_chunks_cap = 0;    // pinned staging of the chunk table
    // BAM feed
    DevBuf d_bamraw, d_bamoff, d_btid, d_bpos, d_brev, d_bumi2, d_bnmask, d_bscore, d_bvalid, d_orig;
    bool use_orig = false;
    u64 n_unmapped = 0, n_records = 0;

    cudaEvent_t ev[UMIGPU_N_STAGES][2];
    bool ev_ok[UMIGPU_N_STAGES];

    // state handed from stage to stage of one run
    bool st_weighted = false, st_need_edges = false;
    int sorted_cur = 0;                   // which ping-pong buffer holds the sorted keys / indices
    DevBuf d_stamp, d_rowptr, d_front[2]; // frontier clustering
    DevBuf d_segblk, d_segnext, d_segflag, d_segbig, d_seghist, d_wbuf[2];   // segmented sort (coordinate-sorted input)
    bool used_seg_sort = false;
    // sharded run (several devices, one dataset): see "shard group" below
    u32 skip_bucket = 0xffffffffu;        // owner: the hot bucket is searched by every device of the group, not here
    bool run_ok = false;                  // the last run returned UMIGPU_OK (its last step is a read-back: the stream is drained)
    bool big_known = false;               // h_sc->n_big_all / m_big_all describe this run's buckets (set by stage_group)
    u32 band = 0, n_bands = 1;            // hot child: this device evaluates the row tiles ti with ti % n_bands == band
    struct Xchg *x = nullptr;
    umigpu_ctx *hot = nullptr;            // child context that runs this device's band of the hot bucket
    umigpu_ctx *helper = nullptr;         // work context of the later multi-index passes (second host thread, own stream and buffers)
    bool is_child = false;
};

static int fail(umigpu_ctx *ctx, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    g_last_error = buf;
    if (ctx) ctx->err = buf;
    return code;
}

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return fail(ctx, e_ == cudaErrorMemoryAllocation ? UMIGPU_ERR_NOMEM : UMIGPU_ERR_CUDA, \
                                           "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
    } while (0)
#define LAUNCH(kern, grid, block, ...)                                                             \
    do {                                                                                           \
        kern<<<(grid), (block), 0, ctx->stream>>>(__VA_ARGS__);                                    \
        ctx->launches++;                                                                           \
        CK(cudaGetLastError());                                                                    \
    } while (0)

// launch with dynamic shared memory above the 48 KB default (opt-in attribute set on every call: cheap)
#define LAUNCH_SMEM(kern, grid, block, smem, ...)                                                  \
    do {                                                                                           \
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem)));  \
        kern<<<(grid), (block), (smem), ctx->stream>>>(__VA_ARGS__);                               \
        ctx->launches++;                                                                           \
        CK(cudaGetLastError());                                                                    \
    } while (0)

static inline int bits_for(u64 range) { int b = 0; while (range) { b++; range >>= 1; } return b; }
static inline u32 grid_for(u64 n, u32 block) { return (u32)std::max<u64>(1, ceil_div_u64(n, block)); }

#define STAGE_BEGIN(s) do { CK(cudaEventRecord(ctx->ev[s][0], ctx->stream)); } while (0)
#define STAGE_END(s)   do { CK(cudaEventRecord(ctx->ev[s][1], ctx->stream)); ctx->ev_ok[s] = true; } while (0)

extern "C" const char *umigpu_version(void) { return "umigpu 0.1 (sm_100a)"; }
extern "C" const char *umigpu_last_error(const umigpu_ctx *ctx) { return ctx ? ctx->err.c_str() : g_last_error.c_str(); }

extern "C" int umigpu_device_init(int32_t device) {
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaFree(nullptr);
    if (e != cudaSuccess) return fail(nullptr, UMIGPU_ERR_CUDA, "umigpu_device_init(%d): %s; libumigpu has no CPU fallback", device, cudaGetErrorString(e));
    return UMIGPU_OK;
}

extern "C" int umigpu_create(const umigpu_config *cfg, umigpu_ctx **out) {
    umigpu_ctx *ctx = nullptr;
    if (!cfg || !out) return fail(nullptr, UMIGPU_ERR_ARG, "umigpu_create: null argument");
    *out = nullptr;
    if (cfg->umi_len < 1 || cfg->umi_len > 32) return fail(nullptr, UMIGPU_ERR_UNSUPPORTED, "umi_len %u outside 1..32", cfg->umi_len);
    if (cfg->k < 0) return fail(nullptr, UMIGPU_ERR_ARG, "k must be >= 0");
    if (cfg->algo < UMIGPU_ALGO_DIR || cfg->algo > UMIGPU_ALGO_CC)
        return fail(nullptr, UMIGPU_ERR_ARG, "Invalid algorithm %d", cfg->algo);       // main.rs:86-91 panics
    if (cfg->merge < UMIGPU_MERGE_ANY || cfg->merge > UMIGPU_MERGE_MAPQUAL)
        return fail(nullptr, UMIGPU_ERR_ARG, "Invalid merge %d", cfg->merge);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, UMIGPU_ERR_CUDA, "no CUDA device (%s); libumigpu has no CPU fallback", cudaGetErrorString(e));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, UMIGPU_ERR_ARG, "device %d out of range (%d devices)", cfg->device, ndev);
    ctx = new (std::nothrow) umigpu_ctx();
    if (!ctx) return fail(nullptr, UMIGPU_ERR_NOMEM, "out of host memory");
    ctx->cfg = *cfg;
    memset(&ctx->ctr, 0, sizeof ctx->ctr);
    for (int s = 0; s < UMIGPU_N_STAGES; s++) { ctx->ev_ok[s] = false; ctx->ev[s][0] = ctx->ev[s][1] = nullptr; }
#define CKC(call) do { cudaError_t e2 = (call); if (e2 != cudaSuccess) { int rc = fail(nullptr, UMIGPU_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e2)); delete ctx; return rc; } } while (0)
    CKC(cudaSetDevice(cfg->device));
    int sms = 0;
    CKC(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg->device));
    ctx->num_sms = sms > 0 ? sms : NUM_SMS_B200;
    if (cfg->stream) { ctx->stream = (cudaStream_t)cfg->stream; ctx->own_stream = false; }
    else { CKC(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)); ctx->own_stream = true; }
    CKC(cudaStreamCreateWithFlags(&ctx->side, cudaStreamNonBlocking));
    CKC(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming)); CKC(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
    for (int s = 0; s < UMIGPU_N_STAGES; s++) { CKC(cudaEventCreate(&ctx->ev[s][0])); CKC(cudaEventCreate(&ctx->ev[s][1])); }
    CKC(ctx->d_sc.reserve(sizeof(DevScalars)));
    CKC(cudaMallocHost((void **)&ctx->h_sc, sizeof(DevScalars)));
    CKC(cudaMallocHost((void **)&ctx->h_sc_init, sizeof(DevScalars)));
#undef CKC
    *out = ctx;
    return umigpu_reset(ctx);
}

static void xchg_release(umigpu_ctx *ctx);

extern "C" void umigpu_destroy(umigpu_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->cfg.device);
    cudaStreamSynchronize(ctx->stream);
    DevBuf *bufs[] = {&ctx->d_tid, &ctx->d_pos, &ctx->d_rev, &ctx->d_umi2, &ctx->d_nmask, &ctx->d_score, &ctx->d_weight, &ctx->d_ascii, &ctx->d_tlen, &ctx->d_btlen,
                      &ctx->d_sc, &ctx->d_key[0][0], &ctx->d_key[0][1], &ctx->d_key[1][0], &ctx->d_key[1][1], &ctx->d_idx[0], &ctx->d_idx[1],
                      &ctx->d_hist, &ctx->d_tiles, &ctx->d_useg, &ctx->d_rep, &ctx->d_planes, &ctx->d_nplane, &ctx->d_bhead, &ctx->d_wsum,
                      &ctx->d_read_uid, &ctx->d_freq, &ctx->d_thr, &ctx->d_repidx, &ctx->d_label, &ctx->d_prio, &ctx->d_bstart, &ctx->d_itemoff,
                      &ctx->d_items, &ctx->d_edges, &ctx->d_keep, &ctx->d_state, &ctx->d_blocked, &ctx->d_bitmap, &ctx->d_kept, &ctx->d_roots,
                      &ctx->d_tileoff, &ctx->d_tsum, &ctx->d_tilestate, &ctx->d_bsum, &ctx->d_blkoff, &ctx->d_blkfirst, &ctx->d_blkcnt, &ctx->d_eq, &ctx->d_pairs, &ctx->d_comp, &ctx->d_ucode, &ctx->d_ubkt, &ctx->d_brank, &ctx->d_bigbid, &ctx->d_bstartbig, &ctx->d_biguid,
                      &ctx->d_miplanes, &ctx->d_minplane, &ctx->d_miucode, &ctx->d_miuid, &ctx->d_cedges, &ctx->d_tidtab, &ctx->d_lintab};
    for (DevBuf *b : bufs) b->release();
    if (ctx->h_sc) cudaFreeHost(ctx->h_sc);
    if (ctx->h_sc_init) cudaFreeHost(ctx->h_sc_init);
    if (ctx->h_kept) cudaFreeHost(ctx->h_kept);
    if (ctx->h_roots) cudaFreeHost(ctx->h_roots);
    if (ctx->h_umirep) cudaFreeHost(ctx->h_umirep);
    if (ctx->h_chunks) cudaFreeHost(ctx->h_chunks);
    ctx->d_chunks.release(); ctx->d_umirep.release();
    ctx->d_stamp.release(); ctx->d_rowptr.release(); ctx->d_front[0].release(); ctx->d_front[1].release();
    ctx->d_segblk.release(); ctx->d_segnext.release(); ctx->d_segflag.release(); ctx->d_segbig.release(); ctx->d_seghist.release();
    ctx->d_wbuf[0].release(); ctx->d_wbuf[1].release();
    if (ctx->hot) { umigpu_destroy(ctx->hot); ctx->hot = nullptr; }
    if (ctx->helper) { umigpu_destroy(ctx->helper); ctx->helper = nullptr; }
    xchg_release(ctx);
    DevBuf *bb[] = {&ctx->d_bamraw, &ctx->d_bamoff, &ctx->d_btid, &ctx->d_bpos, &ctx->d_brev, &ctx->d_bumi2, &ctx->d_bnmask, &ctx->d_bscore, &ctx->d_bvalid, &ctx->d_orig};
    for (DevBuf *b : bb) b->release();
    for (int s = 0; s < UMIGPU_N_STAGES; s++) { if (ctx->ev[s][0]) cudaEventDestroy(ctx->ev[s][0]); if (ctx->ev[s][1]) cudaEventDestroy(ctx->ev[s][1]); }
    if (ctx->side) { cudaStreamSynchronize(ctx->side); cudaStreamDestroy(ctx->side); }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

__global__ void __launch_bounds__(256) iota_kernel(u64 n, u32 *out) {
    u64 i = (u64)blockIdx.x * 256 + threadIdx.x;
    if (i < n) out[i] = (u32)i;
}

static int init_scalars(umigpu_ctx *ctx) {
    DevScalars z; memset(&z, 0, sizeof z);
    z.tid_min = 0x7fffffff; z.tid_max = (i32)0x80000000;
    z.pos_min = 0x7fffffffffffffffLL; z.pos_max = (i64)0x8000000000000000LL;
    z.tlen_min = 0x7fffffffffffffffLL; z.tlen_max = (i64)0x8000000000000000LL;
    z.key_lo = 0x7fffffffffffffffLL; z.key_hi = (i64)0x8000000000000000LL;
    z.hot_bucket = 0xffffffffu;
    // h_sc doubles as the read-back buffer (every read-back ends with a synchronisation, so none is in flight here); the upload
    // comes from a second pinned image that never changes: stream order is all it needs
    *ctx->h_sc = z;
    *ctx->h_sc_init = z;
    CK(cudaMemcpyAsync(ctx->d_sc.p, ctx->h_sc_init, sizeof z, cudaMemcpyHostToDevice, ctx->stream));
    return UMIGPU_OK;
}

extern "C" int umigpu_reset(umigpu_ctx *ctx) {
    if (!ctx) return fail(nullptr, UMIGPU_ERR_ARG, "null context");
    CK(cudaSetDevice(ctx->cfg.device));
    // include/umigpu.h: the caller's pinned / device arrays of a push are free again when the next run, fetch or RESET returns.
    // A successful run ends synchronised; a batch that is abandoned after its pushes (no run), or whose run failed half way, is
    // the case with work still in flight.
    if ((ctx->n_reads > 0 || ctx->n_records > 0) && !(ctx->ran && ctx->run_ok)) CK(cudaStreamSynchronize(ctx->stream));
    ctx->n_reads = 0; ctx->chunks.clear(); ctx->have_score = ctx->have_weight = ctx->have_tlen = -1;
    ctx->ran = false; ctx->n_unique = ctx->n_buckets = 0; ctx->n_edges = 0;
    ctx->use_orig = false; ctx->n_unmapped = 0; ctx->n_records = 0;
    memset(&ctx->ctr, 0, sizeof ctx->ctr);
    for (int s = 0; s < UMIGPU_N_STAGES; s++) ctx->ev_ok[s] = false;
    return init_scalars(ctx);
}

static int read_scalars(umigpu_ctx *ctx) {
    CK(cudaMemcpyAsync(ctx->h_sc, ctx->d_sc.p, sizeof(DevScalars), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return UMIGPU_OK;
}

// ------------------------------------------------------------------------------------------------
// push
// ------------------------------------------------------------------------------------------------
static int push_common(umigpu_ctx *ctx, u64 n, const i32 *tid, const i64 *pos, const u8 *rev, const u8 *ascii,
                       const i32 *score, const i32 *weight, u64 first_index, cudaMemcpyKind kind, const i64 *tlen = nullptr) {
    if (!ctx) return fail(nullptr, UMIGPU_ERR_ARG, "null context");
    CK(cudaSetDevice(ctx->cfg.device));
    if (ctx->ran) return fail(ctx, UMIGPU_ERR_STATE, "push after run: call umigpu_reset first");
    if (n == 0) return UMIGPU_OK;
    if (!ascii) return fail(ctx, UMIGPU_ERR_ARG, "umi_ascii is null");
    if (ctx->n_reads + n > 0xfffffffeull) return fail(ctx, UMIGPU_ERR_UNSUPPORTED, "more than 2^32-2 reads in one batch");
    if (ctx->have_score < 0) { ctx->have_score = score != nullptr; ctx->have_weight = weight != nullptr; }
    if ((score != nullptr) != (ctx->have_score == 1) || (weight != nullptr) != (ctx->have_weight == 1))
        return fail(ctx, UMIGPU_ERR_ARG, "score/weight must be given for all chunks or for none");
    if (ctx->have_tlen < 0) ctx->have_tlen = tlen != nullptr;
    if ((tlen != nullptr) != (ctx->have_tlen == 1)) return fail(ctx, UMIGPU_ERR_ARG, "paired and unpaired pushes cannot be mixed in one batch");
    if (!ctx->chunks.empty()) {
        const Chunk &c = ctx->chunks.back();
        if (first_index < c.first_index + c.n) return fail(ctx, UMIGPU_ERR_ARG, "chunks must be pushed in ascending read-index order");
    }
    const u64 old = ctx->n_reads, tot = old + n;
    const int L = (int)ctx->cfg.umi_len;
    cudaStream_t s = ctx->stream;
    STAGE_BEGIN(UMIGPU_STAGE_PACK);
    // a first chunk that already lives in HBM is borrowed, not copied (the caller keeps it valid until fetch/reset);
    // a later chunk makes reserve_keep copy the borrowed part into owned storage
    const bool zc = kind == cudaMemcpyDeviceToDevice && old == 0;
    auto put = [&](DevBuf &b, const void *src, size_t esz) -> cudaError_t {
        if (zc && src) { b.borrow(src); return cudaSuccess; }
        cudaError_t e = b.reserve_keep(tot * esz, old * esz, s);
        if (e != cudaSuccess) return e;
        return src ? cudaMemcpyAsync((char *)b.p + old * esz, src, n * esz, kind, s) : cudaMemsetAsync((char *)b.p + old * esz, 0, n * esz, s);
    };
    CK(put(ctx->d_tid, tid, 4)); CK(put(ctx->d_pos, pos, 8)); CK(put(ctx->d_rev, rev, 1));
    if (score) CK(put(ctx->d_score, score, 4));
    if (weight) CK(put(ctx->d_weight, weight, 4));
    if (tlen) CK(put(ctx->d_tlen, tlen, 8));
    CK(ctx->d_umi2.reserve_keep(tot * 8, old * 8, s));
    CK(ctx->d_nmask.reserve_keep(tot * 4, old * 4, s));
    const u8 *d_ascii = ascii;
    if (kind == cudaMemcpyHostToDevice) {
        CK(ctx->d_ascii.reserve(n * L));
        CK(cudaMemcpyAsync(ctx->d_ascii.p, ascii, n * L, kind, s));
        d_ascii = ctx->d_ascii.as<u8>();
    }
    LAUNCH(umi_pack_kernel, (u32)std::min<u64>(grid_for(n, PACK_THREADS), (u64)ctx->num_sms * 32), PACK_THREADS, d_ascii, n, L, ctx->d_tid.as<i32>() + old,
           ctx->d_pos.as<i64>() + old, (const i64 *)(tlen ? ctx->d_tlen.as<i64>() + old : nullptr), ctx->d_umi2.as<u64>() + old,
           ctx->d_nmask.as<u32>() + old, ctx->d_sc.as<DevScalars>());
    STAGE_END(UMIGPU_STAGE_PACK);
    if (ctx->use_orig) {
        CK(ctx->d_orig.reserve_keep(tot * 4, old * 4, s));
        LAUNCH(iota_kernel, grid_for(n, 256), 256, n, ctx->d_orig.as<u32>() + old);
    }
    ctx->chunks.push_back({old, n, first_index});
    ctx->n_reads = tot;
    ctx->n_records += n;
    return UMIGPU_OK;
}

extern "C" int umigpu_push_reads(umigpu_ctx *ctx, uint64_t n, const int32_t *tid, const int64_t *unclipped_pos,
                                 const uint8_t *is_reverse, const uint8_t *umi_ascii, const int32_t *score,
                                 const int32_t *weight, uint64_t first_read_index) {
    if (n && (!tid || !unclipped_pos || !is_reverse)) return fail(ctx, UMIGPU_ERR_ARG, "tid/unclipped_pos/is_reverse are null");
    return push_common(ctx, n, tid, unclipped_pos, is_reverse, umi_ascii, score, weight, first_read_index, cudaMemcpyHostToDevice);
}
extern "C" int umigpu_push_reads_device(umigpu_ctx *ctx, uint64_t n, const int32_t *tid, const int64_t *unclipped_pos,
                                        const uint8_t *is_reverse, const uint8_t *umi_ascii, const int32_t *score,
                                        const int32_t *weight, uint64_t first_read_index) {
    if (n && (!tid || !unclipped_pos || !is_reverse)) return fail(ctx, UMIGPU_ERR_ARG, "tid/unclipped_pos/is_reverse are null");
    return push_common(ctx, n, tid, unclipped_pos, is_reverse, umi_ascii, score, weight, first_read_index, cudaMemcpyDeviceToDevice);
}

// Compact host format: 18 (14 when umi_len <= 16) bytes per read over PCIe instead of 29.
extern "C" int umigpu_push_reads_packed(umigpu_ctx *ctx, uint64_t n, const int32_t *tid, const int32_t *pos32, const uint8_t *is_reverse,
                                        const void *umi_2bit, const uint32_t *n_mask, const uint8_t *score8, uint64_t first_read_index) {
    if (!ctx) return fail(nullptr, UMIGPU_ERR_ARG, "null context");
    CK(cudaSetDevice(ctx->cfg.device));
    if (ctx->ran) return fail(ctx, UMIGPU_ERR_STATE, "push after run: call umigpu_reset first");
    if (n == 0) return UMIGPU_OK;
    if (!tid || !pos32 || !is_reverse || !umi_2bit) return fail(ctx, UMIGPU_ERR_ARG, "tid/pos32/is_reverse/umi_2bit are null");
    if (ctx->n_reads + n > 0xfffffffeull) return fail(ctx, UMIGPU_ERR_UNSUPPORTED, "more than 2^32-2 reads in one batch");
    if (ctx->have_score < 0) { ctx->have_score = score8 != nullptr; ctx->have_weight = 0; }
    if ((score8 != nullptr) != (ctx->have_score == 1) || ctx->have_weight == 1)
        return fail(ctx, UMIGPU_ERR_ARG, "score/weight must be given for all chunks or for none");
    if (ctx->have_tlen < 0) ctx->have_tlen = 0;
    if (ctx->have_tlen == 1) return fail(ctx, UMIGPU_ERR_ARG, "paired and unpaired pushes cannot be mixed in one batch");
    if (!ctx->chunks.empty()) {
        const Chunk &c = ctx->chunks.back();
        if (first_read_index < c.first_index + c.n) return fail(ctx, UMIGPU_ERR_ARG, "chunks must be pushed in ascending read-index order");
    }
    const u64 old = ctx->n_reads, tot = old + n;
    const int L = (int)ctx->cfg.umi_len;
    const size_t ub = L <= 16 ? 4 : 8;
    cudaStream_t s = ctx->stream;
    STAGE_BEGIN(UMIGPU_STAGE_PACK);
    CK(ctx->d_tid.reserve_keep(tot * 4, old * 4, s)); CK(ctx->d_pos.reserve_keep(tot * 8, old * 8, s)); CK(ctx->d_rev.reserve_keep(tot, old, s));
    CK(ctx->d_umi2.reserve_keep(tot * 8, old * 8, s)); CK(ctx->d_nmask.reserve_keep(tot * 4, old * 4, s));
    if (score8) CK(ctx->d_score.reserve_keep(tot * 4, old * 4, s));
    // staging area for the compact arrays: [pos32 | umi | n_mask | score8], each 16-byte aligned
    const size_t o_pos = 0, o_umi = (n * 4 + 15) & ~(size_t)15, o_nm = o_umi + ((n * ub + 15) & ~(size_t)15),
                 o_sc = o_nm + (n_mask ? ((n * 4 + 15) & ~(size_t)15) : 0), total = o_sc + (score8 ? n : 0);
    CK(ctx->d_ascii.reserve(total));
    char *st = (char *)ctx->d_ascii.p;
    CK(cudaMemcpyAsync(ctx->d_tid.as<i32>() + old, tid, n * 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->d_rev.as<u8>() + old, is_reverse, n, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(st + o_pos, pos32, n * 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(st + o_umi, umi_2bit, n * ub, cudaMemcpyHostToDevice, s));
    if (n_mask) CK(cudaMemcpyAsync(st + o_nm, n_mask, n * 4, cudaMemcpyHostToDevice, s));
    if (score8) CK(cudaMemcpyAsync(st + o_sc, score8, n, cudaMemcpyHostToDevice, s));
    DevScalars *sc = ctx->d_sc.as<DevScalars>();
    LAUNCH(unpack_compact_kernel, grid_for(n, 256), 256, n, L, (const i32 *)(st + o_pos), (const void *)(st + o_umi),
           n_mask ? (const u32 *)(st + o_nm) : (const u32 *)nullptr, score8 ? (const u8 *)(st + o_sc) : (const u8 *)nullptr,
           ctx->d_pos.as<i64>() + old, ctx->d_umi2.as<u64>() + old, ctx->d_nmask.as<u32>() + old,
           score8 ? ctx->d_score.as<i32>() + old : (i32 *)nullptr, sc);
    LAUNCH(range_reduce_kernel, grid_for(n, 256), 256, n, (const i32 *)(ctx->d_tid.as<i32>() + old), (const i64 *)(ctx->d_pos.as<i64>() + old),
           (const i64 *)nullptr, (const u32 *)(ctx->d_nmask.as<u32>() + old), sc);
    STAGE_END(UMIGPU_STAGE_PACK);
    if (ctx->use_orig) {
        CK(ctx->d_orig.reserve_keep(tot * 4, old * 4, s));
        LAUNCH(iota_kernel, grid_for(n, 256), 256, n, ctx->d_orig.as<u32>() + old);
    }
    ctx->chunks.push_back({old, n, first_read_index});
    ctx->n_reads = tot;
    ctx->n_records += n;
    return UMIGPU_OK;
}

extern "C" int umigpu_push_reads_paired(umigpu_ctx *ctx, uint64_t n, const int32_t *tid, const int64_t *unclipped_pos,
                                        const uint8_t *is_reverse, const int64_t *tlen, const uint8_t *umi_ascii,
                                        const int32_t *score, const int32_t *weight, uint64_t first_read_index) {
    if (n && (!tid || !unclipped_pos || !is_reverse || !tlen)) return fail(ctx, UMIGPU_ERR_ARG, "tid/unclipped_pos/is_reverse/tlen are null");
    return push_common(ctx, n, tid, unclipped_pos, is_reverse, umi_ascii, score, weight, first_read_index, cudaMemcpyHostToDevice, tlen);
}

// ------------------------------------------------------------------------------------------------
// device-wide scan driver
// ------------------------------------------------------------------------------------------------
template <class F, class G>
static int run_scan(umigpu_ctx *ctx, F f, G g, u64 n, u32 *total_dev /* may be null */) {
    u64 ntiles = ceil_div_u64(n, SCAN_TILE);
    CK(ctx->d_tiles.reserve(ntiles * sizeof(u32)));
    u32 *ts = ctx->d_tiles.as<u32>();
    LAUNCH((scan_tile_sums<u32, F>), (u32)ntiles, SCAN_THREADS, f, n, ts);
    LAUNCH((scan_spine<u32>), 1, 1024, ts, ntiles, total_dev);
    LAUNCH((scan_apply<u32, F, G>), (u32)ntiles, SCAN_THREADS, f, g, n, (const u32 *)ts);
    return UMIGPU_OK;
}

// ------------------------------------------------------------------------------------------------
// BAM feed: raw records -> SoA on the device (bam.cuh)
// ------------------------------------------------------------------------------------------------
extern "C" int umigpu_bam_record_offsets(const uint8_t *buf, uint64_t len, uint64_t *offsets, uint64_t max_records,
                                         uint64_t *n_records, uint64_t *consumed) {
    if (!buf || !offsets || !n_records || !consumed) return fail(nullptr, UMIGPU_ERR_ARG, "umigpu_bam_record_offsets: null argument");
    u64 off = 0, n = 0;
    while (n < max_records && off + 4 <= len) {
        u32 bs = (u32)buf[off] | ((u32)buf[off + 1] << 8) | ((u32)buf[off + 2] << 16) | ((u32)buf[off + 3] << 24);
        if (bs < 32) return fail(nullptr, UMIGPU_ERR_ARG, "umigpu_bam_record_offsets: record %llu has block_size %u", (unsigned long long)n, bs);
        if (off + 4 + bs > len) break;            // partial record: the caller refills and continues from *consumed
        offsets[n++] = off;
        off += 4 + (u64)bs;
    }
    offsets[n] = off;
    *n_records = n; *consumed = off;
    return UMIGPU_OK;
}

extern "C" int umigpu_push_bam_records(umigpu_ctx *ctx, uint64_t n, const uint8_t *records, const uint64_t *offsets,
                                       uint8_t umi_sep, uint64_t first_read_index, uint64_t *n_unmapped_out) {
    if (!ctx) return fail(nullptr, UMIGPU_ERR_ARG, "null context");
    CK(cudaSetDevice(ctx->cfg.device));
    if (ctx->ran) return fail(ctx, UMIGPU_ERR_STATE, "push after run: call umigpu_reset first");
    if (n_unmapped_out) *n_unmapped_out = 0;
    if (n == 0) return UMIGPU_OK;
    if (!records || !offsets) return fail(ctx, UMIGPU_ERR_ARG, "null argument");
    if (n > 0xfffffffeull || ctx->n_reads + n > 0xfffffffeull) return fail(ctx, UMIGPU_ERR_UNSUPPORTED, "more than 2^32-2 reads in one batch");
    if (ctx->have_score < 0) { ctx->have_score = 1; ctx->have_weight = 0; }
    if (ctx->have_score != 1 || ctx->have_weight == 1) return fail(ctx, UMIGPU_ERR_ARG, "BAM pushes cannot be mixed with score-less or weighted pushes");
    const bool paired = (ctx->cfg.flags & UMIGPU_FLAG_PAIRED) != 0;
    if (ctx->have_tlen < 0) ctx->have_tlen = paired;
    if (paired != (ctx->have_tlen == 1)) return fail(ctx, UMIGPU_ERR_ARG, "paired and unpaired pushes cannot be mixed in one batch");
    if (!ctx->chunks.empty()) {
        const Chunk &c = ctx->chunks.back();
        if (first_read_index < c.first_index + c.n) return fail(ctx, UMIGPU_ERR_ARG, "chunks must be pushed in ascending read-index order");
    }
    cudaStream_t s = ctx->stream;
    DevScalars *sc = ctx->d_sc.as<DevScalars>();
    const u64 old = ctx->n_reads, tot = old + n, base = offsets[0], bytes = offsets[n] - base;
    STAGE_BEGIN(UMIGPU_STAGE_PACK);
    if (!ctx->use_orig) {              // earlier ASCII chunks get the identity mapping
        CK(ctx->d_orig.reserve_keep(tot * 4, 0, s));
        for (const Chunk &c : ctx->chunks) LAUNCH(iota_kernel, grid_for(c.n, 256), 256, c.n, ctx->d_orig.as<u32>() + c.start);
        ctx->use_orig = true;
    } else CK(ctx->d_orig.reserve_keep(tot * 4, old * 4, s));
    CK(ctx->d_bamraw.reserve(bytes + 16)); CK(ctx->d_bamoff.reserve((n + 1) * 8));
    CK(ctx->d_btid.reserve(n * 4)); CK(ctx->d_bpos.reserve(n * 8)); CK(ctx->d_brev.reserve(n)); CK(ctx->d_bumi2.reserve(n * 8));
    CK(ctx->d_bnmask.reserve(n * 4)); CK(ctx->d_bscore.reserve(n * 4)); CK(ctx->d_bvalid.reserve(n));
    CK(ctx->d_tid.reserve_keep(tot * 4, old * 4, s)); CK(ctx->d_pos.reserve_keep(tot * 8, old * 8, s));
    CK(ctx->d_rev.reserve_keep(tot, old, s)); CK(ctx->d_umi2.reserve_keep(tot * 8, old * 8, s));
    CK(ctx->d_nmask.reserve_keep(tot * 4, old * 4, s)); CK(ctx->d_score.reserve_keep(tot * 4, old * 4, s));
    if (paired) { CK(ctx->d_btlen.reserve(n * 8)); CK(ctx->d_tlen.reserve_keep(tot * 8, old * 8, s)); }
    const u64 prev_unmapped = ctx->h_sc->n_bam_unmapped, prev_mates = ctx->h_sc->n_mates_skipped;
    CK(cudaMemcpyAsync(ctx->d_bamraw.p, records + base, bytes, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->d_bamoff.p, offsets, (n + 1) * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync(&sc->bam_err, 0, 8, s));
    BamDecodeOut tmp{ctx->d_btid.as<i32>(), ctx->d_bpos.as<i64>(), ctx->d_brev.as<u8>(), ctx->d_bumi2.as<u64>(), ctx->d_bnmask.as<u32>(),
                     ctx->d_bscore.as<i32>(), ctx->d_bvalid.as<u8>(), paired ? ctx->d_btlen.as<i64>() : nullptr};
    // offsets are relative to `records`; the device copy starts at records + base
    LAUNCH(bam_decode_kernel, grid_for(n, 128), 128, n, (const u8 *)ctx->d_bamraw.p - base, (const u64 *)ctx->d_bamoff.p, (int)ctx->cfg.umi_len,
           (u32)umi_sep, ctx->cfg.merge == UMIGPU_MERGE_MAPQUAL ? 1 : 0, paired ? 1 : 0, (ctx->cfg.flags & UMIGPU_FLAG_REMOVE_UNPAIRED) ? 1 : 0,
           (ctx->cfg.flags & UMIGPU_FLAG_REMOVE_CHIMERIC) ? 1 : 0, tmp, &sc->bam_err, sc);
    BamDecodeOut dst{ctx->d_tid.as<i32>() + old, ctx->d_pos.as<i64>() + old, ctx->d_rev.as<u8>() + old, ctx->d_umi2.as<u64>() + old,
                     ctx->d_nmask.as<u32>() + old, ctx->d_score.as<i32>() + old, nullptr, paired ? ctx->d_tlen.as<i64>() + old : nullptr};
    int rc = run_scan(ctx, BamValid{ctx->d_bvalid.as<u8>()}, BamCompact{tmp, dst, ctx->d_orig.as<u32>() + old, &sc->bam_valid, n}, n, nullptr);
    if (rc) return rc;
    rc = read_scalars(ctx);
    if (rc) return rc;
    const u32 e = ctx->h_sc->bam_err, m = ctx->h_sc->bam_valid;
    if (e & BAM_ERR_TRUNC) return fail(ctx, UMIGPU_ERR_ARG, "truncated or malformed BAM record");
    if (e & BAM_ERR_NO_SEP) return fail(ctx, UMIGPU_ERR_ARG, "failed to get the umi");                         // utils/read.rs:109
    if (e & BAM_ERR_SHORT) return fail(ctx, UMIGPU_ERR_ARG, "read name too short for a UMI of %u bases", ctx->cfg.umi_len);
    if (e & BAM_ERR_BAD_BASE) return fail(ctx, UMIGPU_ERR_BAD_BASE, "Unknown character in UMI sequence");      // utils/mod.rs:78
    if (m) LAUNCH(range_reduce_kernel, grid_for(m, 256), 256, (u64)m, (const i32 *)(ctx->d_tid.as<i32>() + old), (const i64 *)(ctx->d_pos.as<i64>() + old),
                  (const i64 *)(paired ? ctx->d_tlen.as<i64>() + old : nullptr), (const u32 *)(ctx->d_nmask.as<u32>() + old), sc);
    STAGE_END(UMIGPU_STAGE_PACK);
    if (m) ctx->chunks.push_back({old, m, first_read_index});
    ctx->n_reads = old + m;
    // deduplicate_sam.rs:96-100: skipped mates are not input reads; :104/:119 both count as unmapped
    const u64 unm = ctx->h_sc->n_bam_unmapped - prev_unmapped, mates = ctx->h_sc->n_mates_skipped - prev_mates;
    ctx->n_records += n - mates;
    ctx->n_unmapped += unm;
    if (n_unmapped_out) *n_unmapped_out = unm;
    return UMIGPU_OK;
}

// stable LSD radix sort of the NW-word keys in d_key[0] (+ index payload) over the passes of `plan`;
// returns the buffer index holding the result.  The first pass synthesises idx = position.
#define RS_ITEMS_1 12     // keys per thread, one-word keys  (tile 6144, 2 CTAs/SM)
#define RS_ITEMS_2 8      // keys per thread, two-word keys  (tile 4096)
static int run_sort(umigpu_ctx *ctx, u64 n, int nw, const SortPlan &plan, int *cur_out) {
    DevScalars *sc = ctx->d_sc.as<DevScalars>();
    static int items1 = 0;
    if (!items1) { const char *e = getenv("UMIGPU_RS_ITEMS"); items1 = e ? atoi(e) : RS_ITEMS_1; if (items1 != 8 && items1 != 12 && items1 != 16) items1 = RS_ITEMS_1; }
    const u32 tile = RS_THREADS * (nw == 1 ? items1 : RS_ITEMS_2);
    const u32 ntiles = (u32)ceil_div_u64(n, tile);
    CK(ctx->d_hist.reserve((size_t)RS_MAX_PASSES * RS_RADIX * sizeof(u32)));
    CK(ctx->d_tilestate.reserve((size_t)ntiles * RS_RADIX * 8));
    u32 *ghist = ctx->d_hist.as<u32>();
    CK(cudaMemsetAsync(ghist, 0, (size_t)plan.npass * RS_RADIX * 4, ctx->stream));
    CK(cudaMemsetAsync(&sc->sort_err, 0, 4, ctx->stream));
    KeyArr k0{{ctx->d_key[0][0].as<u64>(), nw == 2 ? ctx->d_key[0][1].as<u64>() : nullptr}};
    u32 ggrid = (u32)std::min<u64>(ceil_div_u64(n, (u64)GH_THREADS * GH_ITEMS), (u64)ctx->num_sms * 8);
    LAUNCH(radix_global_hist, ggrid, GH_THREADS, k0, n, plan, ghist);
    LAUNCH(radix_digit_starts, 1, RS_RADIX, ghist, plan.npass);
    int cur = 0;
    for (int pi = 0; pi < plan.npass; pi++) {
        const SortPass &p = plan.p[pi];
        KeyArr in{{ctx->d_key[cur][0].as<u64>(), ctx->d_key[cur][1].as<u64>()}};
        KeyArr out{{ctx->d_key[cur ^ 1][0].as<u64>(), ctx->d_key[cur ^ 1][1].as<u64>()}};
        const u32 mask = (1u << p.bits) - 1;
        CK(cudaMemsetAsync(ctx->d_tilestate.p, 0, (size_t)ntiles * RS_RADIX * 8, ctx->stream));
        CK(cudaMemsetAsync(&sc->sort_ticket, 0, 4, ctx->stream));
        const size_t dyn = (size_t)tile * (nw * 8 + 4);
#define RS_LAUNCH1(IT) LAUNCH_SMEM((radix_onesweep<1, IT>), ntiles, RS_THREADS, dyn, in, (const u32 *)ctx->d_idx[cur].as<u32>(), out, ctx->d_idx[cur ^ 1].as<u32>(), n, \
                   p.word, p.shift, mask, (const u32 *)(ghist + pi * RS_RADIX), ctx->d_tilestate.as<unsigned long long>(), &sc->sort_ticket, &sc->sort_err, pi == 0 ? 1 : 0)
        if (nw == 1) { if (items1 == 8) RS_LAUNCH1(8); else if (items1 == 16) RS_LAUNCH1(16); else RS_LAUNCH1(12); }
#undef RS_LAUNCH1
        else
            LAUNCH_SMEM((radix_onesweep<2, RS_ITEMS_2>), ntiles, RS_THREADS, dyn, in, (const u32 *)ctx->d_idx[cur].as<u32>(), out, ctx->d_idx[cur ^ 1].as<u32>(), n,
                   p.word, p.shift, mask, (const u32 *)(ghist + pi * RS_RADIX), ctx->d_tilestate.as<unsigned long long>(), &sc->sort_ticket, &sc->sort_err, pi == 0 ? 1 : 0);
        cur ^= 1;
    }
    *cur_out = cur;
    return UMIGPU_OK;
}

// Segmented sort for input that arrives ordered by (contig, position) — seg_sort.cuh.  *done = false: the input is not
// ordered that way (or the key does not fit the packed element): the caller runs the generic sort.  Result in buffer 1.
static bool seg_sort_applies(const KeyLayout &lay, u64 n) {
    const int sbits = 1 + lay.tlen_bits + lay.umi_bits, ib = std::max(1, bits_for(n - 1));
    return lay.nw == 1 && sbits <= 40 && sbits + ib <= 64 && !getenv("UMIGPU_NO_SEG_SORT");
}
// block table + plan scalars of the segmented sort (before the keys are built: the fused key kernel fills the table)
static int seg_sort_prepare(umigpu_ctx *ctx, u64 n) {
    DevScalars *sc = ctx->d_sc.as<DevScalars>();
    const u32 nblk = (u32)ceil_div_u64(n, SEG_BLK);
    CK(ctx->d_segblk.reserve((size_t)nblk * sizeof(SegBlock))); CK(ctx->d_segnext.reserve((size_t)nblk * 4)); CK(ctx->d_segflag.reserve(nblk));
    CK(ctx->d_segbig.reserve(((size_t)nblk + 1) * sizeof(SegBig)));
    CK(cudaMemsetAsync(&sc->seg_n_big, 0, 16, ctx->stream));
    return UMIGPU_OK;
}
// have_table: the block table was filled while the keys were built (build_keys_summary_kernel)
static int run_sort_presorted(umigpu_ctx *ctx, u64 n, const KeyLayout &lay, bool have_table, bool *done) {
    *done = false;
    DevScalars *sc = ctx->d_sc.as<DevScalars>();
    const int sbits = 1 + lay.tlen_bits + lay.umi_bits, ib = std::max(1, bits_for(n - 1));
    if (!seg_sort_applies(lay, n)) return UMIGPU_OK;
    const u32 nblk = (u32)ceil_div_u64(n, SEG_BLK);
    SegPlanOut *plan = reinterpret_cast<SegPlanOut *>(&sc->seg_n_big);
    const u64 *key_in = ctx->d_key[0][0].as<u64>();
    u64 *key_out = ctx->d_key[1][0].as<u64>();
    u32 *idx_out = ctx->d_idx[1].as<u32>();
    if (!have_table) {
        int r0 = seg_sort_prepare(ctx, n);
        if (r0) return r0;
        LAUNCH(seg_block_summary_kernel, nblk, 256, key_in, n, sbits, ctx->d_segblk.as<SegBlock>(), plan);
    }
    LAUNCH(seg_plan_kernel, 1, 1024, (const SegBlock *)ctx->d_segblk.p, nblk, n, (u32)SEG_TILE, ctx->d_segnext.as<u32>(), ctx->d_segflag.as<u8>(),
           ctx->d_segbig.as<SegBig>(), plan);
    int rc = read_scalars(ctx);
    if (rc) return rc;
    if (ctx->h_sc->seg_unsorted) return UMIGPU_OK;
    const u32 n_big = ctx->h_sc->seg_n_big, n_tiles = ctx->h_sc->seg_n_tiles;
    // small segments: one window per block, on the side stream while the big segments' passes run on the main one
    CK(cudaEventRecord(ctx->ev_fork, ctx->stream));
    CK(cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0));
    {
        cudaStream_t main_stream = ctx->stream;
        ctx->stream = n_big ? ctx->side : main_stream;
        int r2 = [&]() -> int {
            LAUNCH_SMEM(seg_window_sort_kernel, nblk, SEG_THREADS, (size_t)SEG_WIN * 8, key_in, n, sbits, (const SegBlock *)ctx->d_segblk.p,
                        (const u32 *)ctx->d_segnext.p, (const u8 *)ctx->d_segflag.p, key_out, idx_out);
            return UMIGPU_OK;
        }();
        cudaError_t ej = cudaEventRecord(ctx->ev_join, ctx->stream);
        ctx->stream = main_stream;
        if (r2) return r2;
        CK(ej);
    }
    if (n_big) {
        const SegPasses sp = seg_passes(sbits);
        const int npass = sp.npass;
        if (npass > SEG_MAX_PASSES) return fail(ctx, UMIGPU_ERR_STATE, "internal: segmented sort with %d passes", npass);
        const size_t hist_bytes = (size_t)n_big * npass * SEG_RADIX * 4;
        CK(ctx->d_seghist.reserve(hist_bytes));
        CK(ctx->d_tilestate.reserve((size_t)n_tiles * SEG_RADIX * 8));
        if (npass > 1) { CK(ctx->d_wbuf[0].reserve(n * 8)); if (npass > 2) CK(ctx->d_wbuf[1].reserve(n * 8)); }
        CK(cudaMemsetAsync(ctx->d_seghist.p, 0, hist_bytes, ctx->stream));
        CK(cudaMemsetAsync(&sc->sort_err, 0, 4, ctx->stream));
        const SegBig *big = (const SegBig *)ctx->d_segbig.p;
        LAUNCH(seg_hist_kernel, std::min<u32>(n_tiles, (u32)ctx->num_sms * 4), SEG_THREADS, key_in, sbits, sp, big, n_big, n_tiles, ctx->d_seghist.as<u32>());
        LAUNCH(seg_digit_starts_kernel, std::min<u32>(n_big * (u32)npass, (u32)ctx->num_sms * 8), SEG_RADIX, ctx->d_seghist.as<u32>(), n_big * (u32)npass);
        int wcur = 0;
        for (int p = 0; p < npass; p++) {
            CK(cudaMemsetAsync(ctx->d_tilestate.p, 0, (size_t)n_tiles * SEG_RADIX * 8, ctx->stream));
            CK(cudaMemsetAsync(&sc->sort_ticket, 0, 4, ctx->stream));
            const u64 *w_in = p == 0 ? nullptr : ctx->d_wbuf[wcur].as<u64>();
            u64 *w_out = p == npass - 1 ? nullptr : ctx->d_wbuf[p == 0 ? 0 : wcur ^ 1].as<u64>();
            SegPassArgs pa{key_in, w_in, w_out, key_out, idx_out, sbits, ib, p, sp, big, n_big, n_tiles, (const u32 *)ctx->d_seghist.p,
                           ctx->d_tilestate.as<unsigned long long>(), &sc->sort_ticket, &sc->sort_err};
            if (sp.rb == 9) LAUNCH_SMEM((seg_onesweep<SEG_ITEMS, 9>), n_tiles, SEG_THREADS, (size_t)SEG_TILE * 8, pa);
            else            LAUNCH_SMEM((seg_onesweep<SEG_ITEMS, 8>), n_tiles, SEG_THREADS, (size_t)SEG_TILE * 8, pa);
            if (p > 0) wcur ^= 1;
        }
        CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
    }
    *done = true;
    return UMIGPU_OK;
}

// One ordering of the unique UMIs of a set of buckets, as seen by the neighbour search.
struct NView {
    const uint2 *planes; const u32 *nplane; const u64 *ucode;
    const u32 *uidmap;          // index in this ordering -> unique id (nullptr = identity: the main order)
    const u32 *bstart; u32 n_buckets;
};

// Work list (tile pairs -> block pairs) and evaluation for one ordering.  *dense is set (and nothing is evaluated)
// when a multi-index pass finds that culling does not thin the work out: the caller then restarts without it.
static int neighbour_pass(umigpu_ctx *ctx, const NView &v, const MiParams &mi, EdgeSink es, bool has_n, int cull, bool allow_blocks, bool *dense, u32 skip) {
    const umigpu_config &cfg = ctx->cfg;
    DevScalars *sc = ctx->d_sc.as<DevScalars>();
    const int k = cfg.k, L = (int)cfg.umi_len;
    const u32 B = v.n_buckets;
    *dense = false;
    if (B == 0) return UMIGPU_OK;
    CK(ctx->d_itemoff.reserve(((size_t)B + 1) * 4));
    CK(ctx->d_tileoff.reserve(((size_t)B + 1) * 4));
    CK(ctx->d_blkoff.reserve(((size_t)B + 1) * 4));
    int rc = run_scan(ctx, BucketItems{v.bstart, skip}, BucketItemsEmit{ctx->d_itemoff.as<u32>(), B}, B, &sc->n_cand);
    if (rc) return rc;
    rc = run_scan(ctx, BucketTiles{v.bstart, skip}, BucketTilesEmit{ctx->d_tileoff.as<u32>(), B}, B, &sc->n_tiles);
    if (rc) return rc;
    rc = run_scan(ctx, BucketBlocks{v.bstart, skip}, BucketTilesEmit{ctx->d_blkoff.as<u32>(), B}, B, &sc->n_blocks);
    if (rc) return rc;
    rc = read_scalars(ctx);
    if (rc) return rc;
    const u32 n_cand = ctx->h_sc->n_cand, n_tiles = ctx->h_sc->n_tiles, n_blocks = ctx->h_sc->n_blocks;
    // the candidate tile pairs are counted in 32 bits: sum_b T_b (T_b + 1) / 2 <= (T_max + 1) * n_tiles / 2 + n_tiles
    {
        const u64 t_max = (ctx->h_sc->max_umis + HT_ROWS - 1) / HT_ROWS;
        if ((t_max + 1) * (u64)n_tiles / 2 + n_tiles >= 0x7fffffffull)
            return fail(ctx, UMIGPU_ERR_UNSUPPORTED, "a bucket of %u unique UMIs needs more than 2^31 tile pairs: split the batch", ctx->h_sc->max_umis);
    }
    if (n_cand == 0) return UMIGPU_OK;
    CK(ctx->d_tsum.reserve((size_t)std::max<u32>(n_tiles, 1) * TS_WORDS * 4));
    CK(ctx->d_bsum.reserve((size_t)std::max<u32>(n_blocks, 1) * TS_WORDS * 4));
    CK(ctx->d_blkfirst.reserve((size_t)std::max<u32>(n_blocks, 1) * 4));
    CK(ctx->d_blkcnt.reserve((size_t)std::max<u32>(n_blocks, 1) * 4));
    LAUNCH(tile_summary_kernel, grid_for((u64)n_tiles * 32, 256), 256, n_tiles, B, (const u32 *)ctx->d_tileoff.p, v.bstart, v.planes, v.nplane, L,
           ctx->d_tsum.as<u32>(), (const u32 *)ctx->d_blkoff.p, ctx->d_bsum.as<u32>(), ctx->d_blkfirst.as<u32>(), ctx->d_blkcnt.as<u32>(),
           v.ucode, mi);
    CK(ctx->d_items.reserve((size_t)n_cand * sizeof(TileItem)));
    CK(cudaMemsetAsync(&sc->n_items, 0, 4, ctx->stream));
    CK(cudaMemsetAsync(&sc->scratch2, 0, 8, ctx->stream));
    LAUNCH(build_items_kernel, grid_for(n_cand, 256), 256, n_cand, B, (const u32 *)ctx->d_itemoff.p, v.bstart, (const u32 *)ctx->d_tileoff.p,
           (const u32 *)ctx->d_blkoff.p, (const u32 *)ctx->d_tsum.p, L, k, cull, mi, ctx->d_items.as<TileItem>(), sc, ctx->band, ctx->n_bands);
    const TileItem *items = ctx->d_items.as<TileItem>();
    ctx->ctr.n_tile_candidates += n_cand;
    // The surviving tile pairs (W of them, counted on the device) are expanded straight away — the expansion reads W from the
    // device — and ONE read-back returns W, the scheduled pair space and the number of block pairs.
    u8 *need = nullptr;
    u64 pcap = 0;
    const bool twice = getenv("UMIGPU_EXPAND_TWICE") != nullptr;       // count first, then fill: the former schedule
    const u32 xgrid = (u32)std::min<u64>(grid_for((u64)n_cand * 32, 256), (u64)ctx->num_sms * 16);
    if (allow_blocks) {
        // ---- sparse form: global one-hot words, block-pair list, one warp per block pair ----
        // The list is filled speculatively into the room there is (the count runs on past it): one expansion when it fits —
        // every run after the first on similar input — and a second one after growing when it does not.
        // A band of a split hot bucket only touches the column blocks near its own row tiles: the expansion flags the column
        // blocks of its pairs and the one-hot words are built for those alone.
        if (ctx->n_bands > 1) {
            CK(ctx->d_blocked.reserve(std::max<u32>(n_blocks, 1)));
            need = ctx->d_blocked.as<u8>();
            CK(cudaMemsetAsync(need, 0, n_blocks, ctx->stream));
        }
        if (!twice) CK(ctx->d_pairs.reserve(std::min<u64>(std::max<u64>((u64)n_cand * 4, (u64)1 << 16), (u64)4 << 20) * sizeof(uint2)));   // first guess: at most 32 MB
        pcap = twice ? 0 : ctx->d_pairs.cap / sizeof(uint2);
        CK(cudaMemsetAsync(&sc->n_block_pairs, 0, 8, ctx->stream));
        LAUNCH(expand_blocks_kernel, xgrid, 256, items, 0u, (const u32 *)ctx->d_bsum.p, L, k, cull, mi, twice ? 0 : 1,
               twice ? (uint2 *)nullptr : ctx->d_pairs.as<uint2>(), (unsigned long long *)&sc->n_block_pairs, twice ? (u8 *)nullptr : need,
               (unsigned long long)pcap, (const u32 *)&sc->n_items);
    }
    rc = read_scalars(ctx);
    if (rc) return rc;
    const u32 W = ctx->h_sc->n_items;
    const u64 item_pairs = ctx->h_sc->scratch2;
    if (W == 0) return UMIGPU_OK;

    if (allow_blocks) {
        const int LP = blk_lp(L), XS = has_n ? 8 : 4;
        const u64 n_pairs = ctx->h_sc->n_block_pairs;
        // more than ~30 % of the scheduled pair space survives block culling: dense work runs better as shared-memory tiles
        // (a band of a split hot bucket never leaves the multi-index scheme: which pass reports a pair must not depend on
        // the rank, and the density a rank sees in its band does)
        const bool is_dense = (double)n_pairs * 16384.0 > 0.30 * (double)item_pairs && !(ctx->n_bands > 1 && mi.part >= 0);
        if (is_dense && mi.part >= 0) { *dense = true; return UMIGPU_OK; }
        if (!is_dense) {
            CK(ctx->d_eq.reserve((size_t)std::max<u32>(n_blocks, 1) * LP * XS * 16));
            if (n_pairs > pcap) {
                CK(ctx->d_pairs.reserve(std::max<u64>(n_pairs, 1) * sizeof(uint2)));
                pcap = ctx->d_pairs.cap / sizeof(uint2);
                CK(cudaMemsetAsync(&sc->n_block_pairs, 0, 8, ctx->stream));
                LAUNCH(expand_blocks_kernel, xgrid, 256, items, W, (const u32 *)ctx->d_bsum.p, L, k, cull, mi, 1, ctx->d_pairs.as<uint2>(),
                       (unsigned long long *)&sc->n_block_pairs, need, (unsigned long long)pcap);
            }
            if (has_n) LAUNCH(onehot_build_kernel<true>, grid_for((u64)n_blocks * 32, 256), 256, n_blocks * 4, (const u32 *)ctx->d_blkfirst.p,
                              (const u32 *)ctx->d_blkcnt.p, v.planes, v.nplane, L, LP, ctx->d_eq.as<u32>(), (const u8 *)need);
            else       LAUNCH(onehot_build_kernel<false>, grid_for((u64)n_blocks * 32, 256), 256, n_blocks * 4, (const u32 *)ctx->d_blkfirst.p,
                              (const u32 *)ctx->d_blkcnt.p, v.planes, v.nplane, L, LP, ctx->d_eq.as<u32>(), (const u8 *)need);
            rc = launch_neighbours_blocks(ctx->stream, ctx->num_sms, ctx->d_pairs.as<uint2>(), n_pairs, (const u32 *)ctx->d_blkfirst.p,
                                          (const u32 *)ctx->d_blkcnt.p, (const u32 *)ctx->d_bsum.p, v.planes, v.nplane, v.ucode, ctx->d_eq.as<uint4>(),
                                          L, k, has_n, cull, es, mi, v.uidmap, (unsigned long long *)&sc->pairs_eval);
            if (rc != 0) return fail(ctx, UMIGPU_ERR_CUDA, "block-pair neighbour kernel: %s", cudaGetErrorString(cudaGetLastError()));
            ctx->launches += 1;
            CK(cudaGetLastError());
            ctx->ctr.n_block_pairs += n_pairs; ctx->ctr.n_tile_items += W;
            return UMIGPU_OK;
        }
    }
    // ---- dense forms (main order only, no multi-index filter) ----
    if (mi.part >= 0 || v.uidmap) return fail(ctx, UMIGPU_ERR_STATE, "internal: dense neighbour kernel reached inside a multi-index pass");
    ctx->ctr.n_tile_items += W;
    if (!(cfg.flags & UMIGPU_FLAG_KERNEL_DIRECT)) {
        rc = launch_neighbours_bitsliced(ctx->stream, ctx->num_sms, items, W, v.planes, v.nplane, ctx->d_bsum.as<u32>(), L, k, has_n, cull, es,
                                         (u32 *)&sc->scratch, (unsigned long long *)&sc->pairs_eval);
        if (rc == 0) { ctx->launches += 1; CK(cudaGetLastError()); return UMIGPU_OK; }
        if (rc < 0) return fail(ctx, UMIGPU_ERR_CUDA, "bit-sliced neighbour kernel: %s", cudaGetErrorString(cudaGetLastError()));
        // rc > 0: configuration not covered by the bit-sliced kernels (k > 3) -> direct kernel
    }
    ctx->used_direct = true;
    ctx->direct_pairs += item_pairs;
    const u32 grid = std::min<u32>(W, (u32)ctx->num_sms * 4);
#define HD(KK, NN) LAUNCH((hamming_tiles_direct<KK, NN>), grid, HT_THREADS, items, W, v.planes, v.nplane, es, k)
    if (!has_n) { if (k == 1) HD(1, false); else if (k == 2) HD(2, false); else HD(0, false); }
    else        { if (k == 1) HD(1, true);  else if (k == 2) HD(2, true);  else HD(0, true); }
#undef HD
    return UMIGPU_OK;
}

// ------------------------------------------------------------------------------------------------
// run
// ------------------------------------------------------------------------------------------------
enum RunMode { RUN_FULL = 0, RUN_EDGES_ONLY = 1 };

// ---- stage 1: key layout, K1b keys, K2 sort, K3 unique / count / merge, bucket segmentation ----
static int stage_group(umigpu_ctx *ctx, bool want_labels, bool force_inf_thr) {
    const u64 n = ctx->n_reads;
    const umigpu_config &cfg = ctx->cfg;
    DevScalars *sc = ctx->d_sc.as<DevScalars>();

    // ---- key layout from the batch's ranges ----
    int rc = read_scalars(ctx);
    if (rc) return rc;
    if (ctx->h_sc->bad_base)
        return fail(ctx, UMIGPU_ERR_BAD_BASE, "Unknown character in UMI sequence");            // utils/mod.rs:78
    KeyLayout lay;
    lay.umi_len = (int)cfg.umi_len;
    lay.has_n = ctx->h_sc->any_n ? 1 : 0;
    lay.umi_bits = lay.has_n ? 3 * lay.umi_len : 2 * lay.umi_len;
    lay.tid_min = ctx->h_sc->tid_min; lay.pos_min = ctx->h_sc->pos_min;
    u64 tid_range = (u64)((i64)ctx->h_sc->tid_max - (i64)ctx->h_sc->tid_min);
    u64 pos_range = (u64)ctx->h_sc->pos_max - (u64)ctx->h_sc->pos_min;
    lay.tid_bits = bits_for(tid_range); lay.pos_bits = bits_for(pos_range);
    lay.tlen_bits = 0; lay.tlen_min = 0;
    if (ctx->have_tlen == 1) {                       // PairedAlignment: (strand, coord, ref, tlen), deduplicate_sam.rs:545-565
        lay.tlen_min = ctx->h_sc->tlen_min;
        lay.tlen_bits = bits_for((u64)ctx->h_sc->tlen_max - (u64)ctx->h_sc->tlen_min);
    }
    lay.bucket_bits = lay.tid_bits + lay.pos_bits + 1 + lay.tlen_bits;
    lay.total_bits = lay.bucket_bits + lay.umi_bits;
    lay.lin_off = nullptr; lay.lin_pmin = nullptr;
    if (lay.total_bits > 64 && tid_range >= 1 && tid_range < 65536 && !getenv("UMIGPU_NO_LINEAR_KEYS")) {
        // Many contigs (a genome with its alt/decoy contigs has thousands): [tid | pos] wastes bits on positions no contig
        // reaches.  Lay the contigs' OCCUPIED position ranges end to end instead — a human genome fits 32 bits — so that the
        // key stays one 64-bit word.  Only tried when the plain layout would need two words.
        const u32 T = (u32)tid_range + 1;
        CK(ctx->d_tidtab.reserve((size_t)T * 16));
        unsigned long long *vmin = ctx->d_tidtab.as<unsigned long long>(), *vmax = vmin + T;
        CK(cudaMemsetAsync(vmin, 0xff, (size_t)T * 8, ctx->stream)); CK(cudaMemsetAsync(vmax, 0, (size_t)T * 8, ctx->stream));
        LAUNCH(tid_range_kernel, (u32)std::min<u64>(grid_for(n, 256), (u64)ctx->num_sms * 16), 256, n, (const i32 *)ctx->d_tid.p, (const i64 *)ctx->d_pos.p,
               lay.tid_min, vmin, vmax);
        std::vector<unsigned long long> h((size_t)T * 2);
        CK(cudaMemcpyAsync(h.data(), vmin, (size_t)T * 16, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        std::vector<unsigned long long> tab((size_t)T * 2);          // [0,T) offsets, [T,2T) per-contig minimum position
        unsigned long long run = 0; bool ok = true;
        for (u32 t = 0; t < T && ok; t++) {
            tab[t] = run; tab[T + t] = 0;
            if (h[T + t] >= h[t]) {
                const unsigned long long span = h[T + t] - h[t];     // biased values: the difference is the true span - 1
                tab[T + t] = (unsigned long long)(i64)(h[t] ^ POS_BIAS);
                if (span >= (1ull << 62) || run + span + 1 >= (1ull << 62)) ok = false; else run += span + 1;
            }
        }
        const int lin_bits = ok ? bits_for(run ? run - 1 : 0) : 999;
        if (lin_bits < lay.tid_bits + lay.pos_bits) {
            CK(ctx->d_lintab.reserve((size_t)T * 16));
            CK(cudaMemcpyAsync(ctx->d_lintab.p, tab.data(), (size_t)T * 16, cudaMemcpyHostToDevice, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));                   // tab is a stack-lifetime pageable buffer
            lay.lin_off = ctx->d_lintab.as<u64>(); lay.lin_pmin = (const i64 *)(ctx->d_lintab.as<u64>() + T);
            lay.tid_bits = 0; lay.pos_bits = lin_bits;
            lay.bucket_bits = lay.pos_bits + 1 + lay.tlen_bits;
            lay.total_bits = lay.bucket_bits + lay.umi_bits;
        }
    }
    if (lay.bucket_bits > 64) return fail(ctx, UMIGPU_ERR_UNSUPPORTED, "bucket key needs %d bits (> 64)", lay.bucket_bits);
    if (lay.total_bits > 128) return fail(ctx, UMIGPU_ERR_UNSUPPORTED, "sort key needs %d bits (> 128): %d-nt UMIs with N leave %d bits for the bucket", lay.total_bits,
                                          lay.umi_len, 128 - lay.umi_bits);
    lay.nw = lay.total_bits <= 64 ? 1 : 2;
    ctx->ctr.key_bits = (u64)lay.total_bits;
    ctx->lay = lay;
    const bool has_n = lay.has_n != 0;

    // ---- K1b keys ----
    STAGE_BEGIN(UMIGPU_STAGE_KEYS);
    for (int b = 0; b < 2; b++) {
        CK(ctx->d_key[b][0].reserve(n * 8));
        if (lay.nw == 2) CK(ctx->d_key[b][1].reserve(n * 8));
        CK(ctx->d_idx[b].reserve(n * 4));
    }
#define BUILD_KEYS(NWV, LINV, K1) LAUNCH((build_keys_kernel<NWV, LINV>), grid_for(n, 256), 256, n, (const i32 *)ctx->d_tid.p, (const i64 *)ctx->d_pos.p, \
        (const u8 *)ctx->d_rev.p, (const i64 *)(lay.tlen_bits ? ctx->d_tlen.p : nullptr), (const u64 *)ctx->d_umi2.p, (const u32 *)ctx->d_nmask.p, lay, \
        ctx->d_key[0][0].as<u64>(), K1)
    // input that may be coordinate-sorted: the keys and the segmented sort's block table come out of one kernel
    const bool fused_table = seg_sort_applies(lay, n) && !getenv("UMIGPU_NO_FUSED_SUMMARY");
    if (fused_table) {
        rc = seg_sort_prepare(ctx, n);
        if (rc) return rc;
        const int sbits = 1 + lay.tlen_bits + lay.umi_bits;
        SegPlanOut *plan = reinterpret_cast<SegPlanOut *>(&sc->seg_n_big);
#define BUILD_KEYS_SEG(LINV) LAUNCH((build_keys_summary_kernel<LINV>), (u32)ceil_div_u64(n, SEG_BLK), 256, n, (const i32 *)ctx->d_tid.p, (const i64 *)ctx->d_pos.p, \
        (const u8 *)ctx->d_rev.p, (const i64 *)(lay.tlen_bits ? ctx->d_tlen.p : nullptr), (const u64 *)ctx->d_umi2.p, (const u32 *)ctx->d_nmask.p, lay, \
        ctx->d_key[0][0].as<u64>(), sbits, ctx->d_segblk.as<SegBlock>(), plan)
        if (lay.lin_off) BUILD_KEYS_SEG(true); else BUILD_KEYS_SEG(false);
#undef BUILD_KEYS_SEG
    }
    else if (lay.nw == 1) { if (lay.lin_off) BUILD_KEYS(1, true, (u64 *)nullptr); else BUILD_KEYS(1, false, (u64 *)nullptr); }
    else { if (lay.lin_off) BUILD_KEYS(2, true, ctx->d_key[0][1].as<u64>()); else BUILD_KEYS(2, false, ctx->d_key[0][1].as<u64>()); }
#undef BUILD_KEYS
    STAGE_END(UMIGPU_STAGE_KEYS);

    // ---- K2 sort ----
    STAGE_BEGIN(UMIGPU_STAGE_SORT);
    int cur = 0;
    bool seg_done = false;
    rc = run_sort_presorted(ctx, n, lay, fused_table, &seg_done);
    if (rc) return rc;
    ctx->used_seg_sort = seg_done;
    if (seg_done) cur = 1;
    else { rc = run_sort(ctx, n, lay.nw, rs_plan(lay.total_bits), &cur); if (rc) return rc; }
    STAGE_END(UMIGPU_STAGE_SORT);
    ctx->sorted_cur = cur;
    SortedKeys sk{ctx->d_key[cur][0].as<u64>(), lay.nw == 2 ? ctx->d_key[cur][1].as<u64>() : nullptr, lay.umi_bits};
    const u32 *sorted_idx = ctx->d_idx[cur].as<u32>();

    // ---- K3 unique / count / merge ----
    STAGE_BEGIN(UMIGPU_STAGE_UNIQUE);
    CK(ctx->d_useg.reserve((n + 1) * 4));
    CK(ctx->d_rep.reserve(n * 8));
    CK(ctx->d_planes.reserve(n * 8));
    CK(ctx->d_ucode.reserve(n * 8));
    if (has_n) CK(ctx->d_nplane.reserve(n * 4));
    CK(ctx->d_bhead.reserve(n));
    const bool weighted = ctx->have_weight == 1;
    ctx->st_weighted = weighted;
    if (weighted) { CK(ctx->d_wsum.reserve(n * 4)); CK(cudaMemsetAsync(ctx->d_wsum.p, 0, n * 4, ctx->stream)); }
    if (want_labels) CK(ctx->d_read_uid.reserve(n * 4));
    const bool use_score = ctx->have_score == 1 && cfg.merge != UMIGPU_MERGE_ANY;
    UniqueEmit ue;
    ue.sk = sk; ue.n = n; ue.idx = sorted_idx; ue.score = use_score ? ctx->d_score.as<i32>() : nullptr;
    ue.weight = weighted ? ctx->d_weight.as<i32>() : nullptr; ue.L = lay.umi_len; ue.has_n = lay.has_n;
    ue.useg = ctx->d_useg.as<u32>(); ue.planes = ctx->d_planes.as<uint2>(); ue.nplane = ctx->d_nplane.as<u32>(); ue.ucode = ctx->d_ucode.as<u64>();
    ue.bhead = ctx->d_bhead.as<u8>(); ue.rep = ctx->d_rep.as<unsigned long long>(); ue.wsum = ctx->d_wsum.as<i32>();
    ue.read_uid = want_labels ? ctx->d_read_uid.as<u32>() : nullptr;
    ue.pend_uid = 0xffffffffu; ue.pend_val = 0; ue.pend_w = 0;
    if (lay.nw == 1 && !weighted && !getenv("UMIGPU_UNIQUE_TWO_PASS")) {
        // one-word keys: ONE pass (ticketed tiles, decoupled look-back); rep[] is zeroed by the tiles that own its slots
        const u64 ntiles = ceil_div_u64(n, SCAN_TILE);
        CK(ctx->d_tiles.reserve(ntiles * sizeof(u32)));
        CK(ctx->d_tilestate.reserve(2 * ntiles * 8));
        unsigned long long *state = ctx->d_tilestate.as<unsigned long long>();
        CK(cudaMemsetAsync(state, 0, ntiles * 8, ctx->stream));
        CK(cudaMemsetAsync(&sc->sort_ticket, 0, 4, ctx->stream));
        LAUNCH(unique_onepass_kernel, (u32)ntiles, SCAN_THREADS, ue, n, (u32)ntiles, state, &sc->sort_ticket, ctx->d_tiles.as<u32>(), state + ntiles,
               &sc->n_unique, &sc->sort_err);
        LAUNCH(unique_carry_kernel, grid_for(ntiles, 256), 256, (u32)ntiles, (const u32 *)ctx->d_tiles.p, (const unsigned long long *)(state + ntiles),
               ctx->d_rep.as<unsigned long long>());
    } else if (lay.nw == 1) {
        // three phases, the apply phase specialised (vector loads, predecessor by shuffle)
        CK(cudaMemsetAsync(ctx->d_rep.p, 0, n * 8, ctx->stream));
        const u64 ntiles = ceil_div_u64(n, SCAN_TILE);
        CK(ctx->d_tiles.reserve(ntiles * sizeof(u32)));
        u32 *ts = ctx->d_tiles.as<u32>();
        LAUNCH((scan_tile_sums<u32, HeadFlag>), (u32)ntiles, SCAN_THREADS, HeadFlag{sk}, n, ts);
        LAUNCH((scan_spine<u32>), 1, 1024, ts, ntiles, &sc->n_unique);
        LAUNCH(unique_apply1_kernel, (u32)ntiles, SCAN_THREADS, ue, n, (const u32 *)ts);
    } else {
        CK(cudaMemsetAsync(ctx->d_rep.p, 0, n * 8, ctx->stream));
        rc = run_scan(ctx, HeadFlag{sk}, ue, n, &sc->n_unique);
        if (rc) return rc;
    }
    rc = read_scalars(ctx);
    if (rc) return rc;
    if (ctx->h_sc->sort_err) return fail(ctx, UMIGPU_ERR_CUDA, "radix sort look-back exceeded its spin budget");
    const u32 U = ctx->h_sc->n_unique;
    ctx->n_unique = U;
    CK(ctx->d_freq.reserve((size_t)U * 4)); CK(ctx->d_thr.reserve((size_t)U * 4)); CK(ctx->d_repidx.reserve((size_t)U * 4));
    CK(ctx->d_label.reserve((size_t)U * 8)); CK(ctx->d_keep.reserve(U));
    const bool inf_thr = force_inf_thr || cfg.algo == UMIGPU_ALGO_CC || cfg.algo == UMIGPU_ALGO_ADJ_UPSTREAM;
    LAUNCH(unique_finalize_kernel, grid_for(U, 256), 256, U, (const u32 *)ctx->d_useg.p, (const unsigned long long *)ctx->d_rep.p,
           weighted ? (const i32 *)ctx->d_wsum.p : (const i32 *)nullptr, cfg.percentage, inf_thr ? 1 : 0, ctx->d_freq.as<i32>(),
           ctx->d_thr.as<i32>(), ctx->d_repidx.as<u32>(), ctx->d_label.as<unsigned long long>(), (const u64 *)ctx->d_ucode.p, lay.umi_len, lay.has_n,
           ctx->d_planes.as<uint2>(), ctx->d_nplane.as<u32>(), lay.umi_bits > 64 ? (const u64 *)ctx->d_umi2.p : (const u64 *)nullptr,
           lay.umi_bits > 64 ? (const u32 *)ctx->d_nmask.p : (const u32 *)nullptr);
    STAGE_END(UMIGPU_STAGE_UNIQUE);

    // ---- buckets ----
    STAGE_BEGIN(UMIGPU_STAGE_WORKLIST);
    CK(ctx->d_bstart.reserve(((size_t)U + 1) * 4));
    CK(ctx->d_ubkt.reserve((size_t)U * 4));
    rc = run_scan(ctx, BucketHead{ctx->d_bhead.as<u8>()}, BucketEmit{ctx->d_bstart.as<u32>(), U, ctx->d_ubkt.as<u32>()}, U, &sc->n_buckets);
    if (rc) return rc;
    // the statistics run off the device-resident bucket count, so that ONE read-back returns the count, the largest bucket and
    // the number / size of the big buckets (stage_neighbours then sizes the multi-index lists without asking again)
    LAUNCH(bucket_stats_kernel, (u32)std::min<u64>(grid_for(U, 256), (u64)ctx->num_sms * 8), 256, 0u, (const u32 *)&sc->n_buckets, (const u32 *)ctx->d_bstart.p, sc);
    rc = read_scalars(ctx);
    if (rc) return rc;
    const u32 B = ctx->h_sc->n_buckets;
    ctx->n_buckets = B;
    ctx->big_known = true;
    STAGE_END(UMIGPU_STAGE_WORKLIST);
    return UMIGPU_OK;
}

// ---- stage 2: K5 neighbour search over the buckets of ctx (unique arrays, bstart, ubkt, freq, thr) -> ctx->d_edges ----
// ctx->skip_bucket is left out (its search is shared by the shard group); a hot child evaluates only its band of row tiles.
struct MiPlan { int P, bpb; int part_lo[MI_MAX_PARTS], part_len[MI_MAX_PARTS]; u32 nbig, m_big; MiParams mi0; };

// Re-orders the unique UMIs of the big buckets of `src` so that part q is the most significant (keys, radix sort, gather).
// Everything that is written lives in the WORK context `ctx` (= src for the serial form, the helper for the threaded one).
static int mi_prepare(umigpu_ctx *ctx, const umigpu_ctx *src, const MiPlan &mp, int q, bool has_n) {
    const u32 U = src->n_unique, m_big = mp.m_big;
    const int pbits = mp.part_len[q] * mp.bpb, rbits = bits_for(mp.nbig - 1);
    CK(ctx->d_key[0][0].reserve((size_t)m_big * 8)); CK(ctx->d_key[1][0].reserve((size_t)m_big * 8));
    CK(ctx->d_idx[0].reserve((size_t)m_big * 4)); CK(ctx->d_idx[1].reserve((size_t)m_big * 4));
    CK(ctx->d_biguid.reserve((size_t)m_big * 4)); CK(ctx->d_miplanes.reserve((size_t)m_big * 8)); CK(ctx->d_miucode.reserve((size_t)m_big * 8));
    CK(ctx->d_miuid.reserve((size_t)m_big * 4)); if (has_n) CK(ctx->d_minplane.reserve((size_t)m_big * 4));
    LAUNCH(mi_keys_kernel, grid_for(U, 256), 256, U, (const u32 *)src->d_ubkt.p, (const u32 *)src->d_brank.p, (const u32 *)src->d_bstart.p,
           (const u32 *)src->d_bstartbig.p, (const u64 *)src->d_ucode.p, mp.part_lo[q] * mp.bpb,
           (unsigned long long)(pbits >= 64 ? ~0ull : ((1ull << pbits) - 1)), pbits, ctx->d_biguid.as<u32>(), ctx->d_key[0][0].as<u64>());
    int cur = 0;
    int r2 = run_sort(ctx, m_big, 1, rs_plan(pbits + rbits), &cur);
    if (r2) return r2;
    LAUNCH(mi_gather_kernel, grid_for(m_big, 256), 256, m_big, (const u32 *)ctx->d_idx[cur].p, (const u32 *)ctx->d_biguid.p,
           (const uint2 *)src->d_planes.p, has_n ? (const u32 *)src->d_nplane.p : (const u32 *)nullptr, (const u64 *)src->d_ucode.p,
           ctx->d_miplanes.as<uint2>(), ctx->d_minplane.as<u32>(), ctx->d_miucode.as<u64>(), ctx->d_miuid.as<u32>());
    return UMIGPU_OK;
}

// Passes 1..P-1 of the multi-index search, on the helper context `ctx` (its own stream, buffers and scalars) from a second host
// thread, while the caller runs pass 0: the two are independent chains of small kernels and scalar read-backs, and both append
// to the same edge sink.  Ends with the helper's stream drained.
static int mi_later_passes(umigpu_ctx *ctx, const umigpu_ctx *src, const MiPlan &mp, EdgeSink es, bool has_n, int cull, bool *dense) {
    CK(cudaSetDevice(ctx->cfg.device));
    DevScalars *sc = ctx->d_sc.as<DevScalars>();
    CK(cudaMemsetAsync(&sc->pairs_eval, 0, sizeof(u64), ctx->stream));
    CK(cudaMemsetAsync(&sc->sort_err, 0, 4, ctx->stream));
    CK(cudaMemcpyAsync(&sc->max_umis, &src->d_sc.as<DevScalars>()->max_umis, 4, cudaMemcpyDeviceToDevice, ctx->stream));
    ctx->ctr.n_tile_candidates = ctx->ctr.n_tile_items = ctx->ctr.n_block_pairs = 0;
    MiParams mi = mp.mi0;
    int rc = UMIGPU_OK;
    for (int q = 1; q < mp.P && !*dense; q++) {
        rc = mi_prepare(ctx, src, mp, q, has_n);
        if (rc) break;
        const NView view{ctx->d_miplanes.as<uint2>(), has_n ? ctx->d_minplane.as<u32>() : (const u32 *)nullptr, ctx->d_miucode.as<u64>(),
                         ctx->d_miuid.as<u32>(), src->d_bstartbig.as<u32>(), mp.nbig};
        mi.part = q;
        rc = neighbour_pass(ctx, view, mi, es, has_n, cull, true, dense, 0xffffffffu);
        if (rc) break;
    }
    const int rc2 = read_scalars(ctx);                         // drains the helper's stream
    if (!rc && rc2) rc = rc2;
    if (!rc && ctx->h_sc->sort_err) rc = fail(ctx, UMIGPU_ERR_CUDA, "radix sort look-back exceeded its spin budget");
    return rc;
}

static int stage_neighbours(umigpu_ctx *ctx, int mode) {
    const umigpu_config &cfg = ctx->cfg;
    DevScalars *sc = ctx->d_sc.as<DevScalars>();
    const KeyLayout &lay = ctx->lay;
    const bool has_n = lay.has_n != 0;
    const u32 U = ctx->n_unique, B = ctx->n_buckets;
    const u32 skip = ctx->skip_bucket;
    int rc;
    const bool need_edges = (mode == RUN_EDGES_ONLY || cfg.algo != UMIGPU_ALGO_ADJ) && cfg.k > 0 && U > B;
    ctx->st_need_edges = need_edges;
    const int cull = (cfg.flags & UMIGPU_FLAG_NO_CULL) ? 0 : 1;
    const int k = cfg.k, L = lay.umi_len;
    const bool allow_blocks = !(cfg.flags & (UMIGPU_FLAG_KERNEL_DIRECT | UMIGPU_FLAG_KERNEL_TILES)) && k >= 1 && k <= 3 && !(has_n && L > 21) && cull;      // (with N beyond 21 nt the 3-bit codes exceed ucode: plane kernels only)
    u64 n_edges = 0;
    STAGE_BEGIN(UMIGPU_STAGE_NEIGHBOURS);
    // ---- multi-index preparation: which buckets are big, their compacted unique list ----
    const int P = k + 1;                                   // parts (pigeonhole)
    bool mi_on = need_edges && allow_blocks && !(cfg.flags & UMIGPU_FLAG_NO_MULTI_INDEX) && L >= 2 * P;
    MiPlan mp; memset(&mp, 0, sizeof mp);
    mp.P = P; mp.bpb = has_n ? 3 : 2;
    mp.mi0.part = -1; mp.mi0.big = MI_BIG;
    const MiParams &mi0 = mp.mi0;
    if (mi_on) {
        int hi = L;
        for (int q = 0; q < P; q++) {                      // part 0 = the most significant positions of the main order
            int len = L / P + (q < L % P ? 1 : 0);
            mp.part_len[q] = len; mp.part_lo[q] = hi - len; hi -= len;
            mp.mi0.pmask[q] = (u32)(((1ull << len) - 1) << mp.part_lo[q]);
            mp.mi0.cmask[q] = (((len * mp.bpb) >= 64 ? ~0ull : ((1ull << (len * mp.bpb)) - 1))) << (mp.part_lo[q] * mp.bpb);
        }
        CK(ctx->d_brank.reserve((size_t)B * 4)); CK(ctx->d_bigbid.reserve((size_t)B * 4));
        // counts already on the host (stage_group's last read-back) unless a bucket is left out or this is a hot child
        const bool known = ctx->big_known && skip == 0xffffffffu && !getenv("UMIGPU_BIG_READBACK");
        if (known && ctx->h_sc->n_big_all == 0) mi_on = false;
        else {
            rc = run_scan(ctx, BucketIsBig{ctx->d_bstart.as<u32>(), MI_BIG, skip}, BucketBigEmit{ctx->d_brank.as<u32>(), ctx->d_bigbid.as<u32>(), B}, B, &sc->n_big);
            if (rc) return rc;
            if (known) mp.nbig = ctx->h_sc->n_big_all;
            else {
                rc = read_scalars(ctx);
                if (rc) return rc;
                mp.nbig = ctx->h_sc->n_big;
            }
            if (mp.nbig == 0) mi_on = false;
        }
        if (mi_on) {
            CK(ctx->d_bstartbig.reserve(((size_t)mp.nbig + 1) * 4));
            rc = run_scan(ctx, BigSize{ctx->d_bstart.as<u32>(), ctx->d_bigbid.as<u32>()}, BigStartEmit{ctx->d_bstartbig.as<u32>(), mp.nbig}, mp.nbig, &sc->m_big);
            if (rc) return rc;
            if (known) mp.m_big = ctx->h_sc->m_big_all;
            else {
                rc = read_scalars(ctx);
                if (rc) return rc;
                mp.m_big = ctx->h_sc->m_big;
            }
        }
    }
    // the later passes run beside pass 0 on a helper context (second host thread, own stream); UMIGPU_MI_SERIAL=1 = one after the other
    const bool threaded = mi_on && P > 1 && !getenv("UMIGPU_MI_SERIAL");
    if (threaded && !ctx->helper) {
        umigpu_config c = ctx->cfg;
        c.stream = nullptr;
        c.flags &= ~UMIGPU_FLAG_LABELS;
        rc = umigpu_create(&c, &ctx->helper);
        if (rc) { ctx->err = g_last_error; return rc; }
        ctx->helper->is_child = true;
    }

    // ---- K5 neighbours ----
    if (need_edges) {
        u64 cap = std::max<u64>((u64)1 << 20, (u64)U * 8);
        if (ctx->d_edges.cap / sizeof(uint2) > cap) cap = ctx->d_edges.cap / sizeof(uint2);
        const NView main_view{ctx->d_planes.as<uint2>(), has_n ? ctx->d_nplane.as<u32>() : (const u32 *)nullptr, ctx->d_ucode.as<u64>(), nullptr,
                              ctx->d_bstart.as<u32>(), B};
        for (int attempt = 0; attempt < 3; attempt++) {
            CK(ctx->d_edges.reserve(cap * sizeof(uint2)));
            CK(cudaMemsetAsync(&sc->edge_count, 0, sizeof(u64), ctx->stream));
            CK(cudaMemsetAsync(&sc->pairs_eval, 0, sizeof(u64), ctx->stream));
            ctx->ctr.n_tile_candidates = ctx->ctr.n_tile_items = ctx->ctr.n_block_pairs = 0;
            ctx->used_direct = false; ctx->direct_pairs = 0;
            EdgeSink es{ctx->d_edges.as<uint2>(), (unsigned long long *)&sc->edge_count, cap, ctx->d_freq.as<i32>(), ctx->d_thr.as<i32>()};
            // Fork: the small buckets depend on nothing pass 0 produces and have no host synchronisation: side stream.  In the
            // serial form the re-ordering for pass 1 goes there as well.
            CK(cudaEventRecord(ctx->ev_fork, ctx->stream));
            CK(cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0));
            {
                cudaStream_t main_stream = ctx->stream;
                ctx->stream = ctx->side;
                int r2 = [&]() -> int {
                    LAUNCH(small_buckets_kernel, grid_for((u64)B * 32, 256), 256, B, (const u32 *)ctx->d_bstart.p, (const uint2 *)ctx->d_planes.p,
                           has_n ? (const u32 *)ctx->d_nplane.p : (const u32 *)nullptr, cfg.k, es, (unsigned long long *)&sc->pairs_eval, skip);
                    if (mi_on && P > 1 && !threaded) return mi_prepare(ctx, ctx, mp, 1, has_n);
                    return UMIGPU_OK;
                }();
                cudaError_t ej = cudaEventRecord(ctx->ev_join, ctx->side);
                ctx->stream = main_stream;
                if (r2) return r2;
                CK(ej);
            }
            bool dense = false, dense_later = false;
            int later_rc = UMIGPU_OK;
            std::thread later;
            umigpu_ctx *hp = ctx->helper;
            if (threaded && mi_on) {
                hp->band = ctx->band; hp->n_bands = ctx->n_bands; hp->lay = ctx->lay;
                // the helper's stream starts after the counters of this attempt have been reset on the main stream
                CK(cudaStreamWaitEvent(hp->stream, ctx->ev_fork, 0));
                // (no exception may cross the C ABI: if the thread cannot be started the passes run here, after pass 0)
                try { later = std::thread([&] { later_rc = mi_later_passes(hp, ctx, mp, es, has_n, cull, &dense_later); }); }
                catch (...) { later_rc = -1000; }
            }
            // pass 0: every bucket with more than 32 unique UMIs in the main order (big buckets filtered on part 0)
            MiParams mi = mi0; mi.part = mi_on ? 0 : -1;
            rc = neighbour_pass(ctx, main_view, mi, es, has_n, cull, allow_blocks, &dense, skip);
            if (later.joinable()) later.join();
            else if (later_rc == -1000) later_rc = mi_later_passes(hp, ctx, mp, es, has_n, cull, &dense_later);
            cudaError_t ew = cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0);      // join (also before a restart or an error return)
            if (threaded && mi_on) {
                ctx->launches += hp->launches; hp->launches = 0;
                if (later_rc) return fail(ctx, later_rc, "%s", hp->err.c_str());
            }
            if (rc) return rc;
            CK(ew);
            if (dense || dense_later) { mi_on = false; continue; }          // culling does not thin this input out: restart without multi-index
            if (threaded && mi_on) {
                ctx->ctr.n_tile_candidates += hp->ctr.n_tile_candidates; ctx->ctr.n_tile_items += hp->ctr.n_tile_items;
                ctx->ctr.n_block_pairs += hp->ctr.n_block_pairs;
            } else {
                // passes 1..k: big buckets only, re-ordered so that part q is the most significant
                for (int q = 1; mi_on && q < P; q++) {
                    if (q > 1) { rc = mi_prepare(ctx, ctx, mp, q, has_n); if (rc) return rc; }
                    const NView view{ctx->d_miplanes.as<uint2>(), has_n ? ctx->d_minplane.as<u32>() : (const u32 *)nullptr, ctx->d_miucode.as<u64>(),
                                     ctx->d_miuid.as<u32>(), ctx->d_bstartbig.as<u32>(), mp.nbig};
                    mi.part = q;
                    rc = neighbour_pass(ctx, view, mi, es, has_n, cull, true, &dense, 0xffffffffu);
                    if (rc) return rc;
                    if (dense) break;
                }
                if (dense) { mi_on = false; continue; }
            }
            rc = read_scalars(ctx);
            if (rc) return rc;
            if (ctx->h_sc->sort_err) return fail(ctx, UMIGPU_ERR_CUDA, "radix sort look-back exceeded its spin budget");
            n_edges = ctx->h_sc->edge_count;
            if (n_edges <= cap) break;
            if (attempt == 2) return fail(ctx, UMIGPU_ERR_CUDA, "edge count changed between passes");
            cap = n_edges;                 // exact count is known now: one more pass with the right size
        }
    } else {
        rc = read_scalars(ctx);
        if (rc) return rc;
    }
    STAGE_END(UMIGPU_STAGE_NEIGHBOURS);
    ctx->n_edges = n_edges;
    ctx->ctr.n_buckets = B; ctx->ctr.total_umis = U; ctx->ctr.max_umis = ctx->h_sc->max_umis;
    ctx->ctr.unordered_pairs = ctx->h_sc->pairs;
    ctx->ctr.pairs_evaluated = ctx->h_sc->pairs_eval + (ctx->used_direct ? ctx->direct_pairs : 0);
    if (threaded && mi_on && ctx->helper) ctx->ctr.pairs_evaluated += ctx->helper->h_sc->pairs_eval;
    ctx->ctr.n_edges = n_edges;
    return UMIGPU_OK;
}

// ---- stage 3: K6 cluster over ctx->d_edges[0, ctx->n_edges) ----
static int stage_cluster(umigpu_ctx *ctx) {
    const umigpu_config &cfg = ctx->cfg;
    DevScalars *sc = ctx->d_sc.as<DevScalars>();
    const u32 U = ctx->n_unique;
    const u64 n_edges = ctx->n_edges;
    int rc;
    STAGE_BEGIN(UMIGPU_STAGE_CLUSTER);
    u8 *keep = ctx->d_keep.as<u8>();
    unsigned long long *label = ctx->d_label.as<unsigned long long>();
    const uint2 *edges = ctx->d_edges.as<uint2>();
    u32 egrid = (u32)std::min<u64>(std::max<u64>(1, ceil_div_u64(n_edges, 256)), (u64)ctx->num_sms * 16);
    u64 sweeps = 0;
    if (cfg.algo == UMIGPU_ALGO_ADJ || n_edges == 0) {
        CK(cudaMemsetAsync(keep, 1, U, ctx->stream));     // adjacency.rs:56 removes only the query itself
    } else if (cfg.algo == UMIGPU_ALGO_ADJ_UPSTREAM) {
        CK(ctx->d_state.reserve(U)); CK(ctx->d_blocked.reserve(U)); CK(ctx->d_prio.reserve((size_t)U * 8));
        CK(cudaMemsetAsync(ctx->d_state.p, 0, U, ctx->stream)); CK(cudaMemsetAsync(ctx->d_blocked.p, 0, U, ctx->stream));
        CK(cudaMemcpyAsync(ctx->d_prio.p, label, (size_t)U * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        const unsigned long long *prio = ctx->d_prio.as<unsigned long long>();
        for (;;) {
            CK(cudaMemsetAsync(&sc->changed, 0, 4, ctx->stream));
            LAUNCH(mis_edge_kernel, egrid, 256, edges, n_edges, prio, ctx->d_state.as<u8>(), ctx->d_blocked.as<u8>());
            LAUNCH(mis_node_kernel, grid_for(U, 256), 256, U, ctx->d_state.as<u8>(), ctx->d_blocked.as<u8>(), sc);
            sweeps++;
            rc = read_scalars(ctx);
            if (rc) return rc;
            if (!ctx->h_sc->changed) break;
        }
        LAUNCH(mis_label_kernel, egrid, 256, edges, n_edges, prio, (const u8 *)ctx->d_state.p, label);
        LAUNCH(mis_keep_kernel, grid_for(U, 256), 256, U, (const u8 *)ctx->d_state.p, keep);
    } else {
      // Three schedules for the same unique fixpoint (label[v] = earliest visited UMI that reaches v):
      //  (1) plain sweeps over the edge list, in place (a label can travel several hops per sweep); from the third sweep on an
      //      edge is skipped unless its source was lowered since the sweep before (stamp array).  Small graphs settle here.
      //  (2) frontier rounds (edges bucketed by source once, then only the out-edges of lowered UMIs): graphs between
      //      UMIGPU_FRONTIER_MIN_EDGES (2 Mi) and UMIGPU_SV_MIN_EDGES that are still moving while few labels change per sweep.
      //  (3) two-phase (cluster.cuh): union-find over the mutual edges, contracted propagation, expand — work independent of
      //      the diameter.  Graphs of >= UMIGPU_SV_MIN_EDGES (default 4 Mi) edges go there after UMIGPU_PLAIN_ROUNDS (default 0)
      //      rounds of 4 plain sweeps: the hot locus of C5 (3.1e7 edges, 40 sweeps = 10.9 ms in round 1) is its case.
      bool converged = false;
      const char *e_sv = getenv("UMIGPU_SV_MIN_EDGES"), *e_pr = getenv("UMIGPU_PLAIN_ROUNDS"), *e_fr = getenv("UMIGPU_FRONTIER_MIN_EDGES");
      const u64 sv_min = e_sv ? strtoull(e_sv, nullptr, 10) : (u64)(4u << 20);
      const int plain_rounds = e_pr ? atoi(e_pr) : 0;
      const u64 fr_min = e_fr ? strtoull(e_fr, nullptr, 10) : (u64)(2u << 20);
      // a saturated UMI space (fastq single bucket: 14 edges per UMI) is a small-world graph: a handful of plain sweeps settle
      // it (C4: 5.9 ms) while union-find + contraction over 1.7e8 mostly mutual edges costs twice that (12.6 ms measured)
      const bool big_graph = n_edges >= sv_min && (e_sv != nullptr || n_edges < 8ull * U);
      const bool sv_forced = e_sv != nullptr || e_pr != nullptr;          // the tests drive the two-phase scheme through these
      const bool may_frontier = fr_min != 0 && n_edges >= fr_min && !sv_forced && !big_graph && U < 0xfffffff0u;
      CK(ctx->d_stamp.reserve((size_t)U * 4));
      CK(cudaMemsetAsync(ctx->d_stamp.p, 0, (size_t)U * 4, ctx->stream));
      u32 *stamp = ctx->d_stamp.as<u32>();
      u32 sweep_no = 0;
      // UMIGPU_FRONTIER_FORCE=1 (tests): no plain sweeps at all, the first frontier is every UMI
      bool use_frontier = getenv("UMIGPU_FRONTIER_FORCE") != nullptr && !sv_forced && U < 0xfffffff0u;
      // sweeps until the fixpoint, in batches of 4 with one read-back; *n_ptr (optional) = device-resident edge count
      // later_batch: sweeps per read-back after the first batch of four (the contracted graph of the two-phase scheme is shallow:
      // two more usually settle it, four would mostly stream edges for nothing)
      auto sweep_batches = [&](const uint2 *el, u64 ne, const unsigned long long *n_ptr, int max_rounds, bool allow_frontier, int later_batch) -> int {
          const u32 g = (u32)std::min<u64>(std::max<u64>(1, ceil_div_u64(ne, 256)), (u64)ctx->num_sms * 16);
          for (int round = 0; !converged && (max_rounds < 0 || round < max_rounds); round++) {
              CK(cudaMemsetAsync(&sc->changed, 0, 4, ctx->stream));
              const int nb = round == 0 ? 4 : later_batch;
              for (int i = 0; i < nb; i++) {
                  if (i == nb - 1) CK(cudaMemsetAsync(&sc->n_lowered, 0, 4, ctx->stream));
                  ++sweep_no;
                  if (sweep_no <= 2) LAUNCH(label_sweep_kernel<false>, g, 256, el, ne, n_ptr, label, sc, stamp, sweep_no);
                  else               LAUNCH(label_sweep_kernel<true>, g, 256, el, ne, n_ptr, label, sc, stamp, sweep_no);
              }
              sweeps += nb;
              int r2 = read_scalars(ctx);
              if (r2) return r2;
              converged = !ctx->h_sc->changed || ctx->h_sc->n_lowered == 0;
              // still moving, but the last sweep lowered few labels compared with the edges it streamed: the frontier form pays
              if (!converged && allow_frontier && (u64)ctx->h_sc->n_lowered * 16 < ne) { use_frontier = true; break; }
          }
          return UMIGPU_OK;
      };
      if (!use_frontier) { rc = sweep_batches(edges, n_edges, nullptr, big_graph ? plain_rounds : -1, may_frontier, 4); if (rc) return rc; }
      if (!converged && use_frontier) {
          // CSR by source: out-degree histogram, scan, scatter (no sort: a row's order does not matter)
          if (n_edges >= 0xffffffffull) return fail(ctx, UMIGPU_ERR_UNSUPPORTED, "more than 2^32 edges in one batch");
          CK(ctx->d_rowptr.reserve(((size_t)U + 2) * 4));
          CK(ctx->d_comp.reserve(std::max<size_t>((size_t)U * 4, 4)));                 // row fill cursors (the two-phase scheme's buffer, unused here)
          CK(ctx->d_cedges.reserve(std::max<u64>(n_edges, 1) * sizeof(u32)));          // column indices
          CK(ctx->d_front[0].reserve((size_t)U * 4)); CK(ctx->d_front[1].reserve((size_t)U * 4));
          CK(cudaMemsetAsync(ctx->d_front[1].p, 0, (size_t)U * 4, ctx->stream));       // degrees live in the second frontier buffer until the scan
          LAUNCH(csr_degree_kernel, egrid, 256, edges, n_edges, ctx->d_front[1].as<u32>());
          rc = run_scan(ctx, DegreeOf{ctx->d_front[1].as<u32>()}, RowEmit{ctx->d_rowptr.as<u32>(), ctx->d_comp.as<u32>(), U}, U, nullptr);
          if (rc) return rc;
          LAUNCH(csr_fill_kernel, egrid, 256, edges, n_edges, ctx->d_comp.as<u32>(), ctx->d_cedges.as<u32>());
          const u32 *col = ctx->d_cedges.as<u32>();
          CK(cudaMemsetAsync(&sc->frontier_cnt[0], 0, 8, ctx->stream));
          LAUNCH(frontier_init_kernel, grid_for(U, 256), 256, U, (const u32 *)stamp, sweep_no, ctx->d_front[0].as<u32>(), &sc->frontier_cnt[0]);
          const u32 fgrid = (u32)ctx->num_sms * 8;
          int in = 0;
          u64 rounds = 0;
          while (!converged && rounds < 65536) {
              for (int i = 0; i < 8; i++) {
                  CK(cudaMemsetAsync(&sc->frontier_cnt[in ^ 1], 0, 4, ctx->stream));
                  LAUNCH(frontier_relax_kernel, fgrid, 256, (const u32 *)ctx->d_rowptr.p, col, label, stamp, (const u32 *)ctx->d_front[in].p,
                         (const u32 *)&sc->frontier_cnt[in], ctx->d_front[in ^ 1].as<u32>(), &sc->frontier_cnt[in ^ 1], ++sweep_no);
                  in ^= 1;
              }
              rounds += 8; sweeps += 8;
              rc = read_scalars(ctx);
              if (rc) return rc;
              converged = ctx->h_sc->frontier_cnt[in] == 0;
          }
          if (!converged) return fail(ctx, UMIGPU_ERR_CUDA, "label propagation did not settle in 65536 frontier rounds");
      }
      if (!converged) {
        // two-phase.  Labels may already have moved (plain rounds): any state with label[v] = priority of a UMI that reaches v
        // leads to the same fixpoint, so nothing is reset.
        CK(ctx->d_comp.reserve(std::max<size_t>((size_t)U * 4, 4)));
        u32 *comp = ctx->d_comp.as<u32>();
        LAUNCH(uf_init_kernel, grid_for(U, 256), 256, U, comp);
        LAUNCH(uf_union_kernel, egrid, 256, edges, n_edges, (const i32 *)ctx->d_freq.p, (const i32 *)ctx->d_thr.p, comp);
        LAUNCH(uf_flatten_kernel, grid_for(U, 256), 256, U, comp, label);
        CK(ctx->d_cedges.reserve(std::max<u64>(n_edges, 1) * sizeof(uint2)));
        CK(cudaMemsetAsync(&sc->scratch, 0, 8, ctx->stream));
        unsigned long long *n_c = (unsigned long long *)&sc->scratch;
        LAUNCH(contract_edges_kernel, egrid, 256, edges, n_edges, (const u32 *)comp, ctx->d_cedges.as<uint2>(), n_c);
        sweeps += 1;
        // Phase B on the contracted list (its length stays on the device); stamps restart: every contracted edge is relaxed once.
        // (Folding the first relaxation into the contraction kernel was tried: the sweeps run in batches of four with one
        // read-back, so it saved no batch and made the contraction 0.5 ms slower on C5.)
        CK(cudaMemsetAsync(ctx->d_stamp.p, 0, (size_t)U * 4, ctx->stream));
        sweep_no = 0;
        rc = sweep_batches(ctx->d_cedges.as<uint2>(), n_edges, n_c, -1, false, 2);
        if (rc) return rc;
        LAUNCH(expand_labels_kernel, grid_for(U, 256), 256, U, (const u32 *)comp, label);
      }
        LAUNCH(keep_from_label_kernel, grid_for(U, 256), 256, U, (const unsigned long long *)label, keep);
    }
    ctx->ctr.n_sweeps = sweeps;
    STAGE_END(UMIGPU_STAGE_CLUSTER);
    return UMIGPU_OK;
}

// ---- stage 4: K7 emit ----
static int stage_emit(umigpu_ctx *ctx, bool want_labels) {
    DevScalars *sc = ctx->d_sc.as<DevScalars>();
    const u64 n = ctx->n_reads;
    const u32 U = ctx->n_unique;
    u8 *keep = ctx->d_keep.as<u8>();
    unsigned long long *label = ctx->d_label.as<unsigned long long>();
    int rc;
    STAGE_BEGIN(UMIGPU_STAGE_EMIT);
    u64 n_words = ceil_div_u64(n, 32);
    CK(ctx->d_bitmap.reserve(n_words * 4));
    CK(cudaMemsetAsync(ctx->d_bitmap.p, 0, n_words * 4, ctx->stream));
    LAUNCH(mark_kept_kernel, grid_for(U, 256), 256, U, (const u8 *)keep, (const u32 *)ctx->d_repidx.p, ctx->d_bitmap.as<u32>());
    CK(ctx->d_kept.reserve((size_t)U * 8));
    // chunk table (push-order position -> caller's read index) for the device-side translation
    const u32 nch = (u32)ctx->chunks.size();
    {
        // context-owned pinned staging: no synchronisation needed for the upload (it stays untouched until the next run)
        if (2 * (size_t)nch > ctx->h_chunks_cap) {
            if (ctx->h_chunks) cudaFreeHost(ctx->h_chunks);
            ctx->h_chunks = nullptr; ctx->h_chunks_cap = 0;
            CK(cudaMallocHost((void **)&ctx->h_chunks, (2 * (size_t)nch + 64) * 8));
            ctx->h_chunks_cap = 2 * (size_t)nch + 64;
        }
        u64 *tab = ctx->h_chunks;
        for (u32 c = 0; c < nch; c++) { tab[c] = ctx->chunks[c].start; tab[nch + c] = ctx->chunks[c].first_index; }
        CK(ctx->d_chunks.reserve(2 * (size_t)nch * 8));
        CK(cudaMemcpyAsync(ctx->d_chunks.p, tab, 2 * (size_t)nch * 8, cudaMemcpyHostToDevice, ctx->stream));
    }
    ChunkMap cm{ctx->d_chunks.as<u64>(), ctx->d_chunks.as<u64>() + nch, nch, ctx->use_orig ? ctx->d_orig.as<u32>() : (const u32 *)nullptr};
    rc = run_scan(ctx, BitmapCount{ctx->d_bitmap.as<u32>()}, BitmapEmit{ctx->d_bitmap.as<u32>(), ctx->d_kept.as<u64>(), n_words, sc, cm}, n_words, nullptr);
    if (rc) return rc;
    if (want_labels) {
        CK(ctx->d_roots.reserve(n * 8)); CK(ctx->d_umirep.reserve(n * 8));
        LAUNCH(read_roots_kernel, grid_for(n, 256), 256, n, (const u32 *)ctx->d_read_uid.p, (const unsigned long long *)label,
               (const u32 *)ctx->d_repidx.p, cm, ctx->d_roots.as<u64>(), ctx->d_umirep.as<u64>());
    }
    STAGE_END(UMIGPU_STAGE_EMIT);
    STAGE_END(UMIGPU_STAGE_TOTAL);
    rc = read_scalars(ctx);
    if (rc) return rc;
    ctx->ctr.n_kept = ctx->h_sc->n_kept;
    ctx->run_ok = true;
    return UMIGPU_OK;
}

static int run_begin(umigpu_ctx *ctx) {
    if (!ctx) return fail(nullptr, UMIGPU_ERR_ARG, "null context");
    CK(cudaSetDevice(ctx->cfg.device));
    if (ctx->ran) return fail(ctx, UMIGPU_ERR_STATE, "run called twice without reset");
    memset(&ctx->ctr, 0, sizeof ctx->ctr);
    ctx->ctr.total_reads = ctx->n_records;          // deduplicate_sam.rs:100 counts every record, mapped or not
    ctx->ctr.n_unmapped = ctx->n_unmapped;
    ctx->ctr.n_unpaired = ctx->h_sc->n_unpaired; ctx->ctr.n_chimeric = ctx->h_sc->n_chimeric; ctx->ctr.n_mates_skipped = ctx->h_sc->n_mates_skipped;
    ctx->ran = true;
    ctx->n_unique = ctx->n_buckets = 0; ctx->n_edges = 0;
    ctx->skip_bucket = 0xffffffffu;
    ctx->big_known = false;
    ctx->run_ok = false;
    return UMIGPU_OK;
}

static int run_internal(umigpu_ctx *ctx, int mode, bool want_labels, bool force_inf_thr) {
    int rc = run_begin(ctx);
    if (rc) return rc;
    if (ctx->n_reads == 0) { ctx->run_ok = true; return UMIGPU_OK; }
    STAGE_BEGIN(UMIGPU_STAGE_TOTAL);
    rc = stage_group(ctx, want_labels, force_inf_thr);
    if (rc) return rc;
    rc = stage_neighbours(ctx, mode);
    if (rc) return rc;
    if (mode == RUN_EDGES_ONLY) { STAGE_END(UMIGPU_STAGE_TOTAL); ctx->run_ok = true; return UMIGPU_OK; }
    rc = stage_cluster(ctx);
    if (rc) return rc;
    return stage_emit(ctx, want_labels);
}

extern "C" int umigpu_run(umigpu_ctx *ctx) {
    if (!ctx) return fail(nullptr, UMIGPU_ERR_ARG, "null context");
    return run_internal(ctx, RUN_FULL, (ctx->cfg.flags & UMIGPU_FLAG_LABELS) != 0, false);
}

static int fetch_internal(umigpu_ctx *ctx, bool want_labels) {
    CK(cudaSetDevice(ctx->cfg.device));
    if (!ctx->ran) return fail(ctx, UMIGPU_ERR_STATE, "fetch before run");
    const u64 nk = ctx->ctr.n_kept, n = ctx->n_reads;
    if (nk > ctx->h_kept_cap) {
        if (ctx->h_kept) cudaFreeHost(ctx->h_kept);
        ctx->h_kept = nullptr; ctx->h_kept_cap = 0;
        CK(cudaMallocHost((void **)&ctx->h_kept, (nk + nk / 4 + 16) * 8));
        ctx->h_kept_cap = nk + nk / 4 + 16;
    }
    if (nk) CK(cudaMemcpyAsync(ctx->h_kept, ctx->d_kept.p, nk * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (want_labels && n) {
        if (n > ctx->h_roots_cap) {
            if (ctx->h_roots) cudaFreeHost(ctx->h_roots);
            if (ctx->h_umirep) cudaFreeHost(ctx->h_umirep);
            ctx->h_roots = ctx->h_umirep = nullptr; ctx->h_roots_cap = 0;
            CK(cudaMallocHost((void **)&ctx->h_roots, (n + 16) * 8));
            CK(cudaMallocHost((void **)&ctx->h_umirep, (n + 16) * 8));
            ctx->h_roots_cap = n + 16;
        }
        CK(cudaMemcpyAsync(ctx->h_roots, ctx->d_roots.p, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(ctx->h_umirep, ctx->d_umirep.p, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return UMIGPU_OK;
}

extern "C" int umigpu_fetch(umigpu_ctx *ctx, umigpu_result *out) {
    if (!ctx || !out) return fail(ctx, UMIGPU_ERR_ARG, "null argument");
    const bool want_labels = (ctx->cfg.flags & UMIGPU_FLAG_LABELS) != 0;
    int rc = fetch_internal(ctx, want_labels);
    if (rc) return rc;
    out->n_kept = ctx->ctr.n_kept;
    out->kept_read_index = ctx->h_kept;
    out->n_reads = ctx->n_reads;
    out->read_cluster_root = want_labels ? ctx->h_roots : nullptr;
    out->read_umi_rep = want_labels ? ctx->h_umirep : nullptr;
    out->counters = ctx->ctr;
    return UMIGPU_OK;
}

extern "C" int umigpu_finish(umigpu_ctx *ctx, umigpu_result *out) {
    int rc = umigpu_run(ctx);
    if (rc) return rc;
    return umigpu_fetch(ctx, out);
}

extern "C" int umigpu_get_counters(umigpu_ctx *ctx, umigpu_counters *out) {
    if (!ctx || !out) return fail(ctx, UMIGPU_ERR_ARG, "null argument");
    if (!ctx->ran) return fail(ctx, UMIGPU_ERR_STATE, "counters before run");
    *out = ctx->ctr;
    return UMIGPU_OK;
}

extern "C" void umigpu_result_free(umigpu_ctx *ctx) {
    if (!ctx) return;
    if (ctx->h_kept) cudaFreeHost(ctx->h_kept);
    if (ctx->h_roots) cudaFreeHost(ctx->h_roots);
    if (ctx->h_umirep) cudaFreeHost(ctx->h_umirep);
    ctx->h_kept = ctx->h_roots = ctx->h_umirep = nullptr; ctx->h_kept_cap = ctx->h_roots_cap = 0;
}

extern "C" int umigpu_stage_ms(umigpu_ctx *ctx, int stage, float *ms) {
    if (!ctx || !ms || stage < 0 || stage >= UMIGPU_N_STAGES) return fail(ctx, UMIGPU_ERR_ARG, "bad stage");
    *ms = 0.0f;
    if (!ctx->ev_ok[stage]) return UMIGPU_OK;
    CK(cudaSetDevice(ctx->cfg.device));
    CK(cudaEventSynchronize(ctx->ev[stage][1]));
    CK(cudaEventElapsedTime(ms, ctx->ev[stage][0], ctx->ev[stage][1]));
    return UMIGPU_OK;
}

extern "C" uint64_t umigpu_launch_count(umigpu_ctx *ctx, int reset) {
    if (!ctx) return 0;
    u64 v = ctx->launches;
    if (reset) ctx->launches = 0;
    return v;
}

// ------------------------------------------------------------------------------------------------
// Algorithm::apply / DataStruct shaped entries
// ------------------------------------------------------------------------------------------------
extern "C" int umigpu_cluster_bucket(umigpu_ctx *ctx, uint64_t n, const uint8_t *umi_ascii, const int32_t *freq,
                                     uint8_t *keep, int32_t *label) {
    if (!ctx) return fail(nullptr, UMIGPU_ERR_ARG, "null context");
    if (n && (!umi_ascii || !freq || !keep || !label)) return fail(ctx, UMIGPU_ERR_ARG, "null argument");
    if (n > 0x7fffffffull) return fail(ctx, UMIGPU_ERR_UNSUPPORTED, "bucket too large");
    int rc = umigpu_reset(ctx);
    if (rc) return rc;
    if (n == 0) return UMIGPU_OK;
    rc = push_common(ctx, n, nullptr, nullptr, nullptr, umi_ascii, nullptr, freq, 0, cudaMemcpyHostToDevice);
    if (rc) return rc;
    rc = run_internal(ctx, RUN_FULL, true, false);
    if (rc) return rc;
    rc = fetch_internal(ctx, true);
    if (rc) return rc;
    if (ctx->n_unique != n) return fail(ctx, UMIGPU_ERR_ARG, "umigpu_cluster_bucket: UMIs of a bucket must be distinct (%u unique of %llu)",
                                         ctx->n_unique, (unsigned long long)n);
    for (u64 i = 0; i < n; i++) { label[i] = (i32)ctx->h_roots[i]; keep[i] = ctx->h_roots[i] == i ? 1 : 0; }
    return UMIGPU_OK;
}

extern "C" int umigpu_remove_near(umigpu_ctx *ctx, uint64_t n, const uint8_t *umi_ascii, const int32_t *freq,
                                  const uint8_t *query, int32_t k, int32_t max_freq, uint8_t *out) {
    if (!ctx) return fail(nullptr, UMIGPU_ERR_ARG, "null context");
    if (!query || (n && (!umi_ascii || !freq || !out))) return fail(ctx, UMIGPU_ERR_ARG, "null argument");
    if (n > 0x7fffffffull) return fail(ctx, UMIGPU_ERR_UNSUPPORTED, "set too large");
    int rc = umigpu_reset(ctx);
    if (rc) return rc;
    const int L = (int)ctx->cfg.umi_len;
    // the query rides along as read n so that it is packed (and validated) by the same kernel
    std::vector<u8> all((n + 1) * L);
    if (n) memcpy(all.data(), umi_ascii, n * L);
    memcpy(all.data() + n * L, query, L);
    rc = push_common(ctx, n + 1, nullptr, nullptr, nullptr, all.data(), nullptr, nullptr, 0, cudaMemcpyHostToDevice);
    if (rc) return rc;
    rc = read_scalars(ctx);
    if (rc) return rc;
    if (ctx->h_sc->bad_base) return fail(ctx, UMIGPU_ERR_BAD_BASE, "Unknown character in UMI sequence");
    if (n == 0) return UMIGPU_OK;
    u64 q2; u32 qn;
    CK(cudaMemcpyAsync(&q2, ctx->d_umi2.as<u64>() + n, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(&qn, ctx->d_nmask.as<u32>() + n, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(ctx->d_freq.reserve(n * 4)); CK(ctx->d_keep.reserve(n));
    CK(cudaMemcpyAsync(ctx->d_freq.p, freq, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH(remove_near_kernel, grid_for(n, 256), 256, (u32)n, (const u64 *)ctx->d_umi2.p, (const u32 *)ctx->d_nmask.p, q2, qn,
           (const i32 *)ctx->d_freq.p, (int)k, max_freq, ctx->d_keep.as<u8>());
    CK(cudaMemcpyAsync(out, ctx->d_keep.p, n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return UMIGPU_OK;
}

// edges (unique ids) -> (src input index << 32 | dst input index)
__global__ void __launch_bounds__(256) edges_to_keys_kernel(const uint2 *__restrict__ edges, u64 n_edges, const u32 *__restrict__ rep_idx,
                                                            u64 *__restrict__ keys) {
    u64 e = (u64)blockIdx.x * 256 + threadIdx.x;
    if (e < n_edges) { uint2 ed = edges[e]; keys[e] = ((u64)rep_idx[ed.x] << 32) | rep_idx[ed.y]; }
}
__global__ void __launch_bounds__(256) csr_from_sorted_kernel(const u64 *__restrict__ keys, u64 n_edges, u32 n_rows,
                                                              u64 *__restrict__ row_ptr, u32 *__restrict__ col) {
    u64 e = (u64)blockIdx.x * 256 + threadIdx.x;
    if (e > n_edges) return;
    // row_ptr[r] = first e with src >= r
    u32 src_here = e < n_edges ? (u32)(keys[e] >> 32) : n_rows;
    u32 src_prev = e > 0 ? (u32)(keys[e - 1] >> 32) : 0;
    if (e == 0) for (u32 r = 0; r <= src_here; r++) row_ptr[r] = 0;
    else for (u32 r = src_prev + 1; r <= src_here; r++) row_ptr[r] = e;
    if (e < n_edges) col[e] = (u32)keys[e];
}

extern "C" int umigpu_neighbours(umigpu_ctx *ctx, uint64_t n, const uint8_t *umi_ascii, const int32_t *freq,
                                 int32_t apply_rule, uint64_t *row_ptr, uint32_t *col, uint64_t col_capacity,
                                 uint64_t *n_edges_out) {
    if (!ctx) return fail(nullptr, UMIGPU_ERR_ARG, "null context");
    if (!row_ptr || !n_edges_out || (n && (!umi_ascii || !freq))) return fail(ctx, UMIGPU_ERR_ARG, "null argument");
    if (n > 0x7fffffffull) return fail(ctx, UMIGPU_ERR_UNSUPPORTED, "bucket too large");
    int rc = umigpu_reset(ctx);
    if (rc) return rc;
    *n_edges_out = 0;
    if (n == 0) { row_ptr[0] = 0; return UMIGPU_OK; }
    rc = push_common(ctx, n, nullptr, nullptr, nullptr, umi_ascii, nullptr, freq, 0, cudaMemcpyHostToDevice);
    if (rc) return rc;
    rc = run_internal(ctx, RUN_EDGES_ONLY, false, apply_rule == 0);
    if (rc) return rc;
    if (ctx->n_unique != n) return fail(ctx, UMIGPU_ERR_ARG, "umigpu_neighbours: UMIs must be distinct");
    const u64 E = ctx->n_edges;
    *n_edges_out = E;
    if (E > col_capacity) return fail(ctx, UMIGPU_ERR_ARG, "col_capacity %llu < %llu edges", (unsigned long long)col_capacity, (unsigned long long)E);
    if (E == 0) { for (u64 i = 0; i <= n; i++) row_ptr[i] = 0; return UMIGPU_OK; }
    // sort (src, dst) in input-index space with the same radix sort, then cut rows
    CK(ctx->d_key[0][0].reserve(E * 8)); CK(ctx->d_key[1][0].reserve(E * 8));
    CK(ctx->d_idx[0].reserve(E * 4)); CK(ctx->d_idx[1].reserve(E * 4));
    LAUNCH(edges_to_keys_kernel, grid_for(E, 256), 256, (const uint2 *)ctx->d_edges.p, E, (const u32 *)ctx->d_repidx.p, ctx->d_key[0][0].as<u64>());
    int nb = bits_for(n - 1);
    int cur = 0;
    // keys are (src << 32 | dst): sort the dst bits, then the src bits
    {
        SortPlan plan; plan.npass = 0;
        for (int part = 0; part < 2; part++) {
            int done = 0, np = (nb + RS_RB - 1) / RS_RB;
            for (int i = 0; i < np; i++) { int b = (nb - done + (np - i) - 1) / (np - i); plan.p[plan.npass++] = {0, part * 32 + done, b}; done += b; }
        }
        if (plan.npass > 0) { rc = run_sort(ctx, E, 1, plan, &cur); if (rc) return rc; }
    }
    CK(ctx->d_rep.reserve((n + 1) * 8)); CK(ctx->d_kept.reserve(E * 4));
    LAUNCH(csr_from_sorted_kernel, grid_for(E + 1, 256), 256, (const u64 *)ctx->d_key[cur][0].p, E, (u32)n, ctx->d_rep.as<u64>(), ctx->d_kept.as<u32>());
    CK(cudaMemcpyAsync(row_ptr, ctx->d_rep.p, (n + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(col, ctx->d_kept.p, E * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return UMIGPU_OK;
}

extern "C" int umigpu_avg_qual(umigpu_ctx *ctx, uint64_t n, const uint8_t *qual, const uint64_t *offsets, int32_t *out) {
    if (!ctx) return fail(nullptr, UMIGPU_ERR_ARG, "null context");
    if (n == 0) return UMIGPU_OK;
    if (!qual || !offsets || !out) return fail(ctx, UMIGPU_ERR_ARG, "null argument");
    CK(cudaSetDevice(ctx->cfg.device));
    const u64 total = offsets[n];
    DevBuf dq, doff, dout;
    int rc = UMIGPU_OK;
#define CKQ(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { rc = fail(ctx, UMIGPU_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); goto done; } } while (0)
    CKQ(dq.reserve(total + 1)); CKQ(doff.reserve((n + 1) * 8)); CKQ(dout.reserve(n * 4));
    CKQ(cudaMemcpyAsync(dq.p, qual, total, cudaMemcpyHostToDevice, ctx->stream));
    CKQ(cudaMemcpyAsync(doff.p, offsets, (n + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    avg_qual_kernel<<<grid_for(n * 32, 256), 256, 0, ctx->stream>>>(n, dq.as<u8>(), doff.as<u64>(), dout.as<i32>());
    ctx->launches++;
    CKQ(cudaGetLastError());
    CKQ(cudaMemcpyAsync(out, dout.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CKQ(cudaStreamSynchronize(ctx->stream));
done:
#undef CKQ
    dq.release(); doff.release(); dout.release();
    return rc;
}

extern "C" int umigpu_int_peak(umigpu_ctx *ctx, double *lop3_ops_per_s, double *popc_ops_per_s) {
    if (!ctx || !lop3_ops_per_s || !popc_ops_per_s) return fail(ctx, UMIGPU_ERR_ARG, "null argument");
    CK(cudaSetDevice(ctx->cfg.device));
    CK(ctx->d_tiles.reserve(64));
    const u32 iters = 1 << 16, grid = (u32)ctx->num_sms * 8;
    cudaEvent_t a = ctx->ev[UMIGPU_STAGE_PACK][0], b = ctx->ev[UMIGPU_STAGE_PACK][1];
    float best[2] = {1e30f, 1e30f};
    for (int rep = 0; rep < 4; rep++) {
        for (int which = 0; which < 2; which++) {
            CK(cudaEventRecord(a, ctx->stream));
            if (which == 0) LAUNCH(int_peak_lop3_kernel, grid, 256, iters, 12345u + rep, ctx->d_tiles.as<u32>());
            else            LAUNCH(int_peak_popc_kernel, grid, 256, iters, 12345u + rep, ctx->d_tiles.as<u32>());
            CK(cudaEventRecord(b, ctx->stream));
            CK(cudaEventSynchronize(b));
            float ms; CK(cudaEventElapsedTime(&ms, a, b));
            if (rep > 0 && ms < best[which]) best[which] = ms;
        }
    }
    const double ops = (double)iters * 8.0 * 256.0 * grid;
    *lop3_ops_per_s = ops / (best[0] * 1e-3);
    *popc_ops_per_s = ops / (best[1] * 1e-3);
    return UMIGPU_OK;
}

// ------------------------------------------------------------------------------------------------
// multi-GPU sharding plan (host helper; buckets are independent)
// ------------------------------------------------------------------------------------------------
struct BKey { i64 pos; i32 tid; u8 rev; bool operator==(const BKey &o) const { return pos == o.pos && tid == o.tid && rev == o.rev; } };
struct BKeyHash {
    size_t operator()(const BKey &k) const {
        u64 x = (u64)k.pos * 0x9E3779B97F4A7C15ull ^ ((u64)(u32)k.tid << 1 | k.rev) * 0xC2B2AE3D27D4EB4Full;
        x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;
        return (size_t)x;
    }
};

extern "C" int umigpu_shard_plan(uint64_t n, const int32_t *tid, const int64_t *unclipped_pos, const uint8_t *is_reverse,
                                 int32_t n_shards, int32_t *shard_of_read, uint64_t *shard_cost) {
    if (n_shards < 1 || (n && (!tid || !unclipped_pos || !is_reverse || !shard_of_read)))
        return fail(nullptr, UMIGPU_ERR_ARG, "umigpu_shard_plan: bad argument");
    try {
        std::unordered_map<BKey, u32, BKeyHash> ids;
        ids.reserve(1 << 16);
        std::vector<u64> count;
        std::vector<u32> bid(n);
        for (u64 i = 0; i < n; i++) {
            BKey k{unclipped_pos[i], tid[i], (u8)(is_reverse[i] ? 1 : 0)};
            auto it = ids.find(k);
            u32 id;
            if (it == ids.end()) { id = (u32)count.size(); ids.emplace(k, id); count.push_back(0); } else id = it->second;
            count[id]++; bid[i] = id;
        }
        const size_t nb = count.size();
        std::vector<u32> order(nb);
        for (size_t b = 0; b < nb; b++) order[b] = (u32)b;
        // cost model: neighbour search ~ reads^2 (upper bound of N_b^2), plus a linear term for the HBM-bound stages
        auto cost = [&](u32 b) { return count[b] * count[b] + 64 * count[b]; };
        std::stable_sort(order.begin(), order.end(), [&](u32 a, u32 b) { return cost(a) > cost(b); });
        typedef std::pair<u64, i32> Load;   // (load, shard): least loaded first, ties to the lower shard id
        std::priority_queue<Load, std::vector<Load>, std::greater<Load>> pq;
        for (i32 s = 0; s < n_shards; s++) pq.push({0, s});
        std::vector<i32> shard_of_bucket(nb);
        std::vector<u64> loads(n_shards, 0);
        for (u32 b : order) {
            Load l = pq.top(); pq.pop();
            shard_of_bucket[b] = l.second;
            l.first += cost(b); loads[l.second] = l.first;
            pq.push(l);
        }
        for (u64 i = 0; i < n; i++) shard_of_read[i] = shard_of_bucket[bid[i]];
        if (shard_cost) for (i32 s = 0; s < n_shards; s++) shard_cost[s] = loads[s];
    } catch (const std::bad_alloc &) {
        return fail(nullptr, UMIGPU_ERR_NOMEM, "umigpu_shard_plan: out of host memory");
    }
    return UMIGPU_OK;
}

extern "C" void umigpu_free(void *p) { free(p); }

// ------------------------------------------------------------------------------------------------
// several GPUs of one box, inputs in ARBITRARY order: hash plan (umigpu_shard_plan), gather per shard, one context + one host
// thread per device, k-way merge.  Fallback of umigpu_dedup_sharded for inputs that are not coordinate-sorted.
// ------------------------------------------------------------------------------------------------

static int dedup_sharded_lpt(const umigpu_config *cfg, int32_t n_devices, const int32_t *device_ids, uint64_t n,
                                    const int32_t *tid, const int64_t *unclipped_pos, const uint8_t *is_reverse,
                                    const uint8_t *umi_ascii, const int32_t *score, uint64_t **kept, uint64_t *n_kept,
                                    umigpu_counters *counters) {
    if (!cfg || !device_ids || n_devices < 1 || !kept || !n_kept) return fail(nullptr, UMIGPU_ERR_ARG, "umigpu_dedup_sharded: bad argument");
    *kept = nullptr; *n_kept = 0;
    if (counters) memset(counters, 0, sizeof *counters);
    if (n == 0) return UMIGPU_OK;
    if (!tid || !unclipped_pos || !is_reverse || !umi_ascii) return fail(nullptr, UMIGPU_ERR_ARG, "umigpu_dedup_sharded: null input");
    const int L = (int)cfg->umi_len;
    std::vector<i32> shard(n);
    std::vector<u64> cost(n_devices);
    int rc = umigpu_shard_plan(n, tid, unclipped_pos, is_reverse, n_devices, shard.data(), cost.data());
    if (rc) return rc;
    struct Part { std::vector<u64> idx; std::vector<i32> tid, score; std::vector<i64> pos; std::vector<u8> rev, umi; std::vector<u64> kept;
                  umigpu_counters ctr; int rc = 0; std::string err; };
    std::vector<Part> parts(n_devices);
    for (u64 i = 0; i < n; i++) parts[shard[i]].idx.push_back(i);
    std::vector<std::thread> th;
    for (int s = 0; s < n_devices; s++) {
        th.emplace_back([&, s] {
            Part &p = parts[s];
            memset(&p.ctr, 0, sizeof p.ctr);
            const u64 m = p.idx.size();
            if (m == 0) return;
            p.tid.resize(m); p.pos.resize(m); p.rev.resize(m); p.umi.resize(m * (size_t)L); if (score) p.score.resize(m);
            for (u64 j = 0; j < m; j++) {
                const u64 i = p.idx[j];
                p.tid[j] = tid[i]; p.pos[j] = unclipped_pos[i]; p.rev[j] = is_reverse[i];
                memcpy(p.umi.data() + j * (size_t)L, umi_ascii + i * (size_t)L, (size_t)L);
                if (score) p.score[j] = score[i];
            }
            umigpu_config c = *cfg; c.device = device_ids[s]; c.stream = nullptr;
            umigpu_ctx *ctx = nullptr;
            p.rc = umigpu_create(&c, &ctx);
            if (p.rc) { p.err = umigpu_last_error(nullptr); return; }
            umigpu_result res;
            p.rc = umigpu_push_reads(ctx, m, p.tid.data(), p.pos.data(), p.rev.data(), p.umi.data(), score ? p.score.data() : nullptr, nullptr, 0);
            if (!p.rc) p.rc = umigpu_finish(ctx, &res);
            if (p.rc) { p.err = umigpu_last_error(ctx); umigpu_destroy(ctx); return; }
            p.kept.resize(res.n_kept);
            for (u64 j = 0; j < res.n_kept; j++) p.kept[j] = p.idx[res.kept_read_index[j]];      // local -> input index (idx is ascending)
            p.ctr = res.counters;
            umigpu_destroy(ctx);
        });
    }
    for (auto &t : th) t.join();
    u64 total = 0;
    for (Part &p : parts) { if (p.rc) return fail(nullptr, p.rc, "shard failed: %s", p.err.c_str()); total += p.kept.size(); }
    u64 *out = (u64 *)malloc(std::max<u64>(total, 1) * sizeof(u64));
    if (!out) return fail(nullptr, UMIGPU_ERR_NOMEM, "out of host memory");
    // k-way merge of ascending lists
    std::vector<size_t> at(n_devices, 0);
    typedef std::pair<u64, int> Head;
    std::priority_queue<Head, std::vector<Head>, std::greater<Head>> pq;
    for (int s = 0; s < n_devices; s++) if (!parts[s].kept.empty()) pq.push({parts[s].kept[0], s});
    u64 o = 0;
    while (!pq.empty()) {
        Head h = pq.top(); pq.pop();
        out[o++] = h.first;
        Part &p = parts[h.second];
        if (++at[h.second] < p.kept.size()) pq.push({p.kept[at[h.second]], h.second});
    }
    *kept = out; *n_kept = total;
    if (counters) {
        for (Part &p : parts) {
            counters->total_reads += p.ctr.total_reads; counters->n_buckets += p.ctr.n_buckets; counters->total_umis += p.ctr.total_umis;
            counters->max_umis = std::max(counters->max_umis, p.ctr.max_umis); counters->n_kept += p.ctr.n_kept;
            counters->unordered_pairs += p.ctr.unordered_pairs; counters->pairs_evaluated += p.ctr.pairs_evaluated; counters->n_edges += p.ctr.n_edges;
            counters->n_tile_items += p.ctr.n_tile_items; counters->n_tile_candidates += p.ctr.n_tile_candidates;
            counters->n_sweeps = std::max(counters->n_sweeps, p.ctr.n_sweeps); counters->n_block_pairs += p.ctr.n_block_pairs;
            counters->key_bits = std::max(counters->key_bits, p.ctr.key_bits);
        }
    }
    return UMIGPU_OK;
}

// ------------------------------------------------------------------------------------------------
// shard group: ONE dataset over several devices (deduplicate_sam.rs:207-213: buckets never interact)
//
// The coordinate-sorted read stream is cut at bucket boundaries into one contiguous slice per device (a contiguous H2D
// range, no gather on the host); every device runs the whole path on its slice and returns ascending kept indices, so the
// merged result is the concatenation in rank order.  No collective.
//
// The exception is a bucket that is bigger than a device's fair share (the hot locus of a skewed run): its reads stay on
// one device (the owner) up to K3, then its neighbour search is split.  The owner publishes the bucket's unique-UMI arrays
// in its EXCHANGE WINDOW (device memory the other devices of the group can address: peer access inside one process,
// CUDA IPC between processes); every device copies them over NVLink, searches the row tiles ti with ti % n == rank
// (stage_neighbours on a child context), and copies its edges into its region of the owner's window; the owner
// appends them to its edge list and clusters.  All transfers are cudaMemcpyAsync (copy engines over NVLink / NVSwitch);
// hand-over is a flag word written after the data in stream order and polled by the consumer's host thread — no kernel
// ever spins on a peer, so nothing can deadlock against an implicit device synchronisation (cudaFree).
// ------------------------------------------------------------------------------------------------
#define XCHG_MAX_RANKS 16
struct XchgSlot { unsigned long long count, epoch; };      // rank r -> owner: count (bit 63 = region overflow) first, epoch last
struct XchgHeader {
    unsigned long long info;        // hot_cnt | has_n << 32 | err << 33 | narrow codes << 34, written before `ready`
    unsigned long long ready;       // epoch of the arrays in this window
    XchgSlot slot[XCHG_MAX_RANKS];
    unsigned long long ucap, ecap, n_ranks;    // what the window was created with (checked at attach)
    unsigned long long pad[27];
};
static_assert(sizeof(XchgHeader) == 512, "exchange header layout");

// In-process groups hand over through HOST memory instead of flag words in the window: the producer's stream runs a host
// function (cudaLaunchHostFunc) after its data is in place, the consumer's host thread spins on a std::atomic.  (Polling a peer
// device's memory with small copies works between processes over CUDA IPC mappings, but inside one process with
// cudaDeviceEnablePeerAccess it returned stale words on an 8-GPU box — rank 0 "waited 10 s for the owner's hot bucket" that
// the owner had published; profiles/r2v_*.)
struct XchgHostFlags {
    std::atomic<unsigned long long> ready[XCHG_MAX_RANKS], info[XCHG_MAX_RANKS];
    std::atomic<unsigned long long> slot_epoch[XCHG_MAX_RANKS][XCHG_MAX_RANKS], slot_count[XCHG_MAX_RANKS][XCHG_MAX_RANKS];   // [owner][rank]
    XchgHostFlags() {
        for (int a = 0; a < XCHG_MAX_RANKS; a++) { ready[a].store(0); info[a].store(0); for (int b = 0; b < XCHG_MAX_RANKS; b++) { slot_epoch[a][b].store(0); slot_count[a][b].store(0); } }
    }
};
struct XchgHostSig { std::atomic<unsigned long long> *first = nullptr, *second = nullptr; unsigned long long v1 = 0, v2 = 0; };
static void CUDART_CB xchg_host_sig_fn(void *p) {
    XchgHostSig *s = static_cast<XchgHostSig *>(p);
    s->first->store(s->v1, std::memory_order_relaxed);
    s->second->store(s->v2, std::memory_order_release);        // the word the consumer waits for goes last
}

struct Xchg {
    int rank = 0, n = 1;
    bool ipc = false;
    std::shared_ptr<XchgHostFlags> hf_own, hf_shared;   // every rank makes one; an in-process group uses rank 0's
    XchgHostFlags *hf = nullptr;              // = hf_shared.get(); null between processes (flag words in the windows instead)
    XchgHostSig sig_pub, sig_put;
    char *local = nullptr;
    std::vector<char *> peer;       // window of every rank as this device addresses it (peer[rank] == local)
    std::vector<char> opened;       // peer[r] came from cudaIpcOpenMemHandle
    u64 ucap = 0, ecap = 0, region = 0, epoch = 0;
    size_t off_code = 0, off_freq = 0, off_inbox = 0, bytes = 0;
    unsigned long long *h_pin = nullptr;      // pinned: [0] info [1] ready [2] count [3] epoch, [8..] poll buffer
    cudaStream_t xs = nullptr;                // polling stream
    double timeout_s = 120.0;
    std::atomic<int> *abort = nullptr;        // in-process groups: set when a rank of the group failed (waits give up at once)
};

// UMIGPU_XCHG_DEBUG=1: one stderr line per hand-over step and rank (which rank waits for what, for how long)
static bool xchg_debug() { static int v = -1; if (v < 0) v = getenv("UMIGPU_XCHG_DEBUG") ? 1 : 0; return v == 1; }
#define XDBG(...) do { if (xchg_debug()) { fprintf(stderr, "[umigpu xchg] " __VA_ARGS__); fputc('\n', stderr); } } while (0)

static void xchg_release(umigpu_ctx *ctx) {
    Xchg *x = ctx->x;
    if (!x) return;
    for (size_t r = 0; r < x->peer.size(); r++) if (x->opened[r] && x->peer[r]) cudaIpcCloseMemHandle(x->peer[r]);
    if (x->local) cudaFree(x->local);
    if (x->h_pin) cudaFreeHost(x->h_pin);
    if (x->xs) cudaStreamDestroy(x->xs);
    delete x;
    ctx->x = nullptr;
}

extern "C" int umigpu_xchg_create(umigpu_ctx *ctx, int32_t rank, int32_t n_ranks, uint64_t max_hot_uniques, uint64_t max_hot_edges,
                                  uint8_t *ipc_handle_out /* 64 bytes, nullable */) {
    if (!ctx) return fail(nullptr, UMIGPU_ERR_ARG, "null context");
    if (n_ranks < 1 || n_ranks > XCHG_MAX_RANKS || rank < 0 || rank >= n_ranks)
        return fail(ctx, UMIGPU_ERR_ARG, "umigpu_xchg_create: rank %d of %d (at most %d ranks)", rank, n_ranks, XCHG_MAX_RANKS);
    CK(cudaSetDevice(ctx->cfg.device));
    CK(cudaStreamSynchronize(ctx->stream));
    xchg_release(ctx);
    Xchg *x = new (std::nothrow) Xchg();
    if (!x) return fail(ctx, UMIGPU_ERR_NOMEM, "out of host memory");
    ctx->x = x;
    x->rank = rank; x->n = n_ranks;
    x->ucap = std::max<u64>(max_hot_uniques, 64);
    x->region = std::max<u64>(ceil_div_u64(std::max<u64>(max_hot_edges, 1), (u64)n_ranks), 1024);
    x->ecap = x->region * (u64)n_ranks;
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    size_t o = sizeof(XchgHeader);
    x->off_code = o;   o = al(o + x->ucap * 8);       // sort codes of the hot bucket's unique UMIs (u32 each when they fit)
    x->off_freq = o;   o = al(o + x->ucap * 4);
    x->off_inbox = o;  o = al(o + x->ecap * 8);
    x->bytes = o;
    if (const char *e = getenv("UMIGPU_XCHG_TIMEOUT_S")) x->timeout_s = atof(e);
    CK(cudaMalloc((void **)&x->local, x->bytes));                 // cudaMalloc (not the async pool): the window is IPC-exportable
    XchgHeader h; memset(&h, 0, sizeof h);
    h.ucap = x->ucap; h.ecap = x->ecap; h.n_ranks = (u64)n_ranks;
    CK(cudaMemcpy(x->local, &h, sizeof h, cudaMemcpyHostToDevice));
    CK(cudaMallocHost((void **)&x->h_pin, (8 + 2 * XCHG_MAX_RANKS) * sizeof(unsigned long long)));
    CK(cudaStreamCreateWithFlags(&x->xs, cudaStreamNonBlocking));
    x->hf_own = std::make_shared<XchgHostFlags>();
    x->peer.assign((size_t)n_ranks, nullptr); x->opened.assign((size_t)n_ranks, 0);
    x->peer[(size_t)rank] = x->local;
    if (ipc_handle_out) {
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
        cudaIpcMemHandle_t hd;
        CK(cudaIpcGetMemHandle(&hd, x->local));
        memcpy(ipc_handle_out, &hd, 64);
    }
    return UMIGPU_OK;
}

static int xchg_check_peer(umigpu_ctx *ctx, int r) {
    Xchg *x = ctx->x;
    XchgHeader h;
    CK(cudaMemcpy(&h, x->peer[(size_t)r], sizeof h, cudaMemcpyDefault));
    if (h.ucap != x->ucap || h.ecap != x->ecap || h.n_ranks != (u64)x->n)
        return fail(ctx, UMIGPU_ERR_ARG, "exchange window of rank %d was created with other sizes (%llu uniques / %llu edges / %llu ranks)", r,
                    (unsigned long long)h.ucap, (unsigned long long)h.ecap, (unsigned long long)h.n_ranks);
    return UMIGPU_OK;
}

extern "C" int umigpu_xchg_attach_ipc(umigpu_ctx *ctx, const uint8_t *handles /* n_ranks x 64 bytes, rank order */) {
    if (!ctx || !ctx->x || !handles) return fail(ctx, UMIGPU_ERR_ARG, "umigpu_xchg_attach_ipc: no exchange window / null handles");
    CK(cudaSetDevice(ctx->cfg.device));
    Xchg *x = ctx->x;
    x->ipc = true;
    for (int r = 0; r < x->n; r++) {
        if (r == x->rank) continue;
        cudaIpcMemHandle_t hd;
        memcpy(&hd, handles + (size_t)r * 64, 64);
        void *p = nullptr;
        CK(cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
        x->peer[(size_t)r] = (char *)p; x->opened[(size_t)r] = 1;
        int rc = xchg_check_peer(ctx, r);
        if (rc) return rc;
    }
    return UMIGPU_OK;
}

extern "C" int umigpu_xchg_attach_local(umigpu_ctx *ctx, umigpu_ctx *const *group /* n_ranks contexts of THIS process, rank order */) {
    if (!ctx || !ctx->x || !group) return fail(ctx, UMIGPU_ERR_ARG, "umigpu_xchg_attach_local: no exchange window / null group");
    CK(cudaSetDevice(ctx->cfg.device));
    Xchg *x = ctx->x;
    if (!group[0] || !group[0]->x || !group[0]->x->hf_own) return fail(ctx, UMIGPU_ERR_ARG, "rank 0 of the group has no exchange window");
    x->hf_shared = group[0]->x->hf_own;          // hand-over words of the whole group: host memory of this process
    x->hf = x->hf_shared.get();
    for (int r = 0; r < x->n; r++) {
        if (r == x->rank) continue;
        umigpu_ctx *o = group[r];
        if (!o || !o->x || !o->x->local) return fail(ctx, UMIGPU_ERR_ARG, "rank %d of the group has no exchange window", r);
        if (o->cfg.device != ctx->cfg.device) {
            int can = 0;
            CK(cudaDeviceCanAccessPeer(&can, ctx->cfg.device, o->cfg.device));
            if (!can) return fail(ctx, UMIGPU_ERR_UNSUPPORTED, "device %d cannot address device %d (no peer access)", ctx->cfg.device, o->cfg.device);
            cudaError_t e = cudaDeviceEnablePeerAccess(o->cfg.device, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError(); else CK(e);
        }
        x->peer[(size_t)r] = o->x->local;
        int rc = xchg_check_peer(ctx, r);
        if (rc) return rc;
    }
    return UMIGPU_OK;
}

// combined (contig, position) key of the cuts: monotone in (tid, pos) for |pos| < 2^35
static inline i64 host_pos_key(i32 tid, i64 pos) { return (i64)((u64)(i64)tid << 36) + pos; }
extern "C" int64_t umigpu_pos_key(int32_t tid, int64_t unclipped_pos) { return host_pos_key(tid, unclipped_pos); }

__global__ void __launch_bounds__(256) slice_range_kernel(u64 n, const i32 *__restrict__ tid, const i64 *__restrict__ pos, DevScalars *sc) {
    const u64 stride = (u64)gridDim.x * 256;
    long long lo = 0x7fffffffffffffffLL, hi = (long long)0x8000000000000000LL;
    u32 bad = 0;
    for (u64 i = (u64)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) {
        const i64 p = pos[i];
        if (p >= (1LL << 35) || p < -(1LL << 35)) bad = 1;
        const long long k = (long long)((u64)(i64)tid[i] << 36) + p;
        lo = k < lo ? k : lo; hi = k > hi ? k : hi;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const long long l2 = __shfl_xor_sync(0xffffffffu, lo, o), h2 = __shfl_xor_sync(0xffffffffu, hi, o);
        lo = l2 < lo ? l2 : lo; hi = h2 > hi ? h2 : hi;
        bad |= __shfl_xor_sync(0xffffffffu, bad, o);
    }
    if (lane_id() == 0) {
        if (lo <= hi) { atomicMin((long long *)&sc->key_lo, lo); atomicMax((long long *)&sc->key_hi, hi); }
        if (bad) sc->hot_pad = 1;
    }
}
// sorted position of one read (push-order position `target`)
__global__ void __launch_bounds__(256) find_read_kernel(u64 n, const u32 *__restrict__ sorted_idx, u32 target, DevScalars *sc) {
    const u64 i = (u64)blockIdx.x * 256 + threadIdx.x;
    if (i < n && sorted_idx[i] == target) sc->scratch = i;
}
// ... -> its unique -> its bucket
__global__ void hot_locate_kernel(DevScalars *sc, const u32 *__restrict__ useg, u32 n_unique, const u32 *__restrict__ ubkt, const u32 *__restrict__ bstart) {
    const u64 p = sc->scratch;
    if (p == ~0ull) { sc->hot_bucket = 0xffffffffu; sc->hot_u0 = sc->hot_cnt = 0; return; }
    u32 lo = 0, hi = n_unique;               // last u with useg[u] <= p
    while (hi - lo > 1) { const u32 mid = (lo + hi) >> 1; if ((u64)useg[mid] <= p) lo = mid; else hi = mid; }
    const u32 b = ubkt[lo];
    sc->hot_bucket = b; sc->hot_u0 = bstart[b]; sc->hot_cnt = bstart[b + 1] - bstart[b];
}
// What crosses NVLink per unique UMI of the hot bucket is its sort code (4 bytes when 2 or 3 bits per base fit 32 bits) and
// its frequency: every rank pulls the arrays from the owner at the same time, so the owner's egress is the bottleneck, and
// bit planes, N plane and threshold are functions of those two.
__global__ void __launch_bounds__(256) hot_export_kernel(u32 n, const u64 *__restrict__ ucode, const i32 *__restrict__ freq, int narrow,
                                                         void *__restrict__ out_code, i32 *__restrict__ out_freq) {
    const u32 u = blockIdx.x * 256 + threadIdx.x;
    if (u >= n) return;
    if (narrow) reinterpret_cast<u32 *>(out_code)[u] = (u32)ucode[u]; else reinterpret_cast<u64 *>(out_code)[u] = ucode[u];
    out_freq[u] = freq[u];
}
__global__ void __launch_bounds__(256) hot_child_expand_kernel(u32 n_unique, const void *__restrict__ code_in, int narrow, const i32 *__restrict__ freq,
                                                               float percentage, int inf_thr, int L, int has_n, uint2 *__restrict__ planes,
                                                               u32 *__restrict__ nplane, u64 *__restrict__ ucode, i32 *__restrict__ thr,
                                                               u32 *__restrict__ bstart, u32 *__restrict__ ubkt) {
    const u32 u = blockIdx.x * 256 + threadIdx.x;
    if (u == 0) { bstart[0] = 0; bstart[1] = n_unique; }
    if (u >= n_unique) return;
    const u64 c = narrow ? (u64)reinterpret_cast<const u32 *>(code_in)[u] : reinterpret_cast<const u64 *>(code_in)[u];
    u32 p0, p1, pn;
    code_to_planes(c, L, has_n, p0, p1, pn);
    planes[u] = make_uint2(p0, p1);
    if (has_n) nplane[u] = pn;
    ucode[u] = c;
    thr[u] = inf_thr ? 0x7fffffff : dir_threshold(percentage, freq[u]);     // same expression as unique_finalize_kernel
    ubkt[u] = 0;
}
__global__ void __launch_bounds__(256) hot_append_kernel(const uint2 *__restrict__ in, u64 n, u32 u0, uint2 *__restrict__ out) {
    const u64 e = (u64)blockIdx.x * 256 + threadIdx.x;
    if (e < n) { const uint2 ed = in[e]; out[e] = make_uint2(ed.x + u0, ed.y + u0); }
}

// host-side wait on a word of a window: small peer reads on the polling stream until pred(value)
template <class Pred>
static int xchg_poll(Xchg *x, umigpu_ctx *ctx /* receives the error message */, const char *src, size_t bytes, Pred pred, const char *what) {
    unsigned long long *buf = x->h_pin + 8;
    const auto t0 = std::chrono::steady_clock::now();
    XDBG("rank %d (device %d) epoch %llu: waiting for %s", x->rank, ctx->cfg.device, (unsigned long long)x->epoch, what);
    for (u64 it = 0;; it++) {
        CK(cudaMemcpyAsync(buf, src, bytes, cudaMemcpyDefault, x->xs));
        CK(cudaStreamSynchronize(x->xs));
        if (pred(buf)) {
            XDBG("rank %d epoch %llu: got %s after %.3f ms (%llu polls)", x->rank, (unsigned long long)x->epoch, what,
                 1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(), (unsigned long long)it + 1);
            return UMIGPU_OK;
        }
        if (x->abort && x->abort->load()) return fail(ctx, UMIGPU_ERR_STATE, "shard group: rank %d gave up waiting for %s: another rank failed", x->rank, what);
        if ((it & 63) == 63) {
            const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (dt > x->timeout_s) return fail(ctx, UMIGPU_ERR_STATE, "shard group: rank %d waited %.0f s for %s (a rank of the group failed or never ran)", x->rank, dt, what);
        }
        if (it > 256) std::this_thread::yield();
    }
}

// in-process hand-over: spin on host words
template <class Pred>
static int xchg_wait_host(Xchg *x, umigpu_ctx *ctx, Pred pred, const char *what) {
    const auto t0 = std::chrono::steady_clock::now();
    XDBG("rank %d (device %d) epoch %llu: waiting (host flags) for %s", x->rank, ctx->cfg.device, (unsigned long long)x->epoch, what);
    for (u64 it = 0;; it++) {
        if (pred()) return UMIGPU_OK;
        if (x->abort && x->abort->load()) return fail(ctx, UMIGPU_ERR_STATE, "shard group: rank %d gave up waiting for %s: another rank failed", x->rank, what);
        if ((it & 1023) == 1023) {
            const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (dt > x->timeout_s) return fail(ctx, UMIGPU_ERR_STATE, "shard group: rank %d waited %.0f s for %s (a rank of the group failed or never ran)", x->rank, dt, what);
        }
        if (it > 64) std::this_thread::yield();
    }
}

// owner: find the hot bucket among this slice's buckets, copy its unique arrays into the window, publish
static int hot_publish(umigpu_ctx *ctx, const umigpu_hot *hot, u32 *u0_out) {
    Xchg *x = ctx->x;
    DevScalars *sc = ctx->d_sc.as<DevScalars>();
    const u64 n = ctx->n_reads;
    const bool has_n = ctx->lay.has_n != 0;
    u64 local = ~0ull;
    for (const Chunk &c : ctx->chunks) if (hot->read_index >= c.first_index && hot->read_index < c.first_index + c.n) local = c.start + (hot->read_index - c.first_index);
    u32 hb = 0xffffffffu, u0 = 0, uh = 0;
    const char *why = nullptr;
    int rc = UMIGPU_OK;
    if (ctx->use_orig) why = "the BAM feed cannot be combined with a split hot bucket";
    else if (ctx->lay.umi_bits > 64) why = "UMIs with N beyond 21 nt cannot be split over devices";
    else if (local == ~0ull || n == 0) why = "the hot bucket's read is not in the owner's slice";
    if (!why) {
        CK(cudaMemsetAsync(&sc->scratch, 0xff, 8, ctx->stream));
        LAUNCH(find_read_kernel, grid_for(n, 256), 256, n, (const u32 *)ctx->d_idx[ctx->sorted_cur].p, (u32)local, sc);
        LAUNCH(hot_locate_kernel, 1, 1, sc, (const u32 *)ctx->d_useg.p, ctx->n_unique, (const u32 *)ctx->d_ubkt.p, (const u32 *)ctx->d_bstart.p);
        rc = read_scalars(ctx);
        if (rc) return rc;
        hb = ctx->h_sc->hot_bucket; u0 = ctx->h_sc->hot_u0; uh = ctx->h_sc->hot_cnt;
        if (hb == 0xffffffffu) why = "the hot bucket's read was not found after the sort";
        else if ((u64)uh > x->ucap) why = "the exchange window is too small for the hot bucket's unique UMIs";
    }
    const int narrow = ctx->lay.umi_bits <= 32 ? 1 : 0;
    if (!why) {
        char *w = x->local;
        LAUNCH(hot_export_kernel, grid_for(uh, 256), 256, uh, (const u64 *)(ctx->d_ucode.as<u64>() + u0), (const i32 *)(ctx->d_freq.as<i32>() + u0), narrow,
               (void *)(w + x->off_code), (i32 *)(w + x->off_freq));
    }
    // info first, ready last (stream order = the order the words land in the window)
    x->h_pin[0] = why ? (1ull << 33) : ((unsigned long long)uh | ((unsigned long long)(has_n ? 1 : 0) << 32) | ((unsigned long long)narrow << 34));
    x->h_pin[1] = x->epoch;
    if (x->hf) {
        x->sig_pub = XchgHostSig{&x->hf->info[x->rank], &x->hf->ready[x->rank], x->h_pin[0], x->epoch};
        CK(cudaLaunchHostFunc(ctx->stream, xchg_host_sig_fn, &x->sig_pub));
    } else {
        CK(cudaMemcpyAsync(x->local + offsetof(XchgHeader, info), &x->h_pin[0], 8, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(x->local + offsetof(XchgHeader, ready), &x->h_pin[1], 8, cudaMemcpyHostToDevice, ctx->stream));
    }
    if (why) { cudaStreamSynchronize(ctx->stream); return fail(ctx, UMIGPU_ERR_UNSUPPORTED, "shard group: %s (%u unique UMIs, window holds %llu)", why, uh, (unsigned long long)x->ucap); }
    ctx->skip_bucket = hb;
    *u0_out = u0;
    XDBG("rank %d epoch %llu: published hot bucket %u (%u unique UMIs from unique id %u)", x->rank, (unsigned long long)x->epoch, hb, uh, u0);
    return UMIGPU_OK;
}

// every rank: import the hot bucket's arrays from the owner, search this rank's band of row tiles, put the edges
static int hot_band(umigpu_ctx *parent, const umigpu_hot *hot);
static int hot_collect(umigpu_ctx *ctx, u32 u0);

// child context of a rank: same configuration, its OWN stream and buffers (the band runs beside the rank's own neighbour
// search, from a second host thread: both are chains of small kernels and scalar read-backs that leave the device idle half
// of the time when run one after the other)
static int hot_child(umigpu_ctx *ctx) {
    if (ctx->hot) return UMIGPU_OK;
    umigpu_config c = ctx->cfg;
    c.stream = nullptr;
    c.flags &= ~UMIGPU_FLAG_LABELS;
    int rc = umigpu_create(&c, &ctx->hot);
    if (rc) { ctx->err = g_last_error; return rc; }
    ctx->hot->is_child = true;
    return UMIGPU_OK;
}

// Runs on its own host thread.  Errors are recorded in the CHILD context (parent->hot->err); only the child's stream, the
// exchange window's staging words [2], [3], [8..] and the polling stream are touched.
static int hot_band(umigpu_ctx *parent, const umigpu_hot *hot) {
    Xchg *x = parent->x;
    umigpu_ctx *ctx = parent->hot;                    // CK / LAUNCH / fail below act on the child
    umigpu_ctx *ch = ctx;
    CK(cudaSetDevice(parent->cfg.device));
    const char *ow = x->peer[(size_t)hot->owner];
    const unsigned long long epoch = x->epoch;
    int rc;
    unsigned long long info;
    if (x->hf) {
        rc = xchg_wait_host(x, ch, [&] { return x->hf->ready[hot->owner].load(std::memory_order_acquire) >= epoch; }, "the owner's hot bucket");
        if (rc) return rc;
        info = x->hf->info[hot->owner].load(std::memory_order_relaxed);
    } else {
        rc = xchg_poll(x, ch, ow + offsetof(XchgHeader, ready), 8, [&](const unsigned long long *b) { return b[0] >= epoch; }, "the owner's hot bucket");
        if (rc) return rc;
        rc = xchg_poll(x, ch, ow + offsetof(XchgHeader, info), 8, [](const unsigned long long *) { return true; }, "the owner's header");
        if (rc) return rc;
        info = x->h_pin[8];
    }
    if ((info >> 33) & 1ull) return fail(ch, UMIGPU_ERR_UNSUPPORTED, "shard group: the owner (rank %d) could not publish the hot bucket", hot->owner);
    const u32 uh = (u32)info;
    const bool has_n = ((info >> 32) & 1ull) != 0;
    const int narrow = (int)((info >> 34) & 1ull);
    rc = umigpu_reset(ch);
    if (rc) return rc;
    CK(cudaEventRecord(parent->ev[UMIGPU_STAGE_HOT_BAND][0], ch->stream));
    ch->ran = true;
    memset(&ch->ctr, 0, sizeof ch->ctr);
    ch->n_edges = 0;
    ch->skip_bucket = 0xffffffffu;
    ch->band = (u32)x->rank; ch->n_bands = (u32)x->n;
    memset(&ch->lay, 0, sizeof ch->lay);
    ch->lay.umi_len = (int)ch->cfg.umi_len; ch->lay.has_n = has_n ? 1 : 0;
    ch->lay.umi_bits = (has_n ? 3 : 2) * ch->lay.umi_len;
    ch->n_unique = uh; ch->n_buckets = uh ? 1 : 0;
    u64 c_edges = 0;
    if (uh > 1 && !(uh <= SMALL_BUCKET && x->rank != 0)) {        // a bucket of <= 32 UMIs is one warp's work: rank 0 takes it
        CK(ch->d_planes.reserve((size_t)uh * 8)); CK(ch->d_ucode.reserve((size_t)uh * 8)); CK(ch->d_freq.reserve((size_t)uh * 4));
        CK(ch->d_thr.reserve((size_t)uh * 4)); CK(ch->d_bstart.reserve(2 * 4)); CK(ch->d_ubkt.reserve((size_t)uh * 4));
        CK(ch->d_umi2.reserve((size_t)uh * 8));                   // staging of the codes as they come over the link
        if (has_n) CK(ch->d_nplane.reserve((size_t)uh * 4));
        cudaStream_t s = ch->stream;
        CK(cudaMemcpyAsync(ch->d_umi2.p, ow + x->off_code, (size_t)uh * (narrow ? 4 : 8), cudaMemcpyDefault, s));
        CK(cudaMemcpyAsync(ch->d_freq.p, ow + x->off_freq, (size_t)uh * 4, cudaMemcpyDefault, s));
        const umigpu_config &cf = ch->cfg;
        const int inf_thr = (cf.algo == UMIGPU_ALGO_CC || cf.algo == UMIGPU_ALGO_ADJ_UPSTREAM) ? 1 : 0;
        LAUNCH(hot_child_expand_kernel, grid_for(uh, 256), 256, uh, (const void *)ch->d_umi2.p, narrow, (const i32 *)ch->d_freq.p, cf.percentage, inf_thr,
               (int)cf.umi_len, has_n ? 1 : 0, ch->d_planes.as<uint2>(), ch->d_nplane.as<u32>(), ch->d_ucode.as<u64>(), ch->d_thr.as<i32>(),
               ch->d_bstart.as<u32>(), ch->d_ubkt.as<u32>());
        LAUNCH(bucket_stats_kernel, 1, 256, 1u, (const u32 *)nullptr, (const u32 *)ch->d_bstart.p, ch->d_sc.as<DevScalars>());
        ch->big_known = false;
        rc = stage_neighbours(ch, RUN_EDGES_ONLY);
        if (rc) return rc;
        c_edges = ch->n_edges;
    }
    // put: edges into this rank's region of the owner's window, then the slot (count, then epoch)
    unsigned long long count = c_edges;
    char *inbox = const_cast<char *>(ow) + x->off_inbox + (size_t)x->rank * x->region * 8;
    if (c_edges > x->region) count |= 1ull << 63;
    else if (c_edges) CK(cudaMemcpyAsync(inbox, ch->d_edges.p, (size_t)c_edges * 8, cudaMemcpyDefault, ch->stream));
    XDBG("rank %d epoch %llu: band done, %llu edges -> owner %d", x->rank, (unsigned long long)epoch, (unsigned long long)c_edges, hot->owner);
    x->h_pin[2] = count; x->h_pin[3] = epoch;
    if (x->hf) {
        x->sig_put = XchgHostSig{&x->hf->slot_count[hot->owner][x->rank], &x->hf->slot_epoch[hot->owner][x->rank], count, epoch};
        CK(cudaLaunchHostFunc(ch->stream, xchg_host_sig_fn, &x->sig_put));
    } else {
        char *slot = const_cast<char *>(ow) + offsetof(XchgHeader, slot) + (size_t)x->rank * sizeof(XchgSlot);
        CK(cudaMemcpyAsync(slot + offsetof(XchgSlot, count), &x->h_pin[2], 8, cudaMemcpyDefault, ch->stream));
        CK(cudaMemcpyAsync(slot + offsetof(XchgSlot, epoch), &x->h_pin[3], 8, cudaMemcpyDefault, ch->stream));
    }
    CK(cudaEventRecord(parent->ev[UMIGPU_STAGE_HOT_BAND][1], ch->stream));
    parent->ev_ok[UMIGPU_STAGE_HOT_BAND] = true;
    return UMIGPU_OK;
}

// owner: wait for every rank's edges, append them (hot-bucket-local ids + u0) to this context's edge list
static int hot_collect(umigpu_ctx *ctx, u32 u0) {
    Xchg *x = ctx->x;
    const unsigned long long epoch = x->epoch;
    const int nr = x->n;
    int rc;
    if (x->hf) {
        rc = xchg_wait_host(x, ctx, [&] { for (int r = 0; r < nr; r++) if (x->hf->slot_epoch[x->rank][r].load(std::memory_order_acquire) < epoch) return false; return true; },
                            "the edges of the other ranks");
        if (rc) return rc;
        for (int r = 0; r < nr; r++) x->h_pin[8 + 2 * r] = x->hf->slot_count[x->rank][r].load(std::memory_order_relaxed);
    } else {
        rc = xchg_poll(x, ctx, x->local + offsetof(XchgHeader, slot), (size_t)nr * sizeof(XchgSlot),
                       [&](const unsigned long long *b) { for (int r = 0; r < nr; r++) if (b[2 * r + 1] < epoch) return false; return true; },
                       "the edges of the other ranks");
        if (rc) return rc;
    }
    u64 cnt[XCHG_MAX_RANKS], total = 0;
    for (int r = 0; r < nr; r++) {
        const unsigned long long c = x->h_pin[8 + 2 * r];
        if (c >> 63) return fail(ctx, UMIGPU_ERR_UNSUPPORTED, "shard group: rank %d found %llu edges in its band of the hot bucket, its window region holds %llu "
                                 "(create the exchange window with a larger max_hot_edges)", r, (unsigned long long)(c & ~(1ull << 63)), (unsigned long long)x->region);
        cnt[r] = c; total += c;
    }
    const u64 own = ctx->n_edges;
    if (total) {
        CK(ctx->d_edges.reserve_keep((own + total) * sizeof(uint2), own * sizeof(uint2), ctx->stream));
        u64 at = own;
        for (int r = 0; r < nr; r++) {
            if (!cnt[r]) continue;
            LAUNCH(hot_append_kernel, grid_for(cnt[r], 256), 256, (const uint2 *)(x->local + x->off_inbox + (size_t)r * x->region * 8), cnt[r], u0,
                   ctx->d_edges.as<uint2>() + at);
            at += cnt[r];
        }
    }
    ctx->n_edges = own + total;
    ctx->ctr.n_edges = ctx->n_edges;
    return UMIGPU_OK;
}

extern "C" int umigpu_run_sharded(umigpu_ctx *ctx, const umigpu_hot *hot, int64_t key_lo, int64_t key_hi) {
    if (!ctx) return fail(nullptr, UMIGPU_ERR_ARG, "null context");
    Xchg *x = ctx->x;
    const bool want_labels = (ctx->cfg.flags & UMIGPU_FLAG_LABELS) != 0;
    const bool edges_possible = ctx->cfg.algo != UMIGPU_ALGO_ADJ && ctx->cfg.k > 0;
    const bool hot_on = hot && hot->present && edges_possible;
    if (hot_on && !x) return fail(ctx, UMIGPU_ERR_STATE, "umigpu_run_sharded: a hot bucket needs an exchange window (umigpu_xchg_create + attach)");
    if (hot_on && (hot->owner < 0 || hot->owner >= x->n)) return fail(ctx, UMIGPU_ERR_ARG, "hot bucket owner %d outside the group", hot->owner);
    int rc = run_begin(ctx);
    if (rc) return rc;
    const u64 n = ctx->n_reads;
    DevScalars *sc = ctx->d_sc.as<DevScalars>();
    STAGE_BEGIN(UMIGPU_STAGE_TOTAL);
    if (n) {
        LAUNCH(slice_range_kernel, (u32)std::min<u64>(grid_for(n, 256), (u64)ctx->num_sms * 16), 256, n, (const i32 *)ctx->d_tid.p, (const i64 *)ctx->d_pos.p, sc);
        rc = stage_group(ctx, want_labels, false);
        if (rc) return rc;
        // every (contig, position) of this slice must lie inside the slice's key range, or a bucket could straddle two devices
        if (ctx->h_sc->hot_pad) return fail(ctx, UMIGPU_ERR_UNSUPPORTED, "sharded run: positions beyond +-2^35");
        if (ctx->h_sc->key_lo < key_lo || (key_hi != INT64_MAX && ctx->h_sc->key_hi >= key_hi))
            return fail(ctx, UMIGPU_ERR_ARG, "sharded run: the slice is not range-partitioned by (contig, position): keys [%lld, %lld] outside [%lld, %lld) — "
                        "the input is not coordinate-sorted (use umigpu_shard_plan)", (long long)ctx->h_sc->key_lo, (long long)ctx->h_sc->key_hi,
                        (long long)key_lo, (long long)key_hi);
    }
    u32 u0 = 0;
    const bool owner = hot_on && x->rank == hot->owner;
    if (hot_on) {
        x->epoch++;
        rc = hot_child(ctx);
        if (rc) return rc;
        if (owner) { rc = hot_publish(ctx, hot, &u0); if (rc) return rc; }
        // this rank's band of the hot bucket beside its own neighbour search: two host threads, two streams
        int band_rc = UMIGPU_OK;
        std::thread band;
        try { band = std::thread([&] { band_rc = hot_band(ctx, hot); }); } catch (...) { band_rc = -1000; }
        if (band_rc == -1000) band_rc = hot_band(ctx, hot);            // no thread: band first, then the own search
        if (n) rc = stage_neighbours(ctx, RUN_FULL);
        if (band.joinable()) band.join();
        ctx->launches += ctx->hot->launches; ctx->hot->launches = 0;
        if (band_rc) return fail(ctx, band_rc, "%s", ctx->hot->err.c_str());
        if (rc) return rc;
    } else if (n) {
        rc = stage_neighbours(ctx, RUN_FULL);
        if (rc) return rc;
    }
    if (owner) { rc = hot_collect(ctx, u0); if (rc) return rc; }
    if (hot_on && ctx->hot) {          // this rank's share of the hot bucket's search
        const umigpu_counters &h = ctx->hot->ctr;
        ctx->ctr.pairs_evaluated += h.pairs_evaluated; ctx->ctr.n_tile_items += h.n_tile_items;
        ctx->ctr.n_tile_candidates += h.n_tile_candidates; ctx->ctr.n_block_pairs += h.n_block_pairs;
    }
    if (!n) { STAGE_END(UMIGPU_STAGE_TOTAL); return UMIGPU_OK; }
    rc = stage_cluster(ctx);
    if (rc) return rc;
    return stage_emit(ctx, want_labels);
}

// ---- plan for a coordinate-sorted stream ----
// Samples the stream (<= 65536 probes), estimates the big buckets from runs of equal (contig, position) in the sample, finds
// the hot bucket's exact extent by binary search, and cuts at bucket starts so that the modelled cost is balanced.  Cost
// model (ns, fitted to the measured stage times of C5 on one B200, profiles/): linear stages 0.11 per read; a bucket of r
// reads adds 1e-8 r^2 for its neighbour search and 0.1 r for its clustering; the hot bucket's search is shared by the
// group, so it only adds its clustering to the owner.  O(samples + log n) host work; the devices verify the cuts.
extern "C" int umigpu_shard_plan_sorted(uint64_t n, const int32_t *tid, const int64_t *unclipped_pos, const uint8_t *is_reverse,
                                        int32_t n_shards, uint64_t hot_min_reads, uint64_t *cuts, int64_t *cut_keys, umigpu_hot *hot,
                                        double *shard_cost) {
    if (n_shards < 1 || n_shards > XCHG_MAX_RANKS || !cuts || !cut_keys || (n && (!tid || !unclipped_pos || !is_reverse)))
        return fail(nullptr, UMIGPU_ERR_ARG, "umigpu_shard_plan_sorted: bad argument");
    if (hot) { hot->present = 0; hot->owner = 0; hot->read_index = 0; hot->reads_est = 0; }
    if (hot_min_reads == 0) hot_min_reads = 1u << 20;
    auto key = [&](u64 i) { return host_pos_key(tid[i], unclipped_pos[i]); };
    auto lower = [&](i64 k) { u64 lo = 0, hi = n; while (lo < hi) { u64 mid = lo + (hi - lo) / 2; if (key(mid) < k) lo = mid + 1; else hi = mid; } return lo; };
    for (int s = 0; s <= n_shards; s++) { cuts[s] = s == n_shards ? n : 0; cut_keys[s] = s == 0 ? INT64_MIN : INT64_MAX; }
    if (shard_cost) for (int s = 0; s < n_shards; s++) shard_cost[s] = 0.0;
    if (n == 0) return UMIGPU_OK;
    // C1H: the split bucket's owner sits on the group's critical path (every rank's band waits for its K1-K3, its clustering
    // waits for every rank's band), so the model charges it enough that it receives little besides the bucket itself
    const double A = 0.11, C2 = 1e-8, C1 = 0.10, C1H = 0.16;
    const u64 m = std::min<u64>(n, 65536);
    const double per = (double)n / (double)m;
    struct Unit { u64 first_sample; double reads, cost; i64 key; bool run; };
    std::vector<Unit> units; units.reserve((size_t)m);
    u64 hot_c = 0; i64 hot_key = 0;
    for (u64 j = 0; j < m;) {
        const i64 kj = key(j * n / m);
        u64 e = j + 1;
        while (e < m && key(e * n / m) == kj) e++;
        const u64 c = e - j;
        if (c >= 4) { units.push_back({j, c * per, 0.0, kj, true}); if (c > hot_c) { hot_c = c; hot_key = kj; } }
        else for (u64 q = j; q < e; q++) units.push_back({q, per, 0.0, kj, false});
        j = e;
    }
    bool have_hot = false;
    u64 hb = 0, he = 0;
    if (hot && n_shards >= 1 && hot_c && (double)hot_c * per >= (double)hot_min_reads) {
        hb = lower(hot_key); he = lower(hot_key + 1);
        if (he > hb) {
            u64 nrev = 0, probes = std::min<u64>(he - hb, 256);
            for (u64 q = 0; q < probes; q++) nrev += is_reverse[hb + q * (he - hb) / probes] ? 1 : 0;
            const u8 maj = nrev * 2 > probes ? 1 : 0;
            u64 r = hb;
            while (r < he && (is_reverse[r] ? 1 : 0) != maj) r++;
            if (r < he) {
                have_hot = true;
                hot->present = 1; hot->read_index = r;
                hot->reads_est = (u64)((double)(he - hb) * (maj ? (double)nrev : (double)(probes - nrev)) / (double)probes) + 1;
            }
        }
    }
    double total = 0.0;
    for (Unit &u : units) {
        u.cost = A * u.reads;
        if (u.run) u.cost += (have_hot && u.key == hot_key) ? C1H * u.reads : C2 * u.reads * u.reads + C1 * u.reads;
        total += u.cost;
    }
    // greedy prefix partition over the units; a cut lands on the first read of a unit's key
    size_t at = 0;
    double acc = 0.0;
    for (int s = 1; s < n_shards; s++) {
        // what is left is shared evenly among the shards that are left (a bucket that outweighs the fair share must not
        // starve the shards behind it)
        const double target = acc + (total - acc) / (double)(n_shards - s + 1);
        while (at < units.size() && acc + units[at].cost * 0.5 < target) { acc += units[at].cost; at++; }
        u64 c = at < units.size() ? lower(units[at].key) : n;
        if (c < cuts[s - 1]) c = cuts[s - 1];
        cuts[s] = c;
        cut_keys[s] = c < n ? key(c) : INT64_MAX;
        // everything up to the cut is accounted for (the unit that holds the cut key may have started in the previous units)
        while (at < units.size() && units[at].first_sample * n / m < c) { acc += units[at].cost; at++; }
    }
    if (have_hot) for (int s = 0; s < n_shards; s++) if (hb >= cuts[s] && hb < cuts[s + 1]) hot->owner = s;
    if (shard_cost) {
        size_t q = 0;
        for (int s = 0; s < n_shards; s++)
            for (; q < units.size() && units[q].first_sample * n / m < cuts[s + 1]; q++) shard_cost[s] += units[q].cost;
    }
    return UMIGPU_OK;
}

// ---- one process, several devices: a persistent group of contexts, one host thread per device per call ----
struct umigpu_group {
    umigpu_config cfg;
    std::vector<int> devices;
    std::vector<umigpu_ctx *> ctx;
    u64 ucap = 0, ecap = 0;
    std::atomic<int> abort{0};
};

extern "C" int umigpu_group_create(const umigpu_config *cfg, int32_t n_devices, const int32_t *device_ids, umigpu_group **out) {
    if (!cfg || !device_ids || !out || n_devices < 1 || n_devices > XCHG_MAX_RANKS) return fail(nullptr, UMIGPU_ERR_ARG, "umigpu_group_create: bad argument");
    *out = nullptr;
    umigpu_group *g = new (std::nothrow) umigpu_group();
    if (!g) return fail(nullptr, UMIGPU_ERR_NOMEM, "out of host memory");
    g->cfg = *cfg;
    for (int r = 0; r < n_devices; r++) {
        umigpu_config c = *cfg; c.device = device_ids[r]; c.stream = nullptr;
        umigpu_ctx *x = nullptr;
        int rc = umigpu_create(&c, &x);
        if (rc) { for (umigpu_ctx *p : g->ctx) umigpu_destroy(p); delete g; return rc; }
        g->ctx.push_back(x); g->devices.push_back(device_ids[r]);
    }
    *out = g;
    return UMIGPU_OK;
}

extern "C" void umigpu_group_destroy(umigpu_group *g) {
    if (!g) return;
    for (umigpu_ctx *p : g->ctx) umigpu_destroy(p);
    delete g;
}

extern "C" umigpu_ctx *umigpu_group_context(umigpu_group *g, int32_t rank) {
    return (g && rank >= 0 && rank < (int)g->ctx.size()) ? g->ctx[(size_t)rank] : nullptr;
}

static int group_windows(umigpu_group *g, u64 ucap, u64 ecap) {
    if (ucap <= g->ucap && ecap <= g->ecap && g->ucap) return UMIGPU_OK;
    ucap = std::max(ucap, g->ucap); ecap = std::max(ecap, g->ecap);
    const int nd = (int)g->ctx.size();
    for (int r = 0; r < nd; r++) { int rc = umigpu_xchg_create(g->ctx[(size_t)r], r, nd, ucap, ecap, nullptr); if (rc) return rc; }
    for (int r = 0; r < nd; r++) { int rc = umigpu_xchg_attach_local(g->ctx[(size_t)r], g->ctx.data()); if (rc) return rc; }
    for (int r = 0; r < nd; r++) g->ctx[(size_t)r]->x->abort = &g->abort;
    g->ucap = ucap; g->ecap = ecap;
    return UMIGPU_OK;
}

extern "C" int umigpu_group_dedup(umigpu_group *g, uint64_t n, const int32_t *tid, const int64_t *unclipped_pos, const uint8_t *is_reverse,
                                  const uint8_t *umi_ascii, const int32_t *score, uint64_t **kept, uint64_t *n_kept, umigpu_counters *counters,
                                  float *rank_ms /* nullable, n_devices */) {
    if (!g || !kept || !n_kept) return fail(nullptr, UMIGPU_ERR_ARG, "umigpu_group_dedup: bad argument");
    *kept = nullptr; *n_kept = 0;
    if (counters) memset(counters, 0, sizeof *counters);
    if (n == 0) return UMIGPU_OK;
    if (!tid || !unclipped_pos || !is_reverse || !umi_ascii) return fail(nullptr, UMIGPU_ERR_ARG, "umigpu_group_dedup: null input");
    const int nd = (int)g->ctx.size();
    const size_t L = g->cfg.umi_len;
    std::vector<u64> cuts((size_t)nd + 1);
    std::vector<i64> ckeys((size_t)nd + 1);
    umigpu_hot hot;
    const char *e_hm = getenv("UMIGPU_HOT_MIN_READS");
    int rc = umigpu_shard_plan_sorted(n, tid, unclipped_pos, is_reverse, nd, nd > 1 ? (e_hm ? strtoull(e_hm, nullptr, 10) : 0) : UINT64_MAX,
                                      cuts.data(), ckeys.data(), &hot, nullptr);
    if (rc) return rc;
    if (hot.present) {
        const char *e_ec = getenv("UMIGPU_XCHG_EDGES");
        rc = group_windows(g, hot.reads_est + 1024, e_ec ? strtoull(e_ec, nullptr, 10) : std::max<u64>((u64)1 << 22, 4 * hot.reads_est));
        if (rc) return rc;
    }
    // a call that failed half way leaves the ranks' epochs apart: flags compare with >=, so any common larger value resynchronises
    {
        u64 e = 0;
        for (umigpu_ctx *c : g->ctx) if (c->x) e = std::max(e, c->x->epoch);
        for (umigpu_ctx *c : g->ctx) if (c->x) c->x->epoch = e;
        g->abort.store(0);
    }
    std::vector<int> rcs((size_t)nd, 0);
    std::vector<std::string> errs((size_t)nd);
    std::vector<umigpu_result> res((size_t)nd);
    std::vector<std::thread> th;
    for (int r = 0; r < nd; r++) {
        th.emplace_back([&, r] {
            umigpu_ctx *c = g->ctx[(size_t)r];
            const u64 a = cuts[(size_t)r], m = cuts[(size_t)r + 1] - a;
            int q = umigpu_reset(c);
            if (!q && m) q = umigpu_push_reads(c, m, tid + a, unclipped_pos + a, is_reverse + a, umi_ascii + a * L, score ? score + a : nullptr, nullptr, a);
            if (!q) q = umigpu_run_sharded(c, &hot, ckeys[(size_t)r], ckeys[(size_t)r + 1]);
            if (!q) q = umigpu_fetch(c, &res[(size_t)r]);
            rcs[(size_t)r] = q;
            if (q) { errs[(size_t)r] = umigpu_last_error(c); g->abort.store(1); }
        });
    }
    for (auto &t : th) t.join();
    bool unsorted = false;
    for (int r = 0; r < nd; r++) if (rcs[(size_t)r] == UMIGPU_ERR_ARG && errs[(size_t)r].find("not range-partitioned") != std::string::npos) unsorted = true;
    if (unsorted) {
        // every rank validates its own slice before anything is exchanged; a rank that passed may be waiting for the owner: its
        // wait times out on its own.  Arbitrary order: gather per shard on the host instead.
        return dedup_sharded_lpt(&g->cfg, nd, g->devices.data(), n, tid, unclipped_pos, is_reverse, umi_ascii, score, kept, n_kept, counters);
    }
    u64 total = 0;
    for (int r = 0; r < nd; r++) { if (rcs[(size_t)r]) return fail(nullptr, rcs[(size_t)r], "shard %d failed: %s", r, errs[(size_t)r].c_str()); total += res[(size_t)r].n_kept; }
    u64 *out = (u64 *)malloc(std::max<u64>(total, 1) * sizeof(u64));
    if (!out) return fail(nullptr, UMIGPU_ERR_NOMEM, "out of host memory");
    u64 o = 0;
    for (int r = 0; r < nd; r++) {                 // slices are ascending index ranges: concatenation is the merge
        if (res[(size_t)r].n_kept) memcpy(out + o, res[(size_t)r].kept_read_index, res[(size_t)r].n_kept * 8);
        o += res[(size_t)r].n_kept;
    }
    *kept = out; *n_kept = total;
    for (int r = 0; r < nd; r++) {
        const umigpu_counters &p = res[(size_t)r].counters;
        if (counters) {
            counters->total_reads += p.total_reads; counters->n_buckets += p.n_buckets; counters->total_umis += p.total_umis;
            counters->max_umis = std::max(counters->max_umis, p.max_umis); counters->n_kept += p.n_kept;
            counters->unordered_pairs += p.unordered_pairs; counters->pairs_evaluated += p.pairs_evaluated; counters->n_edges += p.n_edges;
            counters->n_tile_items += p.n_tile_items; counters->n_tile_candidates += p.n_tile_candidates;
            counters->n_sweeps = std::max(counters->n_sweeps, p.n_sweeps); counters->n_block_pairs += p.n_block_pairs;
            counters->key_bits = std::max(counters->key_bits, p.key_bits);
        }
        if (rank_ms) { float ms = 0; umigpu_stage_ms(g->ctx[(size_t)r], UMIGPU_STAGE_TOTAL, &ms); rank_ms[r] = ms; }
    }
    return UMIGPU_OK;
}

extern "C" int umigpu_dedup_sharded(const umigpu_config *cfg, int32_t n_devices, const int32_t *device_ids, uint64_t n,
                                    const int32_t *tid, const int64_t *unclipped_pos, const uint8_t *is_reverse,
                                    const uint8_t *umi_ascii, const int32_t *score, uint64_t **kept, uint64_t *n_kept,
                                    umigpu_counters *counters) {
    if (!cfg || !device_ids || n_devices < 1 || !kept || !n_kept) return fail(nullptr, UMIGPU_ERR_ARG, "umigpu_dedup_sharded: bad argument");
    umigpu_group *g = nullptr;
    int rc = umigpu_group_create(cfg, n_devices, device_ids, &g);
    if (rc) return rc;
    rc = umigpu_group_dedup(g, n, tid, unclipped_pos, is_reverse, umi_ascii, score, kept, n_kept, counters, nullptr);
    umigpu_group_destroy(g);
    return rc;
}
