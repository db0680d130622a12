/*
 * umigpu.h — C ABI of libumigpu.so: the B200-native (sm_100a) UMI clustering hot path of
 * umi-collapse-rs, as a drop-in behind the reference's CLI / trait surface.
 *
 * Every entry point cites the reference interface it replaces (paths relative to the
 * reference repository, tkob-vh/umi-collapse-rs).  The reference-side binding a maintainer
 * would add (a Rust `extern "C"` crate + one new `impl DeduplicateInterface`) is shown in
 * INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++ or torch types cross this boundary;
 *   - every function returns 0 (UMIGPU_OK) or a negative UMIGPU_ERR_* code, never throws and
 *     never aborts; umigpu_last_error() returns a human-readable message.  (The reference
 *     panics on every error with panic="abort", Cargo.toml:19; the Rust shim turns a non-zero
 *     return into panic! to keep that behaviour.)
 *   - inputs are borrowed for the duration of the call; results are library-owned until
 *     umigpu_result_free / umigpu_destroy;
 *   - a context is bound to one CUDA device and is not thread-safe (the reference is
 *     single-threaded: &mut self on every trait method); a run may start short-lived host threads
 *     of its own (independent stages on a second stream: later multi-index passes, the band of a
 *     split hot bucket) and joins them before it returns;
 *   - there is no CPU fallback: without a CUDA device umigpu_create fails.
 */
#ifndef UMIGPU_H
#define UMIGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UMIGPU_OK                 0
#define UMIGPU_ERR_ARG           -1   /* bad argument / unsupported combination                       */
#define UMIGPU_ERR_CUDA          -2   /* CUDA runtime error (message has the cudaError string)        */
#define UMIGPU_ERR_BAD_BASE      -3   /* UMI byte not in {A,C,G,T,N}: reference panics, utils/mod.rs:78 */
#define UMIGPU_ERR_NOMEM         -4
#define UMIGPU_ERR_UNSUPPORTED   -5   /* key wider than 128 bits, umi_len > 32, > 2^32-2 reads ...    */
#define UMIGPU_ERR_STATE         -6   /* call order violated (e.g. fetch before run)                  */

/* --algo (src/cli.rs:33-36, dispatch src/main.rs:52-92) */
#define UMIGPU_ALGO_DIR           0   /* src/algo/directional.rs:58-91                                */
#define UMIGPU_ALGO_ADJ           1   /* src/algo/adjacency.rs:31-63 AS WRITTEN: remove_near(.,k,0)
                                         removes only the query, every unique UMI is kept   */
#define UMIGPU_ALGO_ADJ_UPSTREAM  2   /* opt-in: adjacency with max_freq = i32::MAX (upstream intent)  */
#define UMIGPU_ALGO_CC            3   /* help text only in the reference: directional with
                                         an infinite threshold = one representative per component     */

/* --merge (src/cli.rs:37-40, src/merge/mod.rs:18-51) */
#define UMIGPU_MERGE_ANY          0   /* AnyMerge: first read of a (bucket, UMI) is its representative */
#define UMIGPU_MERGE_AVGQUAL      1   /* AvgQualMerge: keep existing iff a.avg_qual >= b.avg_qual       */
#define UMIGPU_MERGE_MAPQUAL      2   /* MapQualMerge: same on MAPQ                                     */

/* flags */
#define UMIGPU_FLAG_LABELS        1u  /* also produce the per-read cluster root (ClusterTracker, --tag) */
#define UMIGPU_FLAG_NO_CULL       2u  /* evaluate every tile pair (disable exact prefix culling)        */
#define UMIGPU_FLAG_KERNEL_DIRECT 4u  /* use the direct XOR+popcount tile kernel instead of bit-sliced  */
#define UMIGPU_FLAG_NO_MULTI_INDEX 16u /* big buckets in one pass (disable the pigeonhole multi-index passes) */
#define UMIGPU_FLAG_PAIRED         32u /* --paired (src/cli.rs:49-54): umigpu_push_bam_records applies the paired-end
                                          filters of deduplicate_sam.rs:96-129 and the template length joins the
                                          bucket key (PairedAlignment, deduplicate_sam.rs:545-565)                */
#define UMIGPU_FLAG_REMOVE_UNPAIRED 64u /* --remove-unpaired (src/cli.rs:55-57), BAM feed with FLAG_PAIRED only    */
#define UMIGPU_FLAG_REMOVE_CHIMERIC 128u /* --remove-chimeric (src/cli.rs:58-60), BAM feed with FLAG_PAIRED only   */
#define UMIGPU_FLAG_KERNEL_TILES  8u  /* use the shared-memory tile form of the bit-sliced kernel instead of
                                         the block-pair list form                                        */

typedef struct umigpu_ctx umigpu_ctx;

/* Mirrors the fields of `struct Cli` (src/cli.rs:7-77) that reach the hot path:
 * Directional::new / Adjacency::new capture k and percentage (directional.rs:22-28). */
typedef struct umigpu_config {
    int32_t  k;            /* -k, src/cli.rs:18-19                                                  */
    float    percentage;   /* -p, src/cli.rs:25-26 (f32 like the reference)                         */
    int32_t  algo;         /* UMIGPU_ALGO_*                                                         */
    int32_t  merge;        /* UMIGPU_MERGE_*; selects nothing but whether `score` is honoured        */
    uint32_t umi_len;      /* -u / autodetected by the host (utils/read.rs:87-94); 1..32             */
    int32_t  device;       /* CUDA device ordinal                                                   */
    uint32_t flags;        /* UMIGPU_FLAG_*                                                         */
    uint32_t reserved;
    void    *stream;       /* cudaStream_t to launch on; NULL = a stream owned by the context       */
} umigpu_config;

/* The six end-of-run counters of src/deduplicate_sam.rs:243-267 (a13) plus path statistics. */
typedef struct umigpu_counters {
    uint64_t total_reads;       /* reads pushed (deduplicate_sam.rs:100; filters are applied by the host) */
    uint64_t n_buckets;         /* unique alignment positions, :198                                  */
    uint64_t total_umis;        /* sum over buckets of unique UMIs, :217                             */
    uint64_t max_umis;          /* max unique UMIs in a bucket, :218                                 */
    uint64_t n_kept;            /* reads after deduplicating, :219                                   */
    uint64_t unordered_pairs;   /* sum_b N_b (N_b - 1) / 2: what Naive must cover                    */
    uint64_t pairs_evaluated;   /* pair evaluations the device executed (after exact culling; blocks
                                   on the diagonal are evaluated in full)                             */
    uint64_t n_edges;           /* directed edges that passed the distance and count rule            */
    uint64_t n_tile_items;      /* tile-pair work items executed (survivors of the exact cull)       */
    uint64_t n_tile_candidates; /* tile pairs before culling                                         */
    uint64_t n_sweeps;          /* label-propagation sweeps                                          */
    uint64_t n_block_pairs;     /* (128 x 128) block pairs evaluated by the block-pair kernel          */
    uint64_t n_unmapped;        /* records dropped by the unmapped filter (deduplicate_sam.rs:102-108;
                                   BAM feed only; in paired mode also reads whose mate is unmapped,
                                   :118-121)                                                         */
    uint64_t n_unpaired;        /* paired BAM feed: reads without the paired flag, :111-116          */
    uint64_t n_chimeric;        /* paired BAM feed: tid != mtid, :123-128                            */
    uint64_t n_mates_skipped;   /* paired BAM feed: last-in-template records skipped before they are
                                   counted as input reads, :96-98                                    */
    uint64_t key_bits;          /* width of the sort key this batch needed (<= 64: one-word keys)    */
} umigpu_counters;

typedef struct umigpu_result {
    uint64_t         n_kept;
    const uint64_t  *kept_read_index;  /* ascending input order: what UcWriter::write receives
                                          (deduplicate_sam.rs:227-231), canonical order              */
    uint64_t         n_reads;
    const uint64_t  *read_cluster_root;/* FLAG_LABELS only: per pushed read (push order), the read index
                                          of the emitted representative of its cluster                */
    umigpu_counters  counters;
    const uint64_t  *read_umi_rep;     /* FLAG_LABELS only: per pushed read, the read index of the representative of its own
                                          (bucket, UMI) group — with read_cluster_root this is everything --tag needs
                                          (cluster id / cluster_size / same_umi, src/cli.rs:64-76)                    */
} umigpu_result;

/* stage ids for umigpu_stage_ms */
enum {
    UMIGPU_STAGE_PACK = 0,      /* K1 umi_pack (+H2D in the host entry)     */
    UMIGPU_STAGE_KEYS,          /* K1b build sort keys                      */
    UMIGPU_STAGE_SORT,          /* K2 bucket_group radix sort               */
    UMIGPU_STAGE_UNIQUE,        /* K3 umi_count_merge                       */
    UMIGPU_STAGE_WORKLIST,      /* bucket segmentation + tile work list     */
    UMIGPU_STAGE_NEIGHBOURS,    /* K5 hamming_neighbours                    */
    UMIGPU_STAGE_CLUSTER,       /* K6 cluster (label propagation)           */
    UMIGPU_STAGE_EMIT,          /* K7 emit_compact                          */
    UMIGPU_STAGE_TOTAL,         /* run() first launch -> last launch        */
    UMIGPU_STAGE_HOT_BAND,      /* sharded run: this rank's band of the hot bucket (import, K5, edges out) */
    UMIGPU_N_STAGES
};

const char *umigpu_version(void);
/* Initialises CUDA on `device` (driver start-up + primary context: seconds on a large host).  Optional: a host that
 * calls it from a helper thread before it starts reading its input hides that latency behind the I/O. */
int umigpu_device_init(int32_t device);
/* message of the last failing call on this thread (also valid when ctx == NULL) */
const char *umigpu_last_error(const umigpu_ctx *ctx);

/* Replaces DeduplicateSAM::new + Directional::new/Adjacency::new (deduplicate_sam.rs:46-68,
 * directional.rs:22-28, main.rs:52-92). */
int  umigpu_create(const umigpu_config *cfg, umigpu_ctx **out);
void umigpu_destroy(umigpu_ctx *ctx);
/* forget all pushed reads and results, keep device buffers (next batch / next bucket).  After a run that returned
 * UMIGPU_OK nothing is in flight and the call does not synchronise; after pushes that no successful run followed it waits for
 * their copies and kernels, so the buffer rule of umigpu_push_reads* ("until the next run / fetch / reset returns") holds. */
int  umigpu_reset(umigpu_ctx *ctx);

/*
 * HOT LOOP A, deduplicate_sam.rs:93-177, batched.  One call appends `n` reads (host SoA):
 *   tid            record.tid()                    (bucket key part, :137/:144)
 *   unclipped_pos  get_unclipped_pos(&record)      (utils/mod.rs:96-104)
 *   is_reverse     record.is_reverse()             (0/1)
 *   umi_ascii      n * umi_len bytes, get_umi()    (utils/read.rs:96-111) before to_bitset
 *   score          avg_qual (read.rs:56-63) or MAPQ (read.rs:77-79); may be NULL for MERGE_ANY
 *   weight         NULL = every read counts 1 (the reference); else per-read multiplicity
 *                  (used by umigpu_cluster_bucket to present pre-counted UMIs)
 *   first_read_index  index the caller gives to the first read of this chunk; kept indices are
 *                  reported in this numbering (chunks must be pushed in ascending index order)
 * The copy to the device and the UMI packing kernel are asynchronous on the context's stream:
 * PAGEABLE host buffers may be reused as soon as the call returns (CUDA stages them), PINNED host buffers
 * must stay unmodified until the next umigpu_run / umigpu_fetch / umigpu_reset returns.
 */
int umigpu_push_reads(umigpu_ctx *ctx, uint64_t n, const int32_t *tid, const int64_t *unclipped_pos,
                      const uint8_t *is_reverse, const uint8_t *umi_ascii, const int32_t *score,
                      const int32_t *weight, uint64_t first_read_index);
/* same, but every pointer is a DEVICE pointer valid on ctx's device (inputs already in HBM).  The first chunk of a batch
 * is borrowed, not copied, and the packing kernel that reads umi_ascii is only ENQUEUED on the context's stream: ALL the
 * arrays of the call, umi_ascii included, must stay valid and unchanged until the next umigpu_run / umigpu_fetch /
 * umigpu_reset of this context returns (the same rule as for pinned host buffers above).  The context's stream does not
 * synchronise with other streams: whatever produced the arrays must have completed (or been ordered before cfg.stream)
 * when this is called. */
int umigpu_push_reads_device(umigpu_ctx *ctx, uint64_t n, const int32_t *tid, const int64_t *unclipped_pos,
                             const uint8_t *is_reverse, const uint8_t *umi_ascii, const int32_t *score,
                             const int32_t *weight, uint64_t first_read_index);
/*
 * Paired-end form of umigpu_push_reads: `tlen` (record.insert_size(), deduplicate_sam.rs:138) is the fourth
 * field of the bucket key — Align::Paired(PairedAlignment{strand, coord, ref, tlen}), deduplicate_sam.rs:133-139
 * and :545-565.  The caller has already applied the paired-end filters of :96-129 (the BAM feed below applies
 * them itself under UMIGPU_FLAG_PAIRED).  Paired and unpaired pushes cannot be mixed in one batch.
 */
int umigpu_push_reads_paired(umigpu_ctx *ctx, uint64_t n, const int32_t *tid, const int64_t *unclipped_pos,
                             const uint8_t *is_reverse, const int64_t *tlen, const uint8_t *umi_ascii,
                             const int32_t *score, const int32_t *weight, uint64_t first_read_index);

/*
 * Compact host format of umigpu_push_reads, for hosts that already hold the UMI as a bit set (the reference builds one per
 * read, to_bitset utils/mod.rs:63-83): 14 (umi_len <= 16) or 18 bytes per read cross PCIe instead of 29.
 *   pos32     unclipped position as int32 (BAM positions are 31-bit)
 *   umi_2bit  2 bits per base, A0 C1 G2 T3, base 0 in the most significant used field (bits [2L-2, 2L)); element type
 *             uint32_t when umi_len <= 16, uint64_t otherwise
 *   n_mask    nullable; bit (L-1-b) set = base b is N (its 2-bit field is ignored)
 *   score8    nullable; avg_qual / MAPQ as uint8 (both are < 256 by construction)
 * Same semantics as umigpu_push_reads otherwise (single-end; host pointers).  Bits set beyond 2*umi_len / umi_len are reported
 * like an unknown UMI byte (UMIGPU_ERR_BAD_BASE at run).
 */
int umigpu_push_reads_packed(umigpu_ctx *ctx, uint64_t n, const int32_t *tid, const int32_t *pos32, const uint8_t *is_reverse,
                             const void *umi_2bit, const uint32_t *n_mask, const uint8_t *score8, uint64_t first_read_index);

/*
 * Host feed on the device: `records` holds raw BAM alignment records exactly as they
 * appear in the BGZF-inflated stream (int32 block_size, then block_size bytes), `offsets[i]` is the byte offset
 * of record i's block_size field relative to `records` and offsets[n] the end of the last one
 * (umigpu_bam_record_offsets finds them).  Per record the device evaluates get_unclipped_pos
 * (utils/mod.rs:96-104), UcSAMRead::get_umi after the first `umi_sep` (utils/read.rs:96-111), the score of the
 * configured merge (avg_qual read.rs:56-63 / MAPQ :77-79) and the unmapped filter (deduplicate_sam.rs:102-108);
 * survivors are appended like umigpu_push_reads.  Kept indices are first_read_index + record number within this
 * call.  Host pointers.
 * With UMIGPU_FLAG_PAIRED in the context's flags the device also applies, in the reference's order
 * (deduplicate_sam.rs:96-129): paired last-in-template records are skipped without being counted as input
 * reads; after the unmapped filter, unpaired reads are counted (and dropped under FLAG_REMOVE_UNPAIRED), reads
 * whose mate is unmapped count as unmapped and are dropped, chimeric reads (tid != mtid) are counted (and dropped
 * under FLAG_REMOVE_CHIMERIC); the survivors' insert size joins the bucket key.  *n_unmapped then receives this
 * call's reference `unmapped` increment (:104 + :119).  The mates of the kept reads are selected by the caller
 * (UcWriter::write_reversed, deduplicate_sam.rs:409-462 — see host/umicollapse_gpu.cpp).
 */
int umigpu_push_bam_records(umigpu_ctx *ctx, uint64_t n, const uint8_t *records, const uint64_t *offsets,
                            uint8_t umi_sep, uint64_t first_read_index, uint64_t *n_unmapped);
/* Pure host helper: walks the block_size fields of buf[0..len) and writes up to max_records offsets (+ the end
 * offset); *consumed = bytes covered by whole records (a trailing partial record is left for the next call). */
int umigpu_bam_record_offsets(const uint8_t *buf, uint64_t len, uint64_t *offsets, uint64_t max_records,
                              uint64_t *n_records, uint64_t *consumed);

/* HOT LOOP B, deduplicate_sam.rs:207-233 over all buckets at once: group, count, merge, neighbour
 * search, cluster, compact.  Asynchronous apart from a few scalar read-backs. */
int umigpu_run(umigpu_ctx *ctx);
/* Wait for run() and copy the kept indices (and labels) to host memory owned by the context. */
int umigpu_fetch(umigpu_ctx *ctx, umigpu_result *out);
/* run + fetch */
int umigpu_finish(umigpu_ctx *ctx, umigpu_result *out);
/* counters of the last run without copying the kept list */
int umigpu_get_counters(umigpu_ctx *ctx, umigpu_counters *out);

/*
 * Algorithm::apply-shaped entry (src/algo/mod.rs:13-20, called deduplicate_sam.rs:211-213): one
 * bucket of `n` distinct UMIs (ASCII, n*umi_len) with their frequencies.
 *   keep[i]  = 1 iff UMI i's representative is emitted
 *   label[i] = index of the emitted UMI whose cluster contains UMI i (what ClusterTracker records)
 * Visit order = (freq descending, UMI ascending A<C<G<T<N) — one of the reference's possible orders.
 */
int umigpu_cluster_bucket(umigpu_ctx *ctx, uint64_t n, const uint8_t *umi_ascii, const int32_t *freq,
                          uint8_t *keep, int32_t *label);

/*
 * DataStruct-shaped entries (src/data/mod.rs:11-17, src/data/naive.rs:22-44).
 * umigpu_remove_near: out[i] = 1 iff Naive::remove_near(query, k, max_freq) would remove UMI i:
 *   dist <= k && (dist == 0 || freq[i] <= max_freq).
 * umigpu_neighbours: the whole bucket at once as CSR adjacency over the input indices:
 *   col[row_ptr[i] .. row_ptr[i+1]) = ascending j != i with dist(i,j) <= k and, when apply_rule != 0,
 *   freq[j] <= trunc_i32(f32(percentage) * f32(freq[i] + 1)) (directional.rs:38).  col_capacity is the
 *   size of col; *n_edges always receives the true edge count (call again with a larger col if it
 *   exceeds col_capacity; UMIGPU_ERR_ARG is returned in that case).
 */
int umigpu_remove_near(umigpu_ctx *ctx, uint64_t n, const uint8_t *umi_ascii, const int32_t *freq,
                       const uint8_t *query, int32_t k, int32_t max_freq, uint8_t *out);
int umigpu_neighbours(umigpu_ctx *ctx, uint64_t n, const uint8_t *umi_ascii, const int32_t *freq,
                      int32_t apply_rule, uint64_t *row_ptr, uint32_t *col, uint64_t col_capacity,
                      uint64_t *n_edges);

/* UcSAMRead::new score, utils/read.rs:56-63: per read trunc_i32(f32 sum of qual bytes / f32 len).
 * qual = concatenated phred bytes, offsets[n+1] (host pointers); out = n scores. */
int umigpu_avg_qual(umigpu_ctx *ctx, uint64_t n, const uint8_t *qual, const uint64_t *offsets, int32_t *out);

/* device time of a stage of the last run (CUDA events on the context's stream), milliseconds */
int umigpu_stage_ms(umigpu_ctx *ctx, int stage, float *ms);
/* number of kernels this context has launched since create / since the last call with reset != 0 */
uint64_t umigpu_launch_count(umigpu_ctx *ctx, int reset);
void umigpu_result_free(umigpu_ctx *ctx);

/*
 * ---- Several devices, ONE dataset ------------------------------------------------------------
 * Buckets never interact (deduplicate_sam.rs:207-213 handles each map entry alone), so the path shards by bucket
 * with no collective.  A coordinate-sorted stream is cut at bucket starts into one CONTIGUOUS slice per device
 * (umigpu_shard_plan_sorted): slice r is reads [cuts[r], cuts[r+1]) — a plain pointer offset into the host arrays,
 * one H2D range per device — and because the slices are ascending index ranges the merged kept list is the
 * concatenation of the ranks' kept lists.  The one exchange step: a bucket that outweighs a device's fair share
 * (the hot locus of a skewed run) stays on one device up to the unique/count stage, then its neighbour search is
 * split over all devices of the group through peer-addressable EXCHANGE WINDOWS (cudaMemcpyAsync over NVLink /
 * NVSwitch: unique-UMI arrays out, edges back), and the owner clusters it.
 */
typedef struct umigpu_hot {
    int32_t  present;      /* 0 = no bucket is split                                                            */
    int32_t  owner;        /* rank whose slice holds the bucket                                                 */
    uint64_t read_index;   /* caller's index of ONE read of the bucket (identifies it on the owner)              */
    uint64_t reads_est;    /* estimated reads in the bucket (sizes the exchange windows)                         */
} umigpu_hot;

/* (contig, position) as one ordered int64: the key space of the cuts.  Valid for |unclipped_pos| < 2^35. */
int64_t umigpu_pos_key(int32_t tid, int64_t unclipped_pos);

/*
 * Plan for a stream sorted by (tid, unclipped_pos) (strand and UMI in any order).  Pure host helper, O(65536 probes +
 * log n): cuts[0..n_shards] ascending read indices with cuts[0] = 0, cuts[n_shards] = n, every cut on the first read of
 * a (tid, pos); cut_keys[s] = umigpu_pos_key of read cuts[s] (INT64_MIN / INT64_MAX at the ends): slice s must only hold
 * keys in [cut_keys[s], cut_keys[s+1]) — umigpu_run_sharded verifies that on the device, so an unsorted stream is an
 * error, never a wrong answer.  Slices are balanced on a cost model fitted to the measured stages (linear per read +
 * quadratic search and linear clustering per big bucket).  *hot describes the bucket to split, if one holds at least
 * hot_min_reads reads (0 = default 2^20, UINT64_MAX = never split).  shard_cost (nullable) receives the modelled cost.
 */
int umigpu_shard_plan_sorted(uint64_t n, const int32_t *tid, const int64_t *unclipped_pos, const uint8_t *is_reverse,
                             int32_t n_shards, uint64_t hot_min_reads, uint64_t *cuts /* n_shards + 1 */,
                             int64_t *cut_keys /* n_shards + 1 */, umigpu_hot *hot, double *shard_cost /* n_shards */);

/*
 * Exchange window of one rank: device memory holding the hot bucket's unique-UMI arrays (when this rank owns it) and
 * one inbox region per rank for edges.  Every rank of a group creates its window with the SAME sizes, then attaches to
 * the others': ranks in different processes exchange the 64-byte CUDA IPC handle (ipc_handle_out) by any means and call
 * umigpu_xchg_attach_ipc with all n_ranks handles in rank order; ranks inside one process call umigpu_xchg_attach_local
 * with the group's contexts (peer access is enabled as needed; the same device may appear more than once).
 */
int umigpu_xchg_create(umigpu_ctx *ctx, int32_t rank, int32_t n_ranks, uint64_t max_hot_uniques, uint64_t max_hot_edges,
                       uint8_t *ipc_handle_out /* 64 bytes, nullable */);
int umigpu_xchg_attach_ipc(umigpu_ctx *ctx, const uint8_t *handles /* n_ranks x 64 bytes */);
int umigpu_xchg_attach_local(umigpu_ctx *ctx, umigpu_ctx *const *group /* n_ranks contexts, rank order */);

/*
 * umigpu_run for one rank of a group: the reads pushed into ctx are slice `rank` of the plan (first_read_index =
 * cuts[rank]); key_lo / key_hi = cut_keys[rank] / cut_keys[rank + 1].  With hot->present every rank of the group must
 * make this call (a rank with an empty slice too): the owner publishes the bucket, every rank searches the row tiles
 * ti with ti % n_ranks == rank, the owner gathers the edges and clusters.  hot may be NULL (or present = 0): the ranks
 * are then fully independent and no window is needed.  Results through umigpu_fetch as usual.
 */
int umigpu_run_sharded(umigpu_ctx *ctx, const umigpu_hot *hot, int64_t key_lo, int64_t key_hi);

/*
 * One process, several devices (what the Rust host calls): a persistent group of contexts.  umigpu_group_dedup plans
 * (umigpu_shard_plan_sorted), pushes every device its slice straight from the caller's arrays on its own host thread,
 * runs umigpu_run_sharded and concatenates.  *kept receives a malloc'ed array of *n_kept ascending read indices (free
 * with umigpu_free); counters are summed over ranks (max_umis, n_sweeps, key_bits: maximum); rank_ms (nullable) the
 * device time of every rank's run.  Input that is not coordinate-sorted is detected by the devices and re-run through
 * the hash plan below (gather per shard on the host).  device_ids may repeat a device.
 */
typedef struct umigpu_group umigpu_group;
int  umigpu_group_create(const umigpu_config *cfg, int32_t n_devices, const int32_t *device_ids, umigpu_group **out);
void umigpu_group_destroy(umigpu_group *g);
umigpu_ctx *umigpu_group_context(umigpu_group *g, int32_t rank);
int  umigpu_group_dedup(umigpu_group *g, uint64_t n, const int32_t *tid, const int64_t *unclipped_pos, const uint8_t *is_reverse,
                        const uint8_t *umi_ascii, const int32_t *score, uint64_t **kept, uint64_t *n_kept,
                        umigpu_counters *counters, float *rank_ms);
/* create + dedup + destroy.  cfg->device and cfg->stream are ignored. */
int umigpu_dedup_sharded(const umigpu_config *cfg, int32_t n_devices, const int32_t *device_ids, uint64_t n,
                         const int32_t *tid, const int64_t *unclipped_pos, const uint8_t *is_reverse, const uint8_t *umi_ascii,
                         const int32_t *score, uint64_t **kept, uint64_t *n_kept, umigpu_counters *counters);
void umigpu_free(void *p);

/*
 * Hash plan for reads in ARBITRARY order: longest-processing-time-first over per-bucket cost (reads^2 + 64 reads),
 * shard_of_read[n] receives the shard id; the caller gathers each shard's reads.  Pure host helper, one hash probe per
 * read — use umigpu_shard_plan_sorted whenever the stream is coordinate-sorted (a BAM is).
 */
int umigpu_shard_plan(uint64_t n, const int32_t *tid, const int64_t *unclipped_pos, const uint8_t *is_reverse,
                      int32_t n_shards, int32_t *shard_of_read, uint64_t *shard_cost /* n_shards */);

/* integer-pipe microbenchmark used as the roofline denominator of the neighbour search:
 * ops/s of a dependent-free LOP3 stream and of POPC on the context's device. */
int umigpu_int_peak(umigpu_ctx *ctx, double *lop3_ops_per_s, double *popc_ops_per_s);

#ifdef __cplusplus
}
#endif
#endif /* UMIGPU_H */
