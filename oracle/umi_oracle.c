/*
 * umi_oracle.c — CPU restatement of the umi-collapse-rs UMI clustering hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (umi-collapse-rs_b200/) may
 * include, link, load or call this file.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it, and only as the checker or as
 * the timed CPU baseline.
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures
 * (SURVEY.md §4), and no Rust toolchain exists here, so this restatement could not be
 * checked against reference output.  It is pinned instead against an independent literal
 * Python restatement (oracle/ref_literal.py), hand-derived known-answer buckets and the
 * order-independent invariants of the reference (tests/).
 *
 * Every function cites the reference lines it restates (paths relative to /root/reference).
 *
 * Canonicalisation (the reference itself is non-deterministic):
 *   - HashMap iteration order of the UMIs of a bucket is fixed to ascending UMI string
 *     order with A < C < G < T < N; the stable frequency-descending sort of
 *     src/algo/directional.rs:67-72 then breaks ties in that order;
 *   - survivors are reported in input order (ascending read index).
 * Every output of this oracle is one of the outputs the reference can produce.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <limits.h>

#define ORACLE_OK            0
#define ORACLE_ERR_BAD_BASE -1   /* reference: panic!("Unknown character in UMI sequence") utils/mod.rs:78 */
#define ORACLE_ERR_ARG      -2
#define ORACLE_ERR_NOMEM    -3

#define ALGO_DIR          0      /* src/algo/directional.rs */
#define ALGO_ADJ_REF      1      /* src/algo/adjacency.rs as written (max_freq = 0) */
#define ALGO_ADJ_UPSTREAM 2      /* adjacency with max_freq = i32::MAX (upstream intent; opt-in) */
#define ALGO_CC           3      /* no reference behaviour: directional with threshold = i32::MAX */

#define MERGE_ANY     0          /* src/merge/mod.rs:18-23  */
#define MERGE_AVGQUAL 1          /* src/merge/mod.rs:33-37  */
#define MERGE_MAPQUAL 2          /* src/merge/mod.rs:47-51  */

#define MAXW 4                   /* 64-bit words per BitSet: ceil(3*L/64), L <= 85 */

/* ---- src/utils/read.rs:13-31: ENCODING_DIST = 2, ENCODING_LENGTH = 3, encoding table ---- */
#define ENCODING_DIST   2
#define ENCODING_LENGTH 3

static int encoding_of(uint8_t c) {
    switch (c) {             /* utils/read.rs:22-31 */
    case 'A': return 0;      /* 0b000 */
    case 'T': return 5;      /* 0b101 */
    case 'C': return 6;      /* 0b110 */
    case 'G': return 3;      /* 0b011 */
    case 'N': return 4;      /* 0b100 UNDETERMINED */
    default:  return -1;
    }
}

/* canonical order rank of a base: A < C < G < T < N (this repo's tie-break, not the reference's) */
static int canon_rank(uint8_t c) {
    switch (c) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; default: return 4; }
}

typedef struct {
    int64_t bits[MAXW];      /* utils/bitset.rs:10 */
    int64_t nbits[MAXW];     /* utils/bitset.rs:11 (None == all zero) */
} bitset_t;

static int nwords_for(int umi_len) {   /* utils/bitset.rs:17-18 */
    int length = umi_len * ENCODING_LENGTH;
    return length / 64 + (length % 64 == 0 ? 0 : 1);
}

/* utils/mod.rs:63-83 to_bitset; char_set utils/mod.rs:38-43; char_set_n_bit :45-50 */
static int to_bitset(const uint8_t *s, int umi_len, bitset_t *out) {
    memset(out, 0, sizeof(*out));
    for (int i = 0; i < umi_len; i++) {
        int enc = encoding_of(s[i]);
        if (enc < 0) return ORACLE_ERR_BAD_BASE;
        for (int b = 0; b < ENCODING_LENGTH; b++) {
            int idx = i * ENCODING_LENGTH + b;
            if (enc & (1 << b)) out->bits[idx / 64] |= (int64_t)((uint64_t)1 << (idx % 64));
            if (s[i] == 'N')    out->nbits[idx / 64] |= (int64_t)((uint64_t)1 << (idx % 64));
        }
    }
    return ORACLE_OK;
}

/* utils/bitset.rs:77-91 bit_count_xor */
static int bit_count_xor(const bitset_t *a, const bitset_t *b, int nw) {
    int res = 0;
    for (int i = 0; i < nw; i++) {
        uint64_t x = (uint64_t)(a->nbits[i] ^ b->nbits[i]);
        res += __builtin_popcountll(x | (uint64_t)(a->bits[i] ^ b->bits[i]))
             - __builtin_popcountll(x) / ENCODING_LENGTH;
    }
    return res;
}

/* utils/mod.rs:24-26 umi_dist */
static int umi_dist(const bitset_t *a, const bitset_t *b, int nw) {
    return bit_count_xor(a, b, nw) / ENCODING_DIST;
}

/* algo/directional.rs:38 threshold = (p * (freq + 1) as f32) as i32, f32 arithmetic, saturating cast */
static int32_t dir_threshold(float p, int32_t freq) {
    volatile float t = p * (float)(freq + 1);
    if (t != t) return 0;
    if (t >= 2147483648.0f) return INT32_MAX;
    if (t <= -2147483648.0f) return INT32_MIN;
    return (int32_t)t;
}

int oracle_umi_dist(const uint8_t *a, const uint8_t *b, int umi_len) {
    bitset_t x, y;
    if (umi_len <= 0 || nwords_for(umi_len) > MAXW) return ORACLE_ERR_ARG;
    if (to_bitset(a, umi_len, &x) || to_bitset(b, umi_len, &y)) return ORACLE_ERR_BAD_BASE;
    return umi_dist(&x, &y, nwords_for(umi_len));
}

int32_t oracle_dir_threshold(float p, int32_t freq) { return dir_threshold(p, freq); }

/* ---------------------------------------------------------------------------------------
 * Per-bucket clustering: Algorithm::apply (algo/mod.rs:13-20) over Naive (data/naive.rs).
 * -------------------------------------------------------------------------------------*/
typedef struct {
    bitset_t bs;
    const uint8_t *ascii;
    int32_t freq;
    int64_t rep;            /* read index of the representative */
    int32_t score;
} uentry_t;

static _Thread_local int g_umi_len_cmp;   /* comparator context (per thread: the all-core variant sorts buckets concurrently) */

static int cmp_canon(const void *pa, const void *pb) {
    const uentry_t *a = (const uentry_t *)pa, *b = (const uentry_t *)pb;
    for (int i = 0; i < g_umi_len_cmp; i++) {
        int ra = canon_rank(a->ascii[i]), rb = canon_rank(b->ascii[i]);
        if (ra != rb) return ra < rb ? -1 : 1;
    }
    return 0;
}

/* Naive::remove_near, data/naive.rs:26-40.  `remaining` is the compact list of indices still in
 * the map; it is scanned in full on every call exactly like HashMap::retain.  Removed indices
 * are appended to `near` (the returned HashSet); returns how many. */
static int64_t naive_remove_near(const uentry_t *e, int nw, int32_t *remaining, int64_t *n_remaining,
                                 uint8_t *present, int32_t q, int k, int32_t max_freq,
                                 int32_t *near, uint64_t *dist_calls) {
    int64_t m = *n_remaining, w = 0, nn = 0;
    for (int64_t i = 0; i < m; i++) {
        int32_t o = remaining[i];
        int dist = umi_dist(&e[q].bs, &e[o].bs, nw);                 /* naive.rs:30 */
        if (dist <= k && (dist == 0 || e[o].freq <= max_freq)) {     /* naive.rs:31 */
            near[nn++] = o; present[o] = 0;                          /* naive.rs:32-33 */
        } else {
            remaining[w++] = o;                                      /* naive.rs:35 */
        }
    }
    *dist_calls += (uint64_t)m;
    *n_remaining = w;
    return nn;
}

/* Clusters one bucket.  entries[] must already be in canonical (UMI ascending) order.
 * keep[i]  = 1 iff entry i's representative read is emitted (directional.rs:81-86, adjacency.rs:49-59)
 * label[i] = index of the emitted entry whose visit removed entry i (its cluster root).
 * For ALGO_ADJ_REF only the query itself is ever removed. */
static int cluster_entries(const uentry_t *e, int64_t n, int umi_len, int algo, int k, float p,
                           uint8_t *keep, int32_t *label, uint64_t *dist_calls) {
    int nw = nwords_for(umi_len);
    if (n == 0) return ORACLE_OK;
    int32_t *order = (int32_t *)malloc(sizeof(int32_t) * n);
    int32_t *remaining = (int32_t *)malloc(sizeof(int32_t) * n);
    uint8_t *present = (uint8_t *)malloc(n);
    int32_t *nearbuf = (int32_t *)malloc(sizeof(int32_t) * n);     /* all near-sets, total <= n */
    /* DFS frames (directional.rs:30-54 recursion made explicit; same visiting order) */
    int64_t *fr_begin = (int64_t *)malloc(sizeof(int64_t) * (n + 1));
    int64_t *fr_end = (int64_t *)malloc(sizeof(int64_t) * (n + 1));
    int64_t *fr_pos = (int64_t *)malloc(sizeof(int64_t) * (n + 1));
    int32_t *fr_start = (int32_t *)malloc(sizeof(int32_t) * (n + 1));
    if (!order || !remaining || !present || !nearbuf || !fr_begin || !fr_end || !fr_pos || !fr_start) {
        free(order); free(remaining); free(present); free(nearbuf);
        free(fr_begin); free(fr_end); free(fr_pos); free(fr_start);
        return ORACLE_ERR_NOMEM;
    }
    /* directional.rs:67-72 / adjacency.rs:40-45: stable sort by freq descending.  Insertion of
     * canonical-ordered entries into a counting pass keeps it stable and O(n log n)-free. */
    {
        /* stable merge sort on indices by freq desc */
        int32_t *tmp = (int32_t *)malloc(sizeof(int32_t) * n);
        if (!tmp) return ORACLE_ERR_NOMEM;
        for (int64_t i = 0; i < n; i++) order[i] = (int32_t)i;
        for (int64_t width = 1; width < n; width *= 2) {
            for (int64_t lo = 0; lo < n; lo += 2 * width) {
                int64_t mid = lo + width < n ? lo + width : n, hi = lo + 2 * width < n ? lo + 2 * width : n;
                int64_t a = lo, b = mid, o = lo;
                while (a < mid && b < hi) {
                    if (e[order[b]].freq > e[order[a]].freq) tmp[o++] = order[b++];   /* strictly greater moves ahead */
                    else tmp[o++] = order[a++];
                }
                while (a < mid) tmp[o++] = order[a++];
                while (b < hi) tmp[o++] = order[b++];
            }
            memcpy(order, tmp, sizeof(int32_t) * n);
        }
        free(tmp);
    }
    /* Naive::new, naive.rs:22-24 */
    for (int64_t i = 0; i < n; i++) { remaining[i] = (int32_t)i; present[i] = 1; keep[i] = 0; label[i] = -1; }
    int64_t n_remaining = n;

    for (int64_t oi = 0; oi < n; oi++) {                            /* directional.rs:78 / adjacency.rs:47 */
        int32_t root = order[oi];
        if (!present[root]) continue;                              /* data.contains, naive.rs:42-44 */
        if (algo == ALGO_ADJ_REF || algo == ALGO_ADJ_UPSTREAM) {
            int32_t max_freq = algo == ALGO_ADJ_REF ? 0 : INT32_MAX; /* adjacency.rs:56 passes 0 */
            int64_t nn = naive_remove_near(e, nw, remaining, &n_remaining, present, root, k, max_freq,
                                           nearbuf, dist_calls);
            for (int64_t j = 0; j < nn; j++) label[nearbuf[j]] = root;
        } else {
            /* Directional::visit_and_remove, directional.rs:30-54 */
            int64_t sp = 0, near_top = 0;
            int32_t thr = algo == ALGO_CC ? INT32_MAX : dir_threshold(p, e[root].freq);
            int64_t nn = naive_remove_near(e, nw, remaining, &n_remaining, present, root, k, thr,
                                           nearbuf + near_top, dist_calls);
            fr_begin[sp] = near_top; fr_end[sp] = near_top + nn; fr_pos[sp] = near_top; fr_start[sp] = root;
            near_top += nn; sp++;
            while (sp > 0) {
                int64_t f = sp - 1;
                if (fr_pos[f] == fr_end[f]) { sp--; continue; }
                int32_t v = nearbuf[fr_pos[f]++];
                label[v] = root;
                if (v == fr_start[f]) continue;                     /* directional.rs:48-50 */
                int32_t vthr = algo == ALGO_CC ? INT32_MAX : dir_threshold(p, e[v].freq);
                int64_t vn = naive_remove_near(e, nw, remaining, &n_remaining, present, v, k, vthr,
                                               nearbuf + near_top, dist_calls);
                fr_begin[sp] = near_top; fr_end[sp] = near_top + vn; fr_pos[sp] = near_top; fr_start[sp] = v;
                near_top += vn; sp++;
            }
        }
        keep[root] = 1;                                            /* res.push(&read_freq.read) */
        label[root] = root;
    }
    free(order); free(remaining); free(present); free(nearbuf);
    free(fr_begin); free(fr_end); free(fr_pos); free(fr_start);
    return ORACLE_OK;
}

/* Algorithm::apply-shaped entry: n unique UMIs of one bucket (ASCII, n*umi_len bytes, any order,
 * must be distinct) with their frequencies.  keep/label are indexed like the input. */
int oracle_cluster_bucket(int64_t n, const uint8_t *umi_ascii, int umi_len, const int32_t *freq,
                          int algo, int k, float percentage, uint8_t *keep, int32_t *label,
                          uint64_t *dist_calls) {
    if (n < 0 || umi_len <= 0 || nwords_for(umi_len) > MAXW) return ORACLE_ERR_ARG;
    uentry_t *e = (uentry_t *)malloc(sizeof(uentry_t) * (n ? n : 1));
    if (!e) return ORACLE_ERR_NOMEM;
    for (int64_t i = 0; i < n; i++) {
        e[i].ascii = umi_ascii + i * umi_len;
        if (to_bitset(e[i].ascii, umi_len, &e[i].bs)) { free(e); return ORACLE_ERR_BAD_BASE; }
        e[i].freq = freq[i]; e[i].rep = i; e[i].score = 0;
    }
    g_umi_len_cmp = umi_len;
    qsort(e, n, sizeof(uentry_t), cmp_canon);
    uint8_t *k2 = (uint8_t *)malloc(n ? n : 1);
    int32_t *l2 = (int32_t *)malloc(sizeof(int32_t) * (n ? n : 1));
    uint64_t dc = 0;
    int rc = cluster_entries(e, n, umi_len, algo, k, percentage, k2, l2, &dc);
    if (rc == ORACLE_OK) {
        for (int64_t i = 0; i < n; i++) {
            keep[e[i].rep] = k2[i];
            label[e[i].rep] = l2[i] < 0 ? -1 : (int32_t)e[l2[i]].rep;
        }
    }
    if (dist_calls) *dist_calls = dc;
    free(e); free(k2); free(l2);
    return rc;
}

/* DataStruct::remove_near-shaped entry (data/mod.rs:11-17, naive.rs:26-40): one query against a
 * set of n UMIs with frequencies; out[i] = 1 iff UMI i would be removed/returned. */
int oracle_remove_near(int64_t n, const uint8_t *umi_ascii, int umi_len, const int32_t *freq,
                       const uint8_t *query, int k, int32_t max_freq, uint8_t *out) {
    if (umi_len <= 0 || nwords_for(umi_len) > MAXW) return ORACLE_ERR_ARG;
    int nw = nwords_for(umi_len);
    bitset_t q, o;
    if (to_bitset(query, umi_len, &q)) return ORACLE_ERR_BAD_BASE;
    for (int64_t i = 0; i < n; i++) {
        if (to_bitset(umi_ascii + i * umi_len, umi_len, &o)) return ORACLE_ERR_BAD_BASE;
        int dist = umi_dist(&q, &o, nw);
        out[i] = (dist <= k && (dist == 0 || freq[i] <= max_freq)) ? 1 : 0;
    }
    return ORACLE_OK;
}

/* ---------------------------------------------------------------------------------------
 * Whole path: grouping + count/merge (deduplicate_sam.rs:131-176), cluster loop (:207-233),
 * counters (:243-267).
 * -------------------------------------------------------------------------------------*/
typedef struct {
    int64_t total_reads;       /* deduplicate_sam.rs:100  */
    int64_t n_buckets;         /* :198 align.len()        */
    int64_t total_umis;        /* :217                    */
    int64_t max_umis;          /* :218                    */
    int64_t n_kept;            /* :219 deduped_count      */
    uint64_t dist_calls;       /* number of umi_dist evaluations the faithful path made */
    uint64_t unordered_pairs;  /* sum_b N_b (N_b - 1) / 2 (the pair-comparison metric's numerator) */
} oracle_counters;

typedef struct { int32_t tid; int64_t pos; uint8_t rev; int64_t tlen; } bkey_t;

static uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x;
}
static uint64_t hash_bkey(int32_t tid, int64_t pos, uint8_t rev) {
    return mix64(((uint64_t)(uint32_t)tid << 1 | rev) * 0x9E3779B97F4A7C15ULL ^ mix64((uint64_t)pos));
}
static uint64_t hash_bits(const bitset_t *b, int nw, uint64_t bucket) {
    uint64_t h = bucket * 0x9E3779B97F4A7C15ULL;
    for (int i = 0; i < nw; i++) h = mix64(h ^ (uint64_t)b->bits[i]);
    return h;
}

/* Merge::merge, merge/mod.rs:18-51: true = keep existing */
static int merge_keep_existing(int merge, int32_t a_score, int32_t b_score) {
    if (merge == MERGE_ANY) return 1;
    return a_score >= b_score;
}

/*
 * n reads as SoA: tid, unclipped position (utils/mod.rs:96-104, computed by the caller), strand,
 * UMI (ASCII, n*umi_len), score (avg_qual utils/read.rs:56-63 or MAPQ :77-79; ignored for ANY).
 * kept_out receives the ascending read indices of the survivors (capacity n); per_read_root
 * (optional, capacity n) receives for every read the read index of its cluster's emitted
 * representative (-1 when unknown).  max_bucket_umis_to_cluster: buckets with more unique UMIs
 * are truncated to that many canonical-first UMIs before clustering when > 0 (only used by the
 * bounded cpu_baseline timing; 0 = never truncate).
 */
/* HOT LOOP B (deduplicate_sam.rs:207-233) over a shared bucket counter.  One worker = the reference (it clusters on one
 * thread); oracle_set_threads(n > 1) is the "all-core" variant that SURVEY §8(d) asks to be reported beside
 * it — buckets are independent, so they are handed out to n threads; this is NOT what the reference does. */
static int g_oracle_threads = 1;
void oracle_set_threads(int n) { g_oracle_threads = n; }
typedef struct {
    int64_t nb; const int64_t *bstart, *slot; const uentry_t *ue; const int64_t *read_uid;
    int umi_len, algo, k; float percentage; int64_t max_bucket;
    uint8_t *kept_flag; int64_t *uroot;
    volatile int64_t next; volatile int rc;
} bucket_job_t;
typedef struct { bucket_job_t *job; oracle_counters ctr; } bucket_worker_t;

static void *bucket_worker(void *arg) {
    bucket_worker_t *w = (bucket_worker_t *)arg;
    bucket_job_t *J = w->job;
    oracle_counters *ctr = &w->ctr;
    for (;;) {
        int64_t b = __atomic_fetch_add(&J->next, 1, __ATOMIC_RELAXED);
        if (b >= J->nb || J->rc) break;
        int64_t nbu = J->bstart[b + 1] - J->bstart[b];
        ctr->total_umis += nbu;                                     /* :217 */
        if (nbu > ctr->max_umis) ctr->max_umis = nbu;               /* :218 */
        ctr->unordered_pairs += (uint64_t)nbu * (uint64_t)(nbu - 1) / 2;
        uentry_t *e = (uentry_t *)malloc(sizeof(uentry_t) * nbu);
        uint8_t *keep = (uint8_t *)malloc(nbu);
        int32_t *label = (int32_t *)malloc(sizeof(int32_t) * nbu);
        if (!e || !keep || !label) { J->rc = ORACLE_ERR_NOMEM; break; }
        for (int64_t j = 0; j < nbu; j++) { e[j] = J->ue[J->slot[J->bstart[b] + j]]; }
        g_umi_len_cmp = J->umi_len;
        qsort(e, nbu, sizeof(uentry_t), cmp_canon);
        int64_t ncl = nbu;
        if (J->max_bucket > 0 && ncl > J->max_bucket) ncl = J->max_bucket;
        int rc = cluster_entries(e, ncl, J->umi_len, J->algo, J->k, J->percentage, keep, label, &ctr->dist_calls);
        if (rc) { J->rc = rc; break; }
        for (int64_t j = 0; j < ncl; j++) {
            if (keep[j]) { J->kept_flag[e[j].rep] = 1; ctr->n_kept++; }   /* :219, :227-231 */
        }
        if (J->uroot) {
            /* map entries back to unique ids through their representative's read_uid */
            for (int64_t j = 0; j < ncl; j++)
                J->uroot[J->read_uid[e[j].rep]] = label[j] < 0 ? -1 : e[label[j]].rep;
        }
        free(e); free(keep); free(label);
    }
    return NULL;
}

/* tlen == NULL: Align::Unpaired(Alignment), deduplicate_sam.rs:141-145; otherwise Align::Paired(PairedAlignment)
 * whose Eq/Hash add the template length, deduplicate_sam.rs:133-139 and :545-600. */
int oracle_dedup_paired(int64_t n, const int32_t *tid, const int64_t *pos, const uint8_t *rev, const int64_t *tlen,
                 const uint8_t *umi_ascii, int umi_len, const int32_t *score,
                 int algo, int merge, int k, float percentage,
                 int64_t *kept_out, int64_t *per_read_root, oracle_counters *ctr,
                 int64_t max_bucket_umis_to_cluster) {
    if (n < 0 || umi_len <= 0 || nwords_for(umi_len) > MAXW) return ORACLE_ERR_ARG;
    int nw = nwords_for(umi_len);
    memset(ctr, 0, sizeof(*ctr));
    ctr->total_reads = n;
    if (n == 0) return ORACLE_OK;

    /* bucket table: open addressing, key (rev, pos, tid) — Alignment eq deduplicate_sam.rs:507-514 */
    uint64_t bcap = 16; while (bcap < (uint64_t)n * 2) bcap <<= 1;
    int64_t *btab = (int64_t *)malloc(sizeof(int64_t) * bcap);      /* -> bucket id or -1 */
    bkey_t *bkeys = (bkey_t *)malloc(sizeof(bkey_t) * n);
    /* (bucket, umi) table */
    uint64_t ucap = bcap;
    int64_t *utab = (int64_t *)malloc(sizeof(int64_t) * ucap);      /* -> unique id or -1 */
    uentry_t *ue = (uentry_t *)malloc(sizeof(uentry_t) * n);
    int64_t *ubucket = (int64_t *)malloc(sizeof(int64_t) * n);
    int64_t *read_uid = (int64_t *)malloc(sizeof(int64_t) * n);
    if (!btab || !bkeys || !utab || !ue || !ubucket || !read_uid) return ORACLE_ERR_NOMEM;
    for (uint64_t i = 0; i < bcap; i++) btab[i] = -1;
    for (uint64_t i = 0; i < ucap; i++) utab[i] = -1;
    int64_t nb = 0, nu = 0;

    for (int64_t i = 0; i < n; i++) {                               /* HOT LOOP A, deduplicate_sam.rs:93 */
        /* align.entry(alignment).or_insert_with(...)  :148-150 */
        const int64_t tl = tlen ? tlen[i] : 0;
        uint64_t h = (hash_bkey(tid[i], pos[i], rev[i]) ^ ((uint64_t)tl * 0x9E3779B97F4A7C15ull >> 17)) & (bcap - 1);
        int64_t b;
        for (;;) {
            b = btab[h];
            if (b < 0) { b = nb++; btab[h] = b; bkeys[b].tid = tid[i]; bkeys[b].pos = pos[i]; bkeys[b].rev = rev[i]; bkeys[b].tlen = tl; break; }
            if (bkeys[b].tid == tid[i] && bkeys[b].pos == pos[i] && bkeys[b].rev == rev[i] && bkeys[b].tlen == tl) break;
            h = (h + 1) & (bcap - 1);
        }
        /* get_umi + to_bitset :158 */
        bitset_t bs;
        if (to_bitset(umi_ascii + i * umi_len, umi_len, &bs)) return ORACLE_ERR_BAD_BASE;
        /* umi_reads.entry(umi) :160-176 ; BitSet Eq = bits only, bitset.rs:94-101 */
        uint64_t g = hash_bits(&bs, nw, (uint64_t)b) & (ucap - 1);
        int64_t u;
        for (;;) {
            u = utab[g];
            if (u < 0) {                                            /* Vacant :161-163 */
                u = nu++; utab[g] = u; ue[u].bs = bs; ue[u].ascii = umi_ascii + i * umi_len;
                ue[u].freq = 1; ue[u].rep = i; ue[u].score = score ? score[i] : 0; ubucket[u] = b;
                break;
            }
            if (ubucket[u] == b && memcmp(ue[u].bs.bits, bs.bits, sizeof(int64_t) * nw) == 0) {  /* Occupied :164-175 */
                int keep_existing = merge_keep_existing(merge, ue[u].score, score ? score[i] : 0);
                ue[u].freq += 1;
                if (!keep_existing) { ue[u].rep = i; ue[u].score = score ? score[i] : 0; }
                break;
            }
            g = (g + 1) & (ucap - 1);
        }
        read_uid[i] = u;
    }
    free(btab); free(utab);

    /* gather the unique UMIs bucket by bucket (counting sort by bucket id) */
    int64_t *bstart = (int64_t *)calloc(nb + 1, sizeof(int64_t));
    int64_t *slot = (int64_t *)malloc(sizeof(int64_t) * (nu ? nu : 1));
    if (!bstart || !slot) return ORACLE_ERR_NOMEM;
    for (int64_t u = 0; u < nu; u++) bstart[ubucket[u] + 1]++;
    for (int64_t b = 0; b < nb; b++) bstart[b + 1] += bstart[b];
    {
        int64_t *fill = (int64_t *)malloc(sizeof(int64_t) * (nb ? nb : 1));
        if (!fill) return ORACLE_ERR_NOMEM;
        memcpy(fill, bstart, sizeof(int64_t) * nb);
        for (int64_t u = 0; u < nu; u++) slot[fill[ubucket[u]]++] = u;
        free(fill);
    }

    uint8_t *kept_flag = (uint8_t *)calloc(n, 1);
    int64_t *uroot = (int64_t *)malloc(sizeof(int64_t) * (nu ? nu : 1));   /* unique id -> root read index */
    if (!kept_flag || !uroot) return ORACLE_ERR_NOMEM;
    for (int64_t u = 0; u < nu; u++) uroot[u] = -1;
    ctr->n_buckets = nb;

    {
        bucket_job_t job;
        job.nb = nb; job.bstart = bstart; job.slot = slot; job.ue = ue; job.read_uid = read_uid; job.umi_len = umi_len; job.algo = algo;
        job.k = k; job.percentage = percentage; job.max_bucket = max_bucket_umis_to_cluster; job.kept_flag = kept_flag;
        job.uroot = per_read_root ? uroot : NULL; job.next = 0; job.rc = 0;
        int nt = g_oracle_threads < 1 ? 1 : g_oracle_threads;
        if (nt > 256) nt = 256;
        bucket_worker_t wk[256];
        pthread_t th[256];
        for (int t = 0; t < nt; t++) { wk[t].job = &job; memset(&wk[t].ctr, 0, sizeof(oracle_counters)); }
        if (nt == 1) bucket_worker(&wk[0]);                          /* the reference: one clustering thread */
        else {
            for (int t = 0; t < nt; t++) pthread_create(&th[t], NULL, bucket_worker, &wk[t]);
            for (int t = 0; t < nt; t++) pthread_join(th[t], NULL);
        }
        if (job.rc) return job.rc;
        for (int t = 0; t < nt; t++) {
            ctr->total_umis += wk[t].ctr.total_umis; ctr->unordered_pairs += wk[t].ctr.unordered_pairs;
            ctr->n_kept += wk[t].ctr.n_kept; ctr->dist_calls += wk[t].ctr.dist_calls;
            if (wk[t].ctr.max_umis > ctr->max_umis) ctr->max_umis = wk[t].ctr.max_umis;
        }
    }
    int64_t w = 0;
    for (int64_t i = 0; i < n; i++) if (kept_flag[i]) kept_out[w++] = i;
    if (per_read_root) for (int64_t i = 0; i < n; i++) per_read_root[i] = uroot[read_uid[i]];
    free(kept_flag); free(uroot); free(bstart); free(slot);
    free(bkeys); free(ue); free(ubucket); free(read_uid);
    return ORACLE_OK;
}

/* UcSAMRead::new, utils/read.rs:56-63: avg_qual = (sum of qual bytes as f32 / seq_len as f32) as i32,
 * summed left to right in f32; seq_len == 0 gives NaN -> 0 (Rust saturating cast). */
int32_t oracle_avg_qual(const uint8_t *qual, int64_t len) {
    volatile float avg = 0.0f;
    for (int64_t i = 0; i < len; i++) avg = avg + (float)qual[i];
    volatile float q = avg / (float)len;
    if (q != q) return 0;
    if (q >= 2147483648.0f) return INT32_MAX;
    return (int32_t)q;
}

/* utils/mod.rs:96-104 get_unclipped_pos over a raw BAM CIGAR (op = v & 0xf, len = v >> 4).
 * rust-htslib 0.49.0 (Cargo.lock:751) is not in /root/reference; semantics restated from the BAM
 * specification: end_pos = pos + sum len(M,D,N,=,X); leading/trailing soft clips may sit inside
 * one hard clip.  (parity unpinned — SURVEY.md §8(c)) */
int64_t oracle_unclipped_pos(int64_t pos, int is_reverse, const uint32_t *cigar, int n_cigar) {
    if (!is_reverse) {
        int64_t soft = 0, hard = 0; int i = 0;
        if (i < n_cigar && (cigar[i] & 0xf) == 5) { hard = cigar[i] >> 4; i++; }
        if (i < n_cigar && (cigar[i] & 0xf) == 4) { soft = cigar[i] >> 4; }
        return pos - soft - hard;
    } else {
        int64_t end = pos, soft = 0, hard = 0; int i = n_cigar - 1;
        for (int j = 0; j < n_cigar; j++) {
            int op = cigar[j] & 0xf;
            if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) end += cigar[j] >> 4;
        }
        if (i >= 0 && (cigar[i] & 0xf) == 5) { hard = cigar[i] >> 4; i--; }
        if (i >= 0 && (cigar[i] & 0xf) == 4) { soft = cigar[i] >> 4; }
        return end - 1 + soft + hard;
    }
}

/* One raw BAM alignment record (int32 block_size + body, BAM specification §4.2) -> what HOT LOOP A extracts
 * from it: the unmapped filter (deduplicate_sam.rs:102-108), get_unclipped_pos (utils/mod.rs:96-104), get_umi
 * (utils/read.rs:96-111: first separator byte of the read name, then umi_len bytes) and the score
 * (avg_qual read.rs:56-63 or MAPQ :77-79).  Returns 0, or -10 no separator ("failed to get the umi"),
 * -11 name too short (slice panic in the reference), -12 malformed record. */
/* paired != 0 adds the filters of deduplicate_sam.rs:96-129; *cls receives ORACLE_CLS_* bits, *tlen record.insert_size() */
#define ORACLE_CLS_MATE     1   /* :96-98  skipped before total_read_count */
#define ORACLE_CLS_UNMAPPED 2   /* :102-108 and :118-121                    */
#define ORACLE_CLS_UNPAIRED 4   /* :111-116                                 */
#define ORACLE_CLS_CHIMERIC 8   /* :123-128                                 */
int oracle_bam_decode_paired(const uint8_t *r, uint64_t rec_len, int umi_len, uint8_t sep, int use_mapq,
                      int paired, int remove_unpaired, int remove_chimeric,
                      int32_t *tid, int64_t *pos, uint8_t *rev, uint8_t *umi_out, int32_t *score, uint8_t *valid,
                      int64_t *tlen, int32_t *cls) {
    if (rec_len < 36) return -12;
    uint32_t block_size; memcpy(&block_size, r, 4);
    if ((uint64_t)block_size + 4 > rec_len) return -12;
    int32_t ref_id, p; memcpy(&ref_id, r + 4, 4); memcpy(&p, r + 8, 4);
    uint32_t l_read_name = r[12], mapq = r[13];
    uint16_t n_cigar, flag; memcpy(&n_cigar, r + 16, 2); memcpy(&flag, r + 18, 2);
    uint32_t l_seq; memcpy(&l_seq, r + 20, 4);
    const uint8_t *qname = r + 36, *cigar = qname + l_read_name;
    const uint8_t *qual = cigar + 4 * (uint64_t)n_cigar + (l_seq + 1) / 2;
    if ((uint64_t)(qual - r) + l_seq > rec_len) return -12;
    *cls = 0; *valid = 0;
    if (paired && (flag & 0x1) && (flag & 0x80)) { *cls = ORACLE_CLS_MATE; return ORACLE_OK; }     /* :96-98 */
    if (flag & 0x4) { *cls = ORACLE_CLS_UNMAPPED; return ORACLE_OK; }                              /* :102-108 */
    if (paired) {
        int32_t mtid; memcpy(&mtid, r + 24, 4);
        if (!(flag & 0x1)) { *cls |= ORACLE_CLS_UNPAIRED; if (remove_unpaired) return ORACLE_OK; }           /* :111-116 */
        if ((flag & 0x1) && (flag & 0x8)) { *cls |= ORACLE_CLS_UNMAPPED; return ORACLE_OK; }                /* :118-121 */
        if ((flag & 0x1) && ref_id != mtid) { *cls |= ORACLE_CLS_CHIMERIC; if (remove_chimeric) return ORACLE_OK; } /* :123-128 */
        int32_t isz; memcpy(&isz, r + 32, 4); *tlen = isz;
    }
    *valid = 1;
    uint32_t *cig = (uint32_t *)malloc(4 * (size_t)(n_cigar ? n_cigar : 1));
    memcpy(cig, cigar, 4 * (size_t)n_cigar);
    *rev = (flag & 0x10) ? 1 : 0;
    *pos = oracle_unclipped_pos(p, *rev, cig, n_cigar);
    free(cig);
    *tid = ref_id;
    uint32_t name_len = l_read_name ? l_read_name - 1 : 0, q = 0;
    while (q < name_len && qname[q] != sep) q++;
    if (q >= name_len) return -10;
    if (q + 1 + (uint32_t)umi_len > name_len) return -11;
    memcpy(umi_out, qname + q + 1, umi_len);
    *score = use_mapq ? (int32_t)mapq : oracle_avg_qual(qual, l_seq);
    return ORACLE_OK;
}

int oracle_bam_decode(const uint8_t *r, uint64_t rec_len, int umi_len, uint8_t sep, int use_mapq,
                      int32_t *tid, int64_t *pos, uint8_t *rev, uint8_t *umi_out, int32_t *score, uint8_t *valid) {
    int64_t tlen; int32_t cls;
    return oracle_bam_decode_paired(r, rec_len, umi_len, sep, use_mapq, 0, 0, 0, tid, pos, rev, umi_out, score, valid, &tlen, &cls);
}

int oracle_dedup(int64_t n, const int32_t *tid, const int64_t *pos, const uint8_t *rev,
                 const uint8_t *umi_ascii, int umi_len, const int32_t *score,
                 int algo, int merge, int k, float percentage,
                 int64_t *kept_out, int64_t *per_read_root, oracle_counters *ctr,
                 int64_t max_bucket_umis_to_cluster) {
    return oracle_dedup_paired(n, tid, pos, rev, NULL, umi_ascii, umi_len, score, algo, merge, k, percentage,
                               kept_out, per_read_root, ctr, max_bucket_umis_to_cluster);
}
