// umicollapse_gpu — C++ twin of the umi-collapse-rs command line on top of libumigpu's C ABI.
//
// The reference host is Rust (clap + rust-htslib); this image has no Rust toolchain, so this program is the
// compiled, tested stand-in for it: same flags (src/cli.rs:7-77), same defaults and validation
// (src/main.rs:33-47), same dispatch (src/main.rs:49-92), same end-of-run counters
// (src/deduplicate_sam.rs:243-267).  Host code only moves bytes (BGZF inflate/deflate with zlib, header
// parsing, writing the surviving records); every per-record computation and the clustering run on the GPU.
//
// Differences from the reference, all deliberate: output is in input order, --algo cc
// works, --mode fastq works (the reference's is an empty TODO, main.rs:49-51), --tag is implemented from its help
// text (src/cli.rs:64-76; the reference collects ClusterTrackers and then writes nothing, deduplicate_sam.rs:236-239),
// --paired follows deduplicate_sam.rs:96-129 (filters, on the device) and UcWriter::write_reversed (:409-462, mates).
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_set>
#include <vector>

#include "../../include/umigpu.h"

struct Cli {                                   // src/cli.rs:7-77 (same names, same defaults)
    std::string mode = "bam", input, output, algo_str = "dir", merge_str, data_str = "ngrambktree";
    int k = 1; unsigned umi_length = 0; float percentage = 0.5f; unsigned num_threads = 1; unsigned char umi_separator = '_';
    bool two_pass = false, paired = false, remove_unpaired = false, remove_chimeric = false, keep_unmapped = false, track_clusters = false;
    int device = 0;
};

static std::thread *g_cuda_init = nullptr;     // CUDA start-up helper thread (main): joined before the process exits
[[noreturn]] static void die(const std::string &msg) {
    fprintf(stderr, "umicollapse_gpu: %s\n", msg.c_str());
    if (g_cuda_init && g_cuda_init->joinable() && g_cuda_init->get_id() != std::this_thread::get_id()) g_cuda_init->join();
    exit(2);
}
static void check(int rc, umigpu_ctx *ctx, const char *what) {
    if (rc != 0) die(std::string(what) + " failed (" + std::to_string(rc) + "): " + umigpu_last_error(ctx));   // the reference panics
}

static Cli parse(int argc, char **argv) {
    Cli a;
    auto need = [&](int &i) -> const char * { if (i + 1 >= argc) die(std::string("missing value for ") + argv[i]); return argv[++i]; };
    for (int i = 1; i < argc; i++) {
        std::string s = argv[i];
        if (s == "-m" || s == "--mode") a.mode = need(i);
        else if (s == "-i") a.input = need(i);
        else if (s == "-o") a.output = need(i);
        else if (s == "-k") a.k = atoi(need(i));
        else if (s == "-u") a.umi_length = (unsigned)atoi(need(i));
        else if (s == "-p") a.percentage = (float)atof(need(i));
        else if (s == "--num-threads") a.num_threads = (unsigned)atoi(need(i));
        else if (s == "--umi_sep") { const char *v = need(i); a.umi_separator = (strlen(v) == 1 && (v[0] < '0' || v[0] > '9')) ? (unsigned char)v[0] : (unsigned char)atoi(v); }
        else if (s == "--algo") a.algo_str = need(i);
        else if (s == "--merge") a.merge_str = need(i);
        else if (s == "--data") a.data_str = need(i);
        else if (s == "--two-pass") a.two_pass = true;
        else if (s == "--paired") a.paired = true;
        else if (s == "--remove-unpaired") a.remove_unpaired = true;
        else if (s == "--remove-chimeric") a.remove_chimeric = true;
        else if (s == "--keep-unmapped") a.keep_unmapped = true;
        else if (s == "--tag") a.track_clusters = true;
        else if (s == "--device") a.device = atoi(need(i));
        else die("unexpected argument '" + s + "'");
    }
    if (a.input.empty() || a.output.empty()) die("the following required arguments were not provided: -i <INPUT_FILE> -o <OUITPUT_FILE>");
    if (a.merge_str.empty()) a.merge_str = a.mode == "fastq" ? "avgqual" : "mapqual";                 // main.rs:33-39
    if (a.track_clusters && a.two_pass) die("Cannot track clusters with the two pass algorithm!");    // main.rs:41-43
    if (a.paired && a.keep_unmapped) die("Cannot keep unmapped reads with paired-end reads!");        // main.rs:45-47
    return a;
}

// Uninitialised byte buffer: std::vector would zero-fill gigabytes on one thread before the parallel inflate touches them;
// here the pages are first touched (and faulted in) by the worker threads.  Also wraps a read-only file mapping.
struct RawBuf {
    uint8_t *p = nullptr; size_t n = 0; bool mapped = false;
    RawBuf() = default;
    explicit RawBuf(size_t bytes) : p(bytes ? (uint8_t *)malloc(bytes) : nullptr), n(bytes) { if (bytes && !p) { fprintf(stderr, "umicollapse_gpu: out of memory\n"); exit(2); } }
    RawBuf(const RawBuf &) = delete; RawBuf &operator=(const RawBuf &) = delete;
    RawBuf(RawBuf &&o) noexcept : p(o.p), n(o.n), mapped(o.mapped) { o.p = nullptr; o.n = 0; }
    RawBuf &operator=(RawBuf &&o) noexcept { release(); p = o.p; n = o.n; mapped = o.mapped; o.p = nullptr; o.n = 0; return *this; }
    ~RawBuf() { release(); }
    void release() { if (p) { if (mapped) munmap(p, n); else free(p); } p = nullptr; n = 0; }
    uint8_t *data() { return p; } const uint8_t *data() const { return p; }
    size_t size() const { return n; }
    uint8_t operator[](size_t i) const { return p[i]; }
};
static RawBuf map_file(const std::string &path) {
    int fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) { fprintf(stderr, "umicollapse_gpu: Invalid input path: %s\n", path.c_str()); exit(2); }
    struct stat st; fstat(fd, &st);
    RawBuf b;
    if (st.st_size > 0) {
        void *m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) { fprintf(stderr, "umicollapse_gpu: cannot map %s\n", path.c_str()); exit(2); }
        madvise(m, (size_t)st.st_size, MADV_SEQUENTIAL);
        b.p = (uint8_t *)m; b.n = (size_t)st.st_size; b.mapped = true;
    }
    close(fd);
    return b;
}

static std::vector<uint8_t> read_file(const std::string &p) {
    FILE *f = fopen(p.c_str(), "rb");
    if (!f) die("Invalid input path: " + p);
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<uint8_t> v((size_t)n);
    if (n && fread(v.data(), 1, (size_t)n, f) != (size_t)n) die("short read on " + p);
    fclose(f);
    return v;
}

template <class F> static void parallel_for(size_t n, unsigned threads, F f) {
    threads = std::max(1u, std::min<unsigned>(threads, (unsigned)std::max<size_t>(1, n)));
    std::atomic<size_t> next{0};
    std::vector<std::thread> ts;
    for (unsigned t = 0; t < threads; t++) ts.emplace_back([&] { for (size_t i; (i = next.fetch_add(1)) < n;) f(i); });
    for (auto &t : ts) t.join();
}

// ---- BGZF (SAM/BAM specification §4.1): independent gzip members with a BC extra field ----
static RawBuf bgzf_inflate(const RawBuf &in, unsigned threads) {
    struct Blk { size_t off, csize, usize, uoff; };
    std::vector<Blk> blks;
    size_t off = 0, utotal = 0;
    while (off + 18 <= in.size()) {
        if (in[off] != 0x1f || in[off + 1] != 0x8b) die("not a BGZF stream");
        unsigned xlen = in[off + 10] | (in[off + 11] << 8);
        size_t x = off + 12, xend = x + xlen, bsize = 0;
        while (x + 4 <= xend) { unsigned slen = in[x + 2] | (in[x + 3] << 8); if (in[x] == 'B' && in[x + 1] == 'C') bsize = (size_t)(in[x + 4] | (in[x + 5] << 8)) + 1; x += 4 + slen; }
        if (!bsize || off + bsize > in.size()) die("corrupt BGZF block");
        size_t usize = in[off + bsize - 4] | (in[off + bsize - 3] << 8) | (in[off + bsize - 2] << 16) | ((size_t)in[off + bsize - 1] << 24);
        blks.push_back({off + 12 + xlen, bsize - 12 - xlen - 8, usize, utotal});
        utotal += usize; off += bsize;
    }
    RawBuf out(utotal);
    parallel_for(blks.size(), threads, [&](size_t i) {
        const Blk &b = blks[i];
        if (!b.usize) return;
        z_stream z; memset(&z, 0, sizeof z);
        if (inflateInit2(&z, -15) != Z_OK) die("inflateInit2");
        z.next_in = const_cast<Bytef *>(in.data() + b.off); z.avail_in = (uInt)b.csize;
        z.next_out = out.data() + b.uoff; z.avail_out = (uInt)b.usize;
        if (inflate(&z, Z_FINISH) != Z_STREAM_END) die("inflate failed");
        inflateEnd(&z);
    });
    return out;
}

static void bgzf_write(const std::string &path, const std::vector<const uint8_t *> &ptr, const std::vector<size_t> &len, unsigned threads) {
    // gather into one stream, cut into 0xff00-byte blocks, deflate the blocks in parallel
    std::vector<size_t> ooff(len.size() + 1, 0);
    for (size_t i = 0; i < len.size(); i++) ooff[i + 1] = ooff[i] + len[i];
    const size_t total = ooff.back();
    RawBuf data(total);
    {   // gather on all threads, contiguous slices of the record list
        const size_t nrec = ptr.size(), per = (nrec + threads - 1) / std::max(1u, threads);
        parallel_for(std::max(1u, threads), threads, [&](size_t t) {
            for (size_t i = t * per; i < std::min(nrec, (t + 1) * per); i++) memcpy(data.data() + ooff[i], ptr[i], len[i]);
        });
    }
    const size_t B = 0xff00, nb = (total + B - 1) / B;
    std::vector<std::vector<uint8_t>> comp(nb);
    parallel_for(nb, threads, [&](size_t i) {
        size_t s = i * B, n = std::min(B, total - s);
        std::vector<uint8_t> &c = comp[i];
        c.resize(compressBound((uLong)n) + 32);
        z_stream z; memset(&z, 0, sizeof z);
        if (deflateInit2(&z, 1, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) die("deflateInit2");
        z.next_in = data.data() + s; z.avail_in = (uInt)n; z.next_out = c.data() + 18; z.avail_out = (uInt)(c.size() - 26);
        if (deflate(&z, Z_FINISH) != Z_STREAM_END) die("deflate failed");
        size_t cs = z.total_out; deflateEnd(&z);
        const uint8_t h[12] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0};
        memcpy(c.data(), h, 12); c[12] = 'B'; c[13] = 'C'; c[14] = 2; c[15] = 0;
        size_t bsize = cs + 26 - 1; c[16] = bsize & 0xff; c[17] = (bsize >> 8) & 0xff;
        uint32_t crc = (uint32_t)crc32(crc32(0, nullptr, 0), data.data() + s, (uInt)n), isz = (uint32_t)n;
        memcpy(c.data() + 18 + cs, &crc, 4); memcpy(c.data() + 22 + cs, &isz, 4);
        c.resize(cs + 26);
    });
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) die("cannot open output " + path);
    for (auto &c : comp) fwrite(c.data(), 1, c.size(), f);
    static const uint8_t eof[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    fwrite(eof, 1, 28, f);
    fclose(f);
}

// ---- incremental BGZF reader / writer for --two-pass (bounded host memory) ----
struct BgzfStream {
    FILE *f; unsigned threads; bool eof = false;
    explicit BgzfStream(const std::string &p, unsigned t) : f(fopen(p.c_str(), "rb")), threads(t) { if (!f) die("Invalid input path: " + p); }
    ~BgzfStream() { if (f) fclose(f); }
    // appends the inflated bytes of up to max_blocks BGZF blocks to out; returns false at end of file
    bool next(std::vector<uint8_t> &out, size_t max_blocks = 1024) {
        if (eof) return false;
        std::vector<std::vector<uint8_t>> raw;
        std::vector<size_t> usize;
        while (raw.size() < max_blocks) {
            uint8_t h[18];
            size_t got = fread(h, 1, 18, f);
            if (got == 0) { eof = true; break; }
            if (got != 18 || h[0] != 0x1f || h[1] != 0x8b) die("not a BGZF stream");
            unsigned xlen = h[10] | (h[11] << 8);
            std::vector<uint8_t> blk(h, h + 18);
            // the BC subfield is first in every BGZF block written by htslib / this program
            if (!(h[12] == 'B' && h[13] == 'C')) die("unsupported BGZF extra field layout");
            size_t bsize = (size_t)(h[16] | (h[17] << 8)) + 1;
            blk.resize(bsize);
            if (fread(blk.data() + 18, 1, bsize - 18, f) != bsize - 18) die("truncated BGZF block");
            (void)xlen;
            usize.push_back(blk[bsize - 4] | (blk[bsize - 3] << 8) | (blk[bsize - 2] << 16) | ((size_t)blk[bsize - 1] << 24));
            raw.push_back(std::move(blk));
        }
        if (raw.empty()) return false;
        std::vector<size_t> uoff(raw.size());
        size_t base = out.size(), tot = 0;
        for (size_t i = 0; i < raw.size(); i++) { uoff[i] = tot; tot += usize[i]; }
        out.resize(base + tot);
        parallel_for(raw.size(), threads, [&](size_t i) {
            if (!usize[i]) return;
            const std::vector<uint8_t> &b = raw[i];
            unsigned xlen = b[10] | (b[11] << 8);
            z_stream z; memset(&z, 0, sizeof z);
            if (inflateInit2(&z, -15) != Z_OK) die("inflateInit2");
            z.next_in = const_cast<Bytef *>(b.data() + 12 + xlen); z.avail_in = (uInt)(b.size() - 12 - xlen - 8);
            z.next_out = out.data() + base + uoff[i]; z.avail_out = (uInt)usize[i];
            if (inflate(&z, Z_FINISH) != Z_STREAM_END) die("inflate failed");
            inflateEnd(&z);
        });
        return true;
    }
};
struct BgzfOut {
    std::string path; unsigned threads; FILE *f; std::vector<uint8_t> pend;
    BgzfOut(const std::string &p, unsigned t) : path(p), threads(t), f(fopen(p.c_str(), "wb")) { if (!f) die("cannot open output " + p); }
    void add(const uint8_t *p, size_t n) { pend.insert(pend.end(), p, p + n); if (pend.size() >= (64u << 20)) flush(false); }
    void flush(bool all) {
        const size_t B = 0xff00;
        size_t nb = all ? (pend.size() + B - 1) / B : pend.size() / B;
        if (!nb) return;
        std::vector<std::vector<uint8_t>> comp(nb);
        parallel_for(nb, threads, [&](size_t i) {
            size_t s = i * B, n = std::min(B, pend.size() - s);
            std::vector<uint8_t> &c = comp[i];
            c.resize(compressBound((uLong)n) + 32);
            z_stream z; memset(&z, 0, sizeof z);
            if (deflateInit2(&z, 1, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) die("deflateInit2");
            z.next_in = pend.data() + s; z.avail_in = (uInt)n; z.next_out = c.data() + 18; z.avail_out = (uInt)(c.size() - 26);
            if (deflate(&z, Z_FINISH) != Z_STREAM_END) die("deflate failed");
            size_t cs = z.total_out; deflateEnd(&z);
            const uint8_t h[12] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0};
            memcpy(c.data(), h, 12); c[12] = 'B'; c[13] = 'C'; c[14] = 2; c[15] = 0;
            size_t bsize = cs + 26 - 1; c[16] = bsize & 0xff; c[17] = (bsize >> 8) & 0xff;
            uint32_t crc = (uint32_t)crc32(crc32(0, nullptr, 0), pend.data() + s, (uInt)n), isz = (uint32_t)n;
            memcpy(c.data() + 18 + cs, &crc, 4); memcpy(c.data() + 22 + cs, &isz, 4);
            c.resize(cs + 26);
        });
        for (auto &c : comp) fwrite(c.data(), 1, c.size(), f);
        size_t used = std::min(pend.size(), nb * B);
        pend.erase(pend.begin(), pend.begin() + used);
    }
    void close() {
        flush(true);
        static const uint8_t eof[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        fwrite(eof, 1, 28, f); fclose(f); f = nullptr;
    }
};

static int algo_code(const Cli &a) {                       // main.rs:52-92 (+ cc, which the reference only documents)
    if (a.algo_str == "dir") return UMIGPU_ALGO_DIR;
    if (a.algo_str == "adj") return UMIGPU_ALGO_ADJ;
    if (a.algo_str == "cc") return UMIGPU_ALGO_CC;
    if (a.algo_str == "adj-upstream") return UMIGPU_ALGO_ADJ_UPSTREAM;
    die("Invalid algorithm combination: " + a.algo_str + " , " + a.merge_str + " and " + a.data_str);
}
static int merge_code(const Cli &a) {
    if (a.merge_str == "any") return UMIGPU_MERGE_ANY;
    if (a.merge_str == "avgqual") return UMIGPU_MERGE_AVGQUAL;
    if (a.merge_str == "mapqual") return UMIGPU_MERGE_MAPQUAL;
    die("Invalid algorithm combination: " + a.algo_str + " , " + a.merge_str + " and " + a.data_str);
}

static void report(const umigpu_counters &c, uint64_t unmapped, bool paired = false) {      // deduplicate_sam.rs:243-267
    fprintf(stderr, "Number of input reads: %llu\n", (unsigned long long)c.total_reads);
    fprintf(stderr, "Number of removed unmapped reads: %llu\n", (unsigned long long)unmapped);
    if (paired) {                                                                           // :245-248
        fprintf(stderr, "Number of unpaired reads: %llu\n", (unsigned long long)c.n_unpaired);
        fprintf(stderr, "Number of chimeric reads: %llu\n", (unsigned long long)c.n_chimeric);
    }
    fprintf(stderr, "Number of unique alignment positions: %llu\n", (unsigned long long)c.n_buckets);
    fprintf(stderr, "Number of UMIs: %llu\n", (unsigned long long)c.total_umis);
    fprintf(stderr, "Average number of UMIs per alignment position: %g\n", c.n_buckets ? (double)c.total_umis / (double)c.n_buckets : 0.0);
    fprintf(stderr, "Max number of UMIs over all alignment positions: %llu\n", (unsigned long long)c.max_umis);
    fprintf(stderr, "Number of reads after deduplicating: %llu\n", (unsigned long long)c.n_kept);   // with --tag: "Number of groups of reads" (:261-263)
}

static umigpu_ctx *make_ctx(const Cli &a, unsigned umi_len) {
    umigpu_config cfg; memset(&cfg, 0, sizeof cfg);
    cfg.flags = a.track_clusters ? UMIGPU_FLAG_LABELS : 0;
    if (a.paired) cfg.flags |= UMIGPU_FLAG_PAIRED | (a.remove_unpaired ? UMIGPU_FLAG_REMOVE_UNPAIRED : 0) | (a.remove_chimeric ? UMIGPU_FLAG_REMOVE_CHIMERIC : 0);
    cfg.k = a.k; cfg.percentage = a.percentage; cfg.algo = algo_code(a); cfg.merge = merge_code(a); cfg.umi_len = umi_len; cfg.device = a.device;
    umigpu_ctx *ctx = nullptr;
    check(umigpu_create(&cfg, &ctx), nullptr, "umigpu_create");
    return ctx;
}

// ---- host mirror of the record filters (deduplicate_sam.rs:96-129): used for -u 0 autodetection, which looks at the
// first read that reaches UcSAMRead::new (:152-156), and for batches that are seen before the UMI length is known ----
enum { CLS_MATE = 1, CLS_UNMAPPED = 2, CLS_UNPAIRED = 4, CLS_CHIMERIC = 8 };
static inline unsigned rec_flag(const uint8_t *r) { return r[18] | (r[19] << 8); }
static inline int32_t rec_i32(const uint8_t *r, size_t o) { int32_t v; memcpy(&v, r + o, 4); return v; }
static bool rec_passes(const uint8_t *r, const Cli &a, unsigned *cls_out = nullptr) {
    const unsigned f = rec_flag(r);
    unsigned cls = 0; bool ok = true;
    if (a.paired && (f & 0x1) && (f & 0x80)) { if (cls_out) *cls_out = CLS_MATE; return false; }
    if (f & 0x4) { cls = CLS_UNMAPPED; ok = false; }
    else if (a.paired) {
        if (!(f & 0x1)) { cls |= CLS_UNPAIRED; if (a.remove_unpaired) ok = false; }
        if (ok && (f & 0x1) && (f & 0x8)) { cls |= CLS_UNMAPPED; ok = false; }
        if (ok && (f & 0x1) && rec_i32(r, 4) != rec_i32(r, 24)) { cls |= CLS_CHIMERIC; if (a.remove_chimeric) ok = false; }
    }
    if (cls_out) *cls_out = cls;
    return ok;
}
static void count_unfed(const uint8_t *r, const Cli &a, umigpu_counters &c, uint64_t &unmapped) {
    unsigned cls = 0; rec_passes(r, a, &cls);
    if (!(cls & CLS_MATE)) c.total_reads++; else c.n_mates_skipped++;
    if (cls & CLS_UNMAPPED) unmapped++;
    if (cls & CLS_UNPAIRED) c.n_unpaired++;
    if (cls & CLS_CHIMERIC) c.n_chimeric++;
}
// the caseless regex ^(?:.*?)SEP([ATCGN]+)(?:.*?)$ (utils/read.rs:65-75,87-94): first separator that is followed by an [ATCGN] letter
static unsigned detect_umi_len(const uint8_t *r, uint8_t sep) {
    unsigned l_name = r[12]; const uint8_t *name = r + 36; unsigned len = l_name ? l_name - 1 : 0, p = 0, n = 0;
    for (;; p++) {
        while (p < len && name[p] != sep) p++;
        if (p >= len) die("failed to get the umi");
        if (p + 1 < len && name[p + 1] && strchr("ACGTNacgtn", name[p + 1])) break;
    }
    for (unsigned q = p + 1; q < len && name[q] && strchr("ACGTNacgtn", name[q]); q++) n++;
    return n;
}
// ---- --paired: the mates of the kept reads.  UcWriter::write (deduplicate_sam.rs:382-407) remembers, for every written
// paired read, ReverseRead{qname, mate ref, mate pos}; write_reversed (:409-462) re-reads the input and writes every mapped,
// paired, last-in-template record with a mapped mate whose {qname, ref, pos} is in the set, removing the entry so that a
// repeated mate record is written once.  (ReverseRead's Eq ignores the coordinate while its Hash includes it, :289-322,
// so a lookup only ever finds an entry with the same coordinate short of a hash collision; the key here is all three.)
static std::string mate_key(const uint8_t *name, unsigned name_len, int32_t tid, int32_t pos) {
    std::string k((const char *)name, name_len);
    k.append((const char *)&tid, 4); k.append((const char *)&pos, 4);
    return k;
}
struct MateSet {
    std::unordered_set<std::string> set;
    void remember(const uint8_t *r) {                                  // :393-399
        if (!(rec_flag(r) & 0x1)) return;
        set.insert(mate_key(r + 36, r[12] ? r[12] - 1 : 0, rec_i32(r, 24), rec_i32(r, 28)));
    }
    bool take(const uint8_t *r) {                                      // :429-459
        const unsigned f = rec_flag(r);
        if ((f & 0x4) || !(f & 0x1) || !(f & 0x80) || (f & 0x8)) return false;
        auto it = set.find(mate_key(r + 36, r[12] ? r[12] - 1 : 0, rec_i32(r, 4), rec_i32(r, 8)));
        if (it == set.end()) return false;
        set.erase(it);
        return true;
    }
};

// phase timing on stderr (the reference logs "UMI collapsing reading finished in ... seconds", deduplicate_sam.rs:179-182)
struct PhaseClock {
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    void lap(const char *what) {
        auto n = std::chrono::steady_clock::now();
        fprintf(stderr, "phase %s: %.3f s\n", what, std::chrono::duration<double>(n - t).count());
        t = n;
    }
};

// ---- --mode bam ----
static int run_bam(const Cli &a) {
    PhaseClock pc;
    RawBuf raw = map_file(a.input);
    RawBuf buf = bgzf_inflate(raw, a.num_threads);
    raw.release();
    if (buf.size() < 12 || memcmp(buf.data(), "BAM\1", 4) != 0) die("Invalid input path: not a BAM file");
    int32_t l_text, n_ref; memcpy(&l_text, buf.data() + 4, 4);
    size_t off = 8 + (size_t)l_text; memcpy(&n_ref, buf.data() + off, 4); off += 4;
    for (int r = 0; r < n_ref; r++) { int32_t l; memcpy(&l, buf.data() + off, 4); off += 4 + (size_t)l + 4; }
    const size_t first = off;
    std::vector<uint64_t> offs((buf.size() - first) / 36 + 2);
    uint64_t n = 0, consumed = 0;
    check(umigpu_bam_record_offsets(buf.data() + first, buf.size() - first, offs.data(), offs.size() - 1, &n, &consumed), nullptr, "umigpu_bam_record_offsets");
    pc.lap("read + inflate + record offsets");
    // -u 0: autodetect from the first mapped read (utils/read.rs:65-75,87-94)
    unsigned umi_len = a.umi_length;
    for (uint64_t i = 0; i < n && umi_len == 0; i++) {
        const uint8_t *r = buf.data() + first + offs[i];
        if (!rec_passes(r, a)) continue;
        umi_len = detect_umi_len(r, a.umi_separator);
        break;
    }
    std::vector<const uint8_t *> optr{buf.data()}; std::vector<size_t> olen{first};
    umigpu_counters ctr; memset(&ctr, 0, sizeof ctr);
    uint64_t unmapped = 0;
    if (n && umi_len) {
        umigpu_ctx *ctx = make_ctx(a, umi_len);
        pc.lap("context (CUDA init)");
        const uint64_t CH = 1ull << 22;
        for (uint64_t s = 0; s < n; s += CH) {
            uint64_t e = std::min(n, s + CH), nun = 0;
            check(umigpu_push_bam_records(ctx, e - s, buf.data() + first, offs.data() + s, a.umi_separator, s, &nun), ctx, "umigpu_push_bam_records");
            unmapped += nun;
        }
        umigpu_result res;
        check(umigpu_finish(ctx, &res), ctx, "umigpu_finish");
        pc.lap("device: record decode + dedup (umigpu_push_bam_records, umigpu_finish)");
        ctr = res.counters;
        std::vector<std::vector<uint8_t>> tagged;         // --tag: rewritten records (own storage)
        if (!a.track_clusters) {
            MateSet mates;
            if (a.paired) for (uint64_t j = 0; j < res.n_kept; j++) mates.remember(buf.data() + first + offs[res.kept_read_index[j]]);
            // merge kept indices with (optionally) the unmapped records and (--paired) the kept reads' mates, input order
            uint64_t kpos = 0;
            for (uint64_t i = 0; i < n; i++) {
                const uint8_t *r = buf.data() + first + offs[i];
                bool keep = kpos < res.n_kept && res.kept_read_index[kpos] == i;
                if (keep) kpos++;
                else if (a.keep_unmapped && ((r[18] | (r[19] << 8)) & 4)) keep = true;          // deduplicate_sam.rs:104-106
                else if (a.paired && mates.take(r)) keep = true;                                // :429-459
                if (keep) { optr.push_back(r); olen.push_back(offs[i + 1] - offs[i]); }
            }
        } else {
            // --tag (src/cli.rs:64-76): nothing is removed; every read but the consensus read of its cluster gets the
            // duplicate flag, MI = cluster id, RX = UMI of the consensus read, cs = cluster size (consensus read only),
            // su = reads with exactly this UMI (best read of the UMI only)
            std::vector<uint64_t> rec_of(res.n_reads);                 // pushed read j -> record number
            { uint64_t j = 0; for (uint64_t i = 0; i < n; i++) { const uint8_t *r = buf.data() + first + offs[i]; if (!((r[18] | (r[19] << 8)) & 4)) rec_of[j++] = i; } }
            std::vector<uint32_t> csize(n, 0), same(n, 0), cid(n, 0);
            for (uint64_t j = 0; j < res.n_reads; j++) { csize[res.read_cluster_root[j]]++; same[res.read_umi_rep[j]]++; }
            { uint32_t next = 0; for (uint64_t i = 0; i < n; i++) if (csize[i]) cid[i] = next++; }
            tagged.resize(n);
            uint64_t j = 0;
            for (uint64_t i = 0; i < n; i++) {
                const uint8_t *r = buf.data() + first + offs[i];
                const size_t len = offs[i + 1] - offs[i];
                if ((r[18] | (r[19] << 8)) & 4) { if (a.keep_unmapped) { optr.push_back(r); olen.push_back(len); } continue; }
                const uint64_t root = res.read_cluster_root[j], urep = res.read_umi_rep[j];
                j++;
                std::vector<uint8_t> &t = tagged[i];
                t.assign(r, r + len);
                if (root != i) { unsigned fl = (t[18] | (t[19] << 8)) | 0x400; t[18] = fl & 0xff; t[19] = (fl >> 8) & 0xff; }
                auto tag_i = [&](char x, char y, uint32_t v) { t.push_back(x); t.push_back(y); t.push_back('I'); for (int b = 0; b < 4; b++) t.push_back((v >> (8 * b)) & 0xff); };
                tag_i('M', 'I', cid[root]);
                { // RX:Z = UMI of the consensus read
                    const uint8_t *rr = buf.data() + first + offs[root]; const uint8_t *nm = rr + 36; unsigned nl = rr[12] ? rr[12] - 1 : 0, p = 0;
                    while (p < nl && nm[p] != a.umi_separator) p++;
                    t.push_back('R'); t.push_back('X'); t.push_back('Z');
                    for (unsigned q = 0; q < umi_len && p + 1 + q < nl; q++) t.push_back(nm[p + 1 + q]);
                    t.push_back(0);
                }
                if (root == i) tag_i('c', 's', csize[i]);
                if (urep == i) tag_i('s', 'u', same[i]);
                uint32_t bs = (uint32_t)(t.size() - 4); memcpy(t.data(), &bs, 4);
                optr.push_back(t.data()); olen.push_back(t.size());
            }
        }
        pc.lap("select output records");
        bgzf_write(a.output, optr, olen, a.num_threads);
        pc.lap("deflate + write");
        (void)ctx;    // no umigpu_destroy: the process exits right after the output is closed (freeing pinned buffers and the CUDA context costs seconds)
    } else {
        for (uint64_t i = 0; i < n; i++) {                       // nothing reached the device: every record failed a filter
            const uint8_t *r = buf.data() + first + offs[i];
            count_unfed(r, a, ctr, unmapped);
            if (a.keep_unmapped && (rec_flag(r) & 4)) { optr.push_back(r); olen.push_back(offs[i + 1] - offs[i]); }
        }
        bgzf_write(a.output, optr, olen, a.num_threads);
    }
    report(ctr, unmapped, a.paired);
    return 0;
}

// ---- --mode bam --two-pass (src/cli.rs:45-48: "should use much less memory"): the input is streamed twice, the
// host never holds more than one batch of inflated blocks; the first pass feeds the device, the last one writes the
// survivors; --paired adds a pass in between that collects the mate keys of the kept reads (UcWriter::write, :393-399) ----
static int run_bam_two_pass(const Cli &a) {
    umigpu_ctx *ctx = nullptr;
    unsigned umi_len = a.umi_length;
    uint64_t unmapped = 0;
    umigpu_counters unfed; memset(&unfed, 0, sizeof unfed);       // records of batches seen before the UMI length was known
    std::vector<uint8_t> header;
    MateSet mates;
    umigpu_result res; memset(&res, 0, sizeof res);
    const int n_pass = a.paired ? 3 : 2, last_pass = n_pass - 1;
    for (int pass = 0; pass < n_pass; pass++) {
        BgzfStream in(a.input, a.num_threads);
        std::vector<uint8_t> buf;
        std::vector<uint64_t> offs;
        bool have_header = false;
        uint64_t index = 0, kpos = 0;
        BgzfOut *out = nullptr;
        if (pass == 1 && ctx) check(umigpu_finish(ctx, &res), ctx, "umigpu_finish");
        if (pass == last_pass) {
            out = new BgzfOut(a.output, a.num_threads);
            out->add(header.data(), header.size());
        }
        // UMICOLLAPSE_BATCH_BLOCKS: BGZF blocks inflated per batch (default 1024 = up to 64 MiB; tests use small values)
        const char *bb = getenv("UMICOLLAPSE_BATCH_BLOCKS");
        const size_t batch_blocks = bb ? (size_t)std::max(1, atoi(bb)) : 1024;
        bool more = true;
        while (more) {
            more = in.next(buf, batch_blocks);
            size_t start = 0;
            if (!have_header) {
                if (buf.size() < 12) { if (more) continue; die("Invalid input path: not a BAM file"); }
                if (memcmp(buf.data(), "BAM\1", 4) != 0) die("Invalid input path: not a BAM file");
                int32_t l_text; memcpy(&l_text, buf.data() + 4, 4);
                size_t off = 8 + (size_t)l_text;
                if (buf.size() < off + 4) { if (more) continue; die("truncated BAM header"); }
                int32_t n_ref; memcpy(&n_ref, buf.data() + off, 4); off += 4;
                bool ok = true;
                for (int r = 0; r < n_ref; r++) { if (buf.size() < off + 4) { ok = false; break; } int32_t l; memcpy(&l, buf.data() + off, 4); off += 4 + (size_t)l + 4; }
                if (!ok || buf.size() < off) { if (more) continue; die("truncated BAM header"); }
                if (pass == 0) header.assign(buf.begin(), buf.begin() + off);
                have_header = true; start = off;
            }
            offs.resize((buf.size() - start) / 36 + 2);
            uint64_t n = 0, consumed = 0;
            check(umigpu_bam_record_offsets(buf.data() + start, buf.size() - start, offs.data(), offs.size() - 1, &n, &consumed), nullptr, "umigpu_bam_record_offsets");
            const uint8_t *recs = buf.data() + start;
            if (pass == 0) {
                for (uint64_t i = 0; i < n && umi_len == 0; i++) {             // -u 0: autodetect (utils/read.rs:65-75)
                    const uint8_t *r = recs + offs[i];
                    if (rec_passes(r, a)) umi_len = detect_umi_len(r, a.umi_separator);
                }
                if (n && umi_len) {
                    if (!ctx) ctx = make_ctx(a, umi_len);
                    uint64_t nun = 0;
                    check(umigpu_push_bam_records(ctx, n, recs, offs.data(), a.umi_separator, index, &nun), ctx, "umigpu_push_bam_records");
                    unmapped += nun;
                } else {                    // no read has passed the filters yet: count the batch on the host
                    for (uint64_t i = 0; i < n; i++) count_unfed(recs + offs[i], a, unfed, unmapped);
                }
            } else if (pass < last_pass) {  // --paired: mate keys of the kept reads
                for (uint64_t i = 0; i < n; i++)
                    if (kpos < res.n_kept && res.kept_read_index[kpos] == index + i) { kpos++; mates.remember(recs + offs[i]); }
            } else {
                for (uint64_t i = 0; i < n; i++) {
                    const uint8_t *r = recs + offs[i];
                    bool keep = kpos < res.n_kept && res.kept_read_index[kpos] == index + i;
                    if (keep) kpos++;
                    else if (a.keep_unmapped && ((r[18] | (r[19] << 8)) & 4)) keep = true;
                    else if (a.paired && mates.take(r)) keep = true;
                    if (keep) out->add(r, offs[i + 1] - offs[i]);
                }
            }
            index += n;
            buf.erase(buf.begin(), buf.begin() + start + consumed);         // a trailing partial record stays for the next batch
        }
        if (pass == last_pass) {
            out->close(); delete out;
            umigpu_counters ctr; memset(&ctr, 0, sizeof ctr);
            if (ctx) ctr = res.counters;
            ctr.total_reads += unfed.total_reads; ctr.n_unpaired += unfed.n_unpaired; ctr.n_chimeric += unfed.n_chimeric;
            report(ctr, unmapped, a.paired);
            (void)ctx;    // no umigpu_destroy: see run_bam
        }
    }
    return 0;
}

// ---- --mode fastq: one global bucket (BASELINE config 4); the reference's fastq arm is an empty TODO (main.rs:49-51) ----
static int run_fastq(const Cli &a) {
    std::vector<uint8_t> raw = read_file(a.input);
    std::vector<uint8_t> buf;
    if (raw.size() > 2 && raw[0] == 0x1f && raw[1] == 0x8b) {           // gzip'ed FASTQ
        gzFile g = gzopen(a.input.c_str(), "rb");
        if (!g) die("Invalid input path");
        std::vector<uint8_t> tmp(1 << 22); int got;
        while ((got = gzread(g, tmp.data(), (unsigned)tmp.size())) > 0) buf.insert(buf.end(), tmp.begin(), tmp.begin() + got);
        gzclose(g);
    } else buf.swap(raw);
    // record = 4 lines: @header, sequence, +, quality
    std::vector<size_t> rec_start, hdr_end, qual_start, qual_end;
    size_t p = 0, N = buf.size();
    auto eol = [&](size_t s) { const void *q = memchr(buf.data() + s, '\n', N - s); return q ? (size_t)((const uint8_t *)q - buf.data()) : N; };
    while (p < N) {
        size_t e1 = eol(p); if (e1 >= N) break; size_t e2 = eol(e1 + 1); if (e2 >= N) break; size_t e3 = eol(e2 + 1); if (e3 >= N) break; size_t e4 = eol(e3 + 1);
        if (buf[p] != '@') die("malformed FASTQ record");
        rec_start.push_back(p); hdr_end.push_back(e1); qual_start.push_back(e3 + 1); qual_end.push_back(e4);
        p = e4 + 1;
    }
    const size_t n = rec_start.size();
    unsigned umi_len = a.umi_length;
    if (n && !umi_len) {
        size_t s = rec_start[0] + 1, e = hdr_end[0], q = s;
        for (;; q++) {
            while (q < e && buf[q] != a.umi_separator) q++;
            if (q >= e) die("failed to get the umi");
            if (q + 1 < e && buf[q + 1] && strchr("ACGTNacgtn", buf[q + 1])) break;
        }
        for (size_t t = q + 1; t < e && buf[t] && strchr("ACGTNacgtn", buf[t]); t++) umi_len++;
    }
    std::vector<const uint8_t *> optr; std::vector<size_t> olen;
    umigpu_counters ctr; memset(&ctr, 0, sizeof ctr);
    if (n && umi_len) {
        std::vector<int32_t> tid(n, 0), score(n); std::vector<int64_t> pos(n, 0); std::vector<uint8_t> rev(n, 0), umi(n * (size_t)umi_len);
        parallel_for(n, a.num_threads, [&](size_t i) {
            size_t s = rec_start[i] + 1, e = hdr_end[i], q = s;
            while (q < e && buf[q] != a.umi_separator) q++;
            if (q >= e) die("failed to get the umi");
            if (q + 1 + umi_len > e) die("read name too short for the UMI");
            memcpy(umi.data() + i * umi_len, buf.data() + q + 1, umi_len);
            // avg_qual with the reference's f32 formula (utils/read.rs:56-63) on phred = byte - 33
            size_t ql = qual_end[i] - qual_start[i]; if (ql && buf[qual_end[i] - 1] == '\r') ql--;
            float sum = 0.0f; for (size_t t = 0; t < ql; t++) sum += (float)(buf[qual_start[i] + t] - 33);
            float qv = sum / (float)ql; score[i] = (qv != qv) ? 0 : (int32_t)qv;
        });
        umigpu_ctx *ctx = make_ctx(a, umi_len);
        check(umigpu_push_reads(ctx, n, tid.data(), pos.data(), rev.data(), umi.data(), score.data(), nullptr, 0), ctx, "umigpu_push_reads");
        umigpu_result res;
        check(umigpu_finish(ctx, &res), ctx, "umigpu_finish");
        ctr = res.counters;
        for (uint64_t j = 0; j < res.n_kept; j++) { size_t i = (size_t)res.kept_read_index[j]; optr.push_back(buf.data() + rec_start[i]); olen.push_back(std::min(N, qual_end[i] + 1) - rec_start[i]); }
        FILE *f = fopen(a.output.c_str(), "wb"); if (!f) die("cannot open output " + a.output);
        for (size_t j = 0; j < optr.size(); j++) fwrite(optr[j], 1, olen[j], f);
        fclose(f);
        (void)ctx;    // no umigpu_destroy: the process exits right after the output is closed (freeing pinned buffers and the CUDA context costs seconds)
    } else { FILE *f = fopen(a.output.c_str(), "wb"); if (f) fclose(f); ctr.total_reads = n; }
    report(ctr, 0);
    return 0;
}

int main(int argc, char **argv) {
    Cli a = parse(argc, argv);
    auto t0 = std::chrono::steady_clock::now();
    if (a.paired && a.mode == "fastq") die("--paired applies to --mode bam only");
    if (a.paired && a.track_clusters) die("--tag with --paired is not implemented (the reference's tag pass is an empty TODO, deduplicate_sam.rs:236-239)");
    if (a.track_clusters && a.mode == "fastq") die("--tag is implemented for --mode bam only");
    // CUDA start-up takes seconds on a large host: do it on a helper thread while the input is read and inflated
    std::thread cuda_init([dev = a.device] { umigpu_device_init(dev); });
    g_cuda_init = &cuda_init;
    int rc;
    struct Joiner { std::thread &t; ~Joiner() { if (t.joinable()) t.join(); g_cuda_init = nullptr; } } joiner{cuda_init};
    if (a.mode == "fastq") rc = run_fastq(a);
    else if (a.mode == "bam" || a.mode == "sam") rc = (a.two_pass && !a.track_clusters) ? run_bam_two_pass(a) : run_bam(a);
    else die("unknown mode " + a.mode);
    fprintf(stderr, "UMI collapsing finished in %.3f seconds\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());   // main.rs:97-102
    // the output is closed: skip the CUDA context teardown and the release of GBs of host buffers
    if (cuda_init.joinable()) cuda_init.join();
    g_cuda_init = nullptr;
    fflush(stdout); fflush(stderr);
    _exit(rc);
}
