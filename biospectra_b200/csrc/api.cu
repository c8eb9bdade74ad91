// api.cu -- C ABI of libbiospectra_gpu.so (see include/biospectra_gpu.h).
#include "../../include/biospectra_gpu.h"
#include "gpu_index.cuh"
#include "filter3.cuh"

#include <dirent.h>
#include <errno.h>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <sys/stat.h>
#include <unistd.h>

#define BSP_API extern "C" __attribute__((visibility("default")))

static thread_local std::string g_last_error;

static const char* IDX_FILE = "biospectra.idx";
static const u64 IDX_MAGIC = 0x3130584449505342ull;   // "BSPIDX01"

// ------------------------------------------------------------------ handles
struct bsp_indexer {
    std::mutex mu;
    bsp_config conf;
    std::string index_path;
    RefStore* ref = nullptr;
    cudaStream_t st = nullptr;
    std::string last_error;
    ~bsp_indexer() { delete ref; if (st) cudaStreamDestroy(st); }
};

// page-locked host buffer that only grows (copies from / to it are truly asynchronous)
struct PinBuf {
    u8* p = nullptr; size_t bytes = 0;
    PinBuf() {}
    PinBuf(const PinBuf&) = delete;
    PinBuf& operator=(const PinBuf&) = delete;
    ~PinBuf() { if (p) cudaFreeHost(p); }
    void ensure(size_t need) {
        if (need <= bytes) return;
        if (p) { cudaFreeHost(p); p = nullptr; bytes = 0; }
        const size_t want = need + need / 4 + 256;
        CUDA_CHECK(cudaMallocHost((void**)&p, want));
        bytes = want;
    }
};

// Per pipeline slot: device scratch of one sub-batch, its stream, events and pinned host mailboxes.
struct Scratch {
    DevBuf<u8> seq, scan_tmp; DevBuf<i64> seq_off, tok_off, tok_cap;
    DevBuf<u64> tok_key; DevBuf<u32> tok_pos, tok_df, clause_row; DevBuf<float> clause_val;
    DevBuf<i32> n_tok, n_clauses, msm, route, status_in, list_dense, list_pos, list_filter, counts, read_ngap;
    DevBuf<FilterStats> fstats; DevBuf<float2> vrange;
    // gap clauses of the sub-batch (gapjoin.cuh): the clauses, counters ([0] clauses registered, [1] matches appended,
    // [2] chunks), per-clause offsets of the sorted match lists, and the join's working arrays
    DevBuf<GapClause> gaps; DevBuf<u32> gap_ctr, gap_off;
    i64 gc = 0, gc_min = 0;             // capacity of `gaps`; what the last overflow asked for
    DevBuf<u32> gj_keys, gj_keys_alt, gj_vals, gj_vals_alt, gj_mg, gj_mcol, gj_v, gj_v_alt; DevBuf<uint2> gj_chunks;
    DevBuf<float> gj_mfreq; DevBuf<u64> gj_k64, gj_k64_alt; DevBuf<unsigned long long> gj_bytes;
    RadixSortTemp gj_rst; DevBuf<u32> gj_sizes; DevBuf<GjCand> gj_cand;
    const u8* in_seq = nullptr; const i64* in_off = nullptr; i64 in_bytes = 0;   // phase 1 inputs (re-run when `gaps` was too short)
    DevBuf<i32> pos_slot;               // read -> index in list_pos (the positional scorer's partial lists are per list entry)
    DevBuf<i32> ppart_n; DevBuf<u32> ppart_doc; DevBuf<float> ppart_score;
    int slot_index = 0;
    bool front_fused = false;
    // filter3.cuh: per-read headers / clause records / exception cells, candidate lists, read-wide thresholds
    DevBuf<F3Hdr> f3_hdr; DevBuf<uint4> f3_rec; DevBuf<uint2> f3_rec_hi, f3_cand; DevBuf<F3Cell> f3_cells; DevBuf<F3Long> f3_longs;
    DevBuf<i32> f3_cand_n, f3_long_over; DevBuf<u32> f3_tau, f3_prolog;
    bool filter_v3 = false;
    DevBuf<i32> part_n; DevBuf<u32> part_doc; DevBuf<float> part_score;
    DevBuf<i32> o_status, o_label, o_nhits, o_ntop, o_doc; DevBuf<float> o_score;
    DevBuf<i32> o_taxlabel, o_taxidx;
    cudaStream_t st = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    u8* pin_in = nullptr; size_t pin_in_bytes = 0;
    PinBuf pin_routes;                  // routes of the sub-batch on their way to bsp_last_routes
    bool routes_pending = false;
    i32* h_counts = nullptr;            // pinned: [0..3] route counts, [4] filter overflow, [5] gap clauses, [6] most clauses, [7..9] gap join
    FilterStats* h_fstats = nullptr;    // pinned
    // state between the two phases of a sub-batch
    i32 n = 0; int n_tiles = 1, parts_pos = 0, parts_stride = 1, dense_threads = 32, filter_threads = 32, filter_tiles = 1;
    QueryParams qp; bool busy = false;
    i32 r0 = 0;                         // first read of the sub-batch in the caller's arrays
    i32 cnt[4] = {0, 0, 0, 0};          // route counts between the two halves of phase 2: dense, positional, filter, overflow
    i64 launches = 0, launches_gap = 0; bool filter_launched = false;
    void init() {
        CUDA_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        for (auto& e : ev) CUDA_CHECK(cudaEventCreate(&e));
        CUDA_CHECK(cudaMallocHost((void**)&h_counts, 16 * sizeof(i32)));
        CUDA_CHECK(cudaMallocHost((void**)&h_fstats, sizeof(FilterStats)));
    }
    ~Scratch() {
        if (pin_in) cudaFreeHost(pin_in);
        if (h_counts) cudaFreeHost(h_counts);
        if (h_fstats) cudaFreeHost(h_fstats);
        for (auto& e : ev) if (e) cudaEventDestroy(e);
        if (st) cudaStreamDestroy(st);
    }
};

// bsp_classify_one: one pending request per blocked caller
struct OneRead {
    const char* seq; i32 len;
    i32 status = BSP_READ_ERROR, label = 0, n_hits = 0, n_top = 0;
    i32 doc[TOPN]; float score[TOPN];
    int rc = BSP_OK; std::string err;
    bool done = false;
};
// Leader/follower micro-batcher: the first caller of a batch collects company, runs the GPU batch for everybody and
// hands the results back; no background thread.
struct MicroBatcher {
    std::mutex mu;
    std::condition_variable cv_collect, cv_done;
    std::vector<OneRead*> pending;
    bool leader = false;
    int gpu_busy = 0;                       // micro-batches that left the queue and have not finished
    std::mutex run_mu;                      // one micro-batch at a time owns the pinned staging below
    PinBuf pin_seq, pin_off, pin_i32, pin_f32;
    i32 max_reads = 4096;
    i64 wait_us = 200;
    i32 last_size = 0;                      // size of the previous micro-batch
    i64 batches = 0, reads = 0;
};

struct bsp_classifier {
    std::mutex mu;
    MicroBatcher mb;
    bsp_config conf;
    RefStore* ref = nullptr;
    GpuIndex ix;
    cudaStream_t st = nullptr;          // build / introspection stream (= slot 0's stream)
    std::string last_error;
    Scratch sc[BSP_SLOTS];              // sub-batches in flight: the next one's scorer is queued before the current one ends
    // scoring constants
    DevBuf<float> idf_by_df, doc_w, tf16;
    // dense-table side of the per-doc weights: by table column (GpuIndex::docmap), the norm classes present and, per group of
    // 32 columns, the class with the smallest weight (BM25 filter: Bm25Classes)
    DevBuf<float> col_w, class_k; DevBuf<u8> group_class; int n_classes = 0; float k_min = 0.0f, k_max = 0.0f;
    DevBuf<float2> group_w;             // per group of 32 columns: (smallest, largest) per-doc weight
    int sm_count = 148;
    // interned ranked lineages (bsp_classifier_set_lineages)
    DevBuf<i64> lin_off; DevBuf<i32> lin_tax; DevBuf<u8> lin_flags; i64 lin_docs = 0;
    i64 global_docs = 0, global_sum_ttf = 0, doc_base_global = 0;
    // last batch
    std::vector<i32> last_routes;
    i64 counters[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    double score_ms = 0, pipe_ms = 0;
    // gap join of the last batch: clauses, chunks, matches, position bytes fetched, kernel time; tail entries wanted next time
    i64 gj_clauses = 0, gj_chunks = 0, gj_matches = 0, gj_bytes = 0, gj_cands = 0; double gj_ms = 0, gj_verify_ms = 0; size_t gap_tail_want = 0;
    ~bsp_classifier() { delete ref; }
};

// Error text is kept per calling thread only (bsp_last_error reads it back on the same thread): entry points such as
// bsp_classify_one are called concurrently on one handle, so a per-handle string would be written by several threads.
static int fail(std::string* /*unused: per-handle slot*/, int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}

#define API_TRY(h) try {
#define API_CATCH(h)                                                                           \
    } catch (const BspError& e) { return fail((h) ? &(h)->last_error : nullptr, e.code, e.what()); } \
    catch (const std::bad_alloc&) { return fail((h) ? &(h)->last_error : nullptr, BSP_ERR_NOMEM, "out of host memory"); } \
    catch (const std::exception& e) { return fail((h) ? &(h)->last_error : nullptr, BSP_ERR_STATE, e.what()); }

BSP_API const char* bsp_last_error(const void* handle) {
    (void)handle;                       // the message of the last failing call made by THIS thread
    return g_last_error.c_str();
}
BSP_API const char* bsp_version(void) { return "biospectra-b200 0.1 (sm_100a)"; }

BSP_API void bsp_config_default(bsp_config* c) {         // Configuration.java:37-44
    memset(c, 0, sizeof *c);
    c->kmer_size = 10; c->kmer_skips = 5; c->min_strand_kmer = 0;
    c->query_algorithm = BSP_PAIRED_PROXIMITY; c->scoring_algorithm = BSP_TFIDF;
    c->worker_threads = 4; c->index_ram_buffer = 16; c->device = 0;
    c->query_term_min_should_match = 0.5;
    c->part_rank = 0; c->part_count = 1; c->dense_tables = 0; c->positional = 0; c->full_topn = 0;
}

static i64 env_i64(const char* name, i64 dflt) {
    const char* v = getenv(name);
    if (!v || !*v) return dflt;
    return atoll(v);
}

// ------------------------------------------------------------------ filesystem
static void wipe_directory(const std::string& path) {      // Indexer.java:253-264 cleanUpDirectory
    DIR* d = opendir(path.c_str());
    if (!d) return;
    struct dirent* de;
    while ((de = readdir(d)) != nullptr) {
        if (!strcmp(de->d_name, ".") || !strcmp(de->d_name, "..")) continue;
        std::string p = path + "/" + de->d_name;
        struct stat sb;
        if (lstat(p.c_str(), &sb) != 0) continue;
        if (S_ISDIR(sb.st_mode)) { wipe_directory(p); rmdir(p.c_str()); }
        else unlink(p.c_str());
    }
    closedir(d);
}
static void mkdirs(const std::string& path) {
    std::string cur;
    for (size_t i = 0; i <= path.size(); i++) {
        if (i == path.size() || path[i] == '/') {
            if (!cur.empty()) mkdir(cur.c_str(), 0777);
        }
        if (i < path.size()) cur.push_back(path[i]);
    }
}
static bool is_dir(const std::string& p) {
    struct stat sb;
    return stat(p.c_str(), &sb) == 0 && S_ISDIR(sb.st_mode);
}

// ------------------------------------------------------------------ indexer
BSP_API int bsp_index_create(const bsp_config* conf, const char* index_path, bsp_indexer** out) {
    if (out) *out = nullptr;
    bsp_indexer* h = nullptr;
    API_TRY(h)
    if (!conf) return fail(nullptr, BSP_ERR_ARG, "conf is null");                                   // Indexer.java:65-67
    if (!index_path) return fail(nullptr, BSP_ERR_ARG, "indexPath is null");                        // :69-71
    if (conf->kmer_size <= 0) return fail(nullptr, BSP_ERR_ARG, "kmerSize must be larger than 0");  // :73-75
    if (conf->worker_threads <= 0) return fail(nullptr, BSP_ERR_ARG, "workerThreads must be larger than 0"); // :77-79
    if (conf->kmer_size > 31) return fail(nullptr, BSP_ERR_UNSUPPORTED, "kmerSize > 31 is not supported");
    if (!out) return fail(nullptr, BSP_ERR_ARG, "out is null");
    CUDA_CHECK(cudaSetDevice(conf->device));
    h = new bsp_indexer();
    h->conf = *conf;
    h->index_path = index_path;
    if (*index_path) {
        mkdirs(h->index_path);                                                                      // :85-87
        if (is_dir(h->index_path)) wipe_directory(h->index_path);                                   // :89-91
    }
    CUDA_CHECK(cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking));
    h->ref = new RefStore();
    h->ref->k = conf->kmer_size;
    h->ref->min_strand = conf->min_strand_kmer ? 1 : 0;
    *out = h;
    return BSP_OK;
    } catch (const BspError& e) { delete h; return fail(nullptr, e.code, e.what()); }
    catch (const std::exception& e) { delete h; return fail(nullptr, BSP_ERR_STATE, e.what()); }
}

static int index_add(bsp_indexer* h, const char* filename, const char* header, const char* taxonomy,
                     const void* seq, i64 seq_len, bool on_device) {
    if (!h) return fail(nullptr, BSP_ERR_ARG, "handle is null");
    std::lock_guard<std::mutex> lk(h->mu);
    API_TRY(h)
    if (!h->ref) return fail(&h->last_error, BSP_ERR_STATE, "indexer is closed");
    if (seq_len < 0 || (!seq && seq_len > 0)) return fail(&h->last_error, BSP_ERR_ARG, "sequence is null");
    CUDA_CHECK(cudaSetDevice(h->conf.device));
    EntryMeta m;
    m.filename = filename ? filename : "";
    m.header = header ? header : "";
    m.taxonomy = taxonomy ? taxonomy : "";
    m.len = seq_len;
    int rev_ok = on_device ? h->ref->append_device((const u8*)seq, seq_len, h->st)
                           : h->ref->append_host((const char*)seq, seq_len, h->st);
    h->ref->add(std::move(m), rev_ok);
    return BSP_OK;
    API_CATCH(h)
}

BSP_API int bsp_index_add_entry(bsp_indexer* h, const char* filename, const char* header, const char* taxonomy,
                                const char* seq, int64_t seq_len) {
    return index_add(h, filename, header, taxonomy, seq, seq_len, false);
}
BSP_API int bsp_index_add_entry_device(bsp_indexer* h, const char* filename, const char* header, const char* taxonomy,
                                       const void* d_seq, int64_t seq_len) {
    return index_add(h, filename, header, taxonomy, d_seq, seq_len, true);
}

// ---- index file: header, entry table, strings, packed words, mask words
static void wr(FILE* f, const void* p, size_t n) {
    if (n && fwrite(p, 1, n, f) != n) throw BspError(BSP_ERR_IO, std::string("write failed: ") + strerror(errno));
}
static void rd(FILE* f, void* p, size_t n) {
    if (n && fread(p, 1, n, f) != n) throw BspError(BSP_ERR_IO, "index file truncated");
}
static void wr_str(FILE* f, const std::string& s) { u64 n = s.size(); wr(f, &n, 8); wr(f, s.data(), s.size()); }
static std::string rd_str(FILE* f) {
    u64 n; rd(f, &n, 8);
    if (n > ((u64)1 << 32)) throw BspError(BSP_ERR_IO, "corrupt index file (string length)");
    std::string s(n, '\0'); rd(f, &s[0], n); return s;
}

// Two page-locked bounce buffers for the index file: the copy of one chunk to / from the device overlaps the file I/O of
// the other (truly asynchronous copies need page-locked memory; the round-1 code went through one pageable buffer with a
// stream synchronize per 64 MB).
struct IoRing {
    static constexpr size_t CH = (size_t)16 << 20;      // words per chunk
    u32* buf[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    IoRing() {
        for (int i = 0; i < 2; i++) {
            CUDA_CHECK(cudaMallocHost((void**)&buf[i], CH * 4));
            CUDA_CHECK(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
        }
    }
    ~IoRing() { for (int i = 0; i < 2; i++) { if (buf[i]) cudaFreeHost(buf[i]); if (ev[i]) cudaEventDestroy(ev[i]); } }
};

static void write_index_file(const std::string& dir, RefStore& ref, cudaStream_t st) {
    std::string tmp = dir + "/" + IDX_FILE + ".tmp", fin = dir + "/" + IDX_FILE;
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f) throw BspError(BSP_ERR_IO, "cannot create " + tmp + ": " + strerror(errno));
    try {
        u64 hdr[8] = {IDX_MAGIC, 1, (u64)ref.k, (u64)ref.min_strand, (u64)ref.entries.size(), (u64)ref.used_bases, 0, 0};
        wr(f, hdr, sizeof hdr);
        for (const EntryMeta& m : ref.entries) {
            i64 v[3] = {m.len, m.base_off, (i64)m.rev_ok};
            wr(f, v, sizeof v);
            wr_str(f, m.filename); wr_str(f, m.header); wr_str(f, m.taxonomy);
        }
        IoRing ring;
        const size_t CH = IoRing::CH;
        for (int which = 0; which < 2; which++) {
            const u32* src = which == 0 ? ref.packed.p : ref.mask.p;
            size_t words = (size_t)(which == 0 ? ref.used_bases / 16 : ref.used_bases / 32);
            const size_t chunks = (words + CH - 1) / CH;
            // chunk c travels to the host while chunk c - 1 is written to the file
            for (size_t c = 0; c <= chunks; c++) {
                if (c < chunks) {
                    const size_t o = c * CH, n = std::min(CH, words - o);
                    CUDA_CHECK(cudaMemcpyAsync(ring.buf[c & 1], src + o, n * 4, cudaMemcpyDeviceToHost, st));
                    CUDA_CHECK(cudaEventRecord(ring.ev[c & 1], st));
                }
                if (c > 0) {
                    const size_t o = (c - 1) * CH, n = std::min(CH, words - o);
                    CUDA_CHECK(cudaEventSynchronize(ring.ev[(c - 1) & 1]));
                    wr(f, ring.buf[(c - 1) & 1], n * 4);
                }
            }
        }
        if (fflush(f) != 0) throw BspError(BSP_ERR_IO, "flush failed");
        fsync(fileno(f));
        fclose(f); f = nullptr;
        if (rename(tmp.c_str(), fin.c_str()) != 0) throw BspError(BSP_ERR_IO, "rename failed: " + std::string(strerror(errno)));
    } catch (...) {
        if (f) fclose(f);
        unlink(tmp.c_str());
        throw;
    }
}

static RefStore* read_index_file(const std::string& dir, int part_rank, int part_count, cudaStream_t st) {
    std::string fin = dir + "/" + IDX_FILE;
    FILE* f = fopen(fin.c_str(), "rb");
    if (!f) throw BspError(BSP_ERR_IO, "cannot open " + fin + ": " + strerror(errno));
    RefStore* ref = new RefStore();
    try {
        u64 hdr[8];
        rd(f, hdr, sizeof hdr);
        if (hdr[0] != IDX_MAGIC || hdr[1] != 1) throw BspError(BSP_ERR_IO, "not a biospectra-b200 index (bad magic/version)");
        ref->k = (int)hdr[2]; ref->min_strand = (int)hdr[3];
        u64 nE = hdr[4]; i64 used = (i64)hdr[5];
        std::vector<EntryMeta> all(nE);
        for (u64 e = 0; e < nE; e++) {
            i64 v[3]; rd(f, v, sizeof v);
            all[e].len = v[0]; all[e].base_off = v[1]; all[e].rev_ok = (int)v[2];
            all[e].filename = rd_str(f); all[e].header = rd_str(f); all[e].taxonomy = rd_str(f);
        }
        long data_pos = ftell(f);
        // partition: contiguous entry ranges balanced by bases
        u64 e0 = 0, e1 = nE;
        if (part_count > 1) {
            i64 total = 0;
            for (auto& m : all) total += m.len;
            i64 lo = total * part_rank / part_count, hi = total * (part_rank + 1) / part_count;
            i64 run = 0; e0 = nE; e1 = nE;
            bool s0 = false;
            for (u64 e = 0; e < nE; e++) {
                i64 mid = run + all[e].len / 2;
                if (!s0 && mid >= lo) { e0 = e; s0 = true; }
                if (mid >= hi) { e1 = e; break; }
                run += all[e].len;
            }
            if (!s0) e0 = nE;
            if (part_rank == part_count - 1) e1 = nE;
            if (e1 < e0) e1 = e0;
        }
        i64 b0 = e0 < nE ? all[e0].base_off : used;
        i64 b1 = e1 < nE ? all[e1].base_off : used;
        i64 span = b1 - b0;
        ref->reserve_bases(span + 128, st);
        IoRing ring;
        const size_t CH = IoRing::CH;
        for (int which = 0; which < 2; which++) {
            u32* dst = which == 0 ? ref->packed.p : ref->mask.p;
            i64 per = which == 0 ? 16 : 32;
            i64 file_base = data_pos + (which == 0 ? 0 : (used / 16) * 4);
            size_t words = (size_t)(span / per);
            if (fseek(f, file_base + (b0 / per) * 4, SEEK_SET) != 0) throw BspError(BSP_ERR_IO, "seek failed");
            // chunk c is read from the file while chunk c - 1 travels to the device
            size_t c = 0;
            for (size_t o = 0; o < words; o += CH, c++) {
                size_t n = std::min(CH, words - o);
                if (c >= 2) CUDA_CHECK(cudaEventSynchronize(ring.ev[c & 1]));       // the buffer's previous copy has left it
                rd(f, ring.buf[c & 1], n * 4);
                CUDA_CHECK(cudaMemcpyAsync(dst + o, ring.buf[c & 1], n * 4, cudaMemcpyHostToDevice, st));
                CUDA_CHECK(cudaEventRecord(ring.ev[c & 1], st));
            }
            CUDA_CHECK(cudaStreamSynchronize(st));
        }
        for (u64 e = e0; e < e1; e++) {
            EntryMeta m = std::move(all[e]);
            m.base_off -= b0;
            ref->entries.push_back(std::move(m));
        }
        ref->used_bases = span;
        fclose(f);
    } catch (...) {
        fclose(f);
        delete ref;
        throw;
    }
    return ref;
}

// ------------------------------------------------------------------ classifier setup
static void upload_scoring_constants(bsp_classifier* h) {
    GpuIndex& ix = h->ix;
    const i64 D = (i64)ix.docs.size();
    const i64 maxdoc = h->global_docs;
    // idf by df (host libm log: the same function the oracle uses)
    std::vector<float> idf(maxdoc + 2);
    for (i64 df = 0; df <= maxdoc + 1; df++) {
        if (h->conf.scoring_algorithm == BSP_BM25)
            idf[df] = (float)std::log(1.0 + ((double)(maxdoc - df) + 0.5) / ((double)df + 0.5));      // BM25Similarity#idf
        else
            idf[df] = (float)(std::log((double)maxdoc / (double)(df + 1)) + 1.0);                     // DefaultSimilarity#idf
    }
    h->idf_by_df.alloc(idf.size());
    CUDA_CHECK(cudaMemcpyAsync(h->idf_by_df.p, idf.data(), idf.size() * 4, cudaMemcpyHostToDevice, h->st));
    // per-doc weight
    float cache[256];
    if (h->conf.scoring_algorithm == BSP_BM25) {
        volatile float avgdl = h->global_sum_ttf <= 0 ? 1.0f : (float)((double)h->global_sum_ttf / (double)maxdoc);
        const float k1 = 1.2f, b = 0.75f;
        for (int x = 0; x < 256; x++) {
            volatile float f = h_byte315_to_float(x);
            volatile float ff = f * f;
            volatile float ln = 1.0f / ff;
            volatile float t1 = b * ln;
            volatile float t2 = t1 / avgdl;
            volatile float t3 = (1.0f - b) + t2;
            cache[x] = k1 * t3;
        }
    } else {
        for (int x = 0; x < 256; x++) cache[x] = h_byte315_to_float(x);
    }
    std::vector<float> dw(round_up(std::max<i64>(D, 1), 32 * 32) + 32, 1.0f);
    for (i64 d = 0; d < D; d++) dw[d] = cache[ix.doc_norm[d]];
    h->doc_w.alloc(dw.size());
    CUDA_CHECK(cudaMemcpyAsync(h->doc_w.p, dw.data(), dw.size() * 4, cudaMemcpyHostToDevice, h->st));
    // the same weights by table column, the norm classes in column order (columns are sorted by norm byte)
    std::vector<float> cw(ix.h_docmap.size() + 32 * 32, 1.0f), ck;
    std::vector<u8> gcls(ix.h_docmap.size() / 32 + 33, 0);
    int last_byte = -1;
    for (size_t c = 0; c < ix.h_docmap.size(); c++) {
        const u32 d = ix.h_docmap[c];
        if (d == 0xFFFFFFFFu) continue;
        const int nb = ix.doc_norm[d];
        cw[c] = cache[nb];
        if (nb != last_byte) { ck.push_back(cache[nb]); last_byte = nb; }
        // class of the group = the one with the smallest weight among its columns (BM25: an upper bound for the others)
        const size_t g = c / 32;
        const int cls = std::min((int)ck.size() - 1, 127);
        if (c % 32 == 0) gcls[g] = (u8)cls;
        else if ((gcls[g] & 0x7F) != cls) gcls[g] = (u8)(0x80 | (ck[cls] < ck[gcls[g] & 0x7F] ? cls : (gcls[g] & 0x7F)));
    }
    // groups with padding columns (the last one) and classes beyond the 7-bit index are "mixed": per-doc weights are loaded
    for (size_t g = 0; g < gcls.size(); g++) {
        bool pad = false;
        for (size_t c = g * 32; c < g * 32 + 32; c++) pad |= c >= ix.h_docmap.size() || ix.h_docmap[c] == 0xFFFFFFFFu;
        if (pad || (int)ck.size() > 128 || env_i64("BSP_NO_UNIFORM", 0) != 0) gcls[g] |= 0x80;
    }
    if (ck.empty()) ck.push_back(1.0f);
    h->n_classes = (int)ck.size();
    h->k_min = *std::min_element(ck.begin(), ck.end());
    h->k_max = *std::max_element(ck.begin(), ck.end());
    std::vector<float2> gw(gcls.size() + 512, make_float2(1.0f, 1.0f));
    for (size_t g = 0; g * 32 < ix.h_docmap.size(); g++) {
        float lo = 3.0e38f, hi = 0.0f;
        for (size_t c = g * 32; c < g * 32 + 32 && c < ix.h_docmap.size(); c++)
            if (ix.h_docmap[c] != 0xFFFFFFFFu) { lo = std::min(lo, cw[c]); hi = std::max(hi, cw[c]); }
        if (hi > 0.0f || lo < 3.0e38f) gw[g] = make_float2(lo, hi);
    }
    h->group_w.alloc(gw.size());
    CUDA_CHECK(cudaMemcpyAsync(h->group_w.p, gw.data(), gw.size() * sizeof(float2), cudaMemcpyHostToDevice, h->st));
    h->col_w.alloc(cw.size()); h->class_k.alloc(ck.size()); h->group_class.alloc(gcls.size());
    CUDA_CHECK(cudaMemcpyAsync(h->col_w.p, cw.data(), cw.size() * 4, cudaMemcpyHostToDevice, h->st));
    CUDA_CHECK(cudaMemcpyAsync(h->class_k.p, ck.data(), ck.size() * 4, cudaMemcpyHostToDevice, h->st));
    CUDA_CHECK(cudaMemcpyAsync(h->group_class.p, gcls.data(), gcls.size(), cudaMemcpyHostToDevice, h->st));
    // nibble -> per-clause frequency factor: tf-idf tf(freq) = (float)sqrt(freq) (DefaultSimilarity#tf); BM25: freq
    float tl[16];
    for (int c = 0; c < 16; c++) tl[c] = h->conf.scoring_algorithm == BSP_BM25 ? (float)c : (float)std::sqrt((double)c);
    h->tf16.alloc(16);
    CUDA_CHECK(cudaMemcpyAsync(h->tf16.p, tl, sizeof tl, cudaMemcpyHostToDevice, h->st));
    CUDA_CHECK(cudaStreamSynchronize(h->st));
}

static bsp_classifier* make_classifier(const bsp_config* conf, RefStore* ref) {
    bsp_classifier* h = new bsp_classifier();
    try {
        h->conf = *conf;
        h->ref = ref;
        for (int i = 0; i < BSP_SLOTS; i++) { h->sc[i].init(); h->sc[i].slot_index = i; }
        h->st = h->sc[0].st;
        h->mb.max_reads = (i32)std::max<i64>(1, env_i64("BSP_MICROBATCH_MAX", 4096));
        h->mb.wait_us = std::max<i64>(0, env_i64("BSP_MICROBATCH_US", 200));
        // the filter kernel's cp.async ring needs up to 64 KB of dynamic shared memory (512 threads x 8 stages x 16 B);
        // its residency is bounded by shared memory: ask for the largest carve-out
#define FK_ATTR(TI, AT)                                                                                                          \
        CUDA_CHECK(cudaFuncSetAttribute(filter_score_kernel<TI, AT, SIM_TFIDF>, cudaFuncAttributeMaxDynamicSharedMemorySize, FK_STAGES * FK_MAXT * 16)); \
        CUDA_CHECK(cudaFuncSetAttribute(filter_score_kernel<TI, AT, SIM_TFIDF>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)); \
        CUDA_CHECK(cudaFuncSetAttribute(filter_score_kernel<TI, AT, SIM_BM25>, cudaFuncAttributeMaxDynamicSharedMemorySize, FK_STAGES * FK_MAXT * 16 + FK_BM25_LUT_BYTES)); \
        CUDA_CHECK(cudaFuncSetAttribute(filter_score_kernel<TI, AT, SIM_BM25>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared))
        FK_ATTR(false, false); FK_ATTR(false, true); FK_ATTR(true, false); FK_ATTR(true, true);
#undef FK_ATTR
        CUDA_CHECK(cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, conf->device));
        CUDA_CHECK(cudaFuncSetAttribute(filter3_stream_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f3_smem_bytes(FK3_MAXW, false)));
        CUDA_CHECK(cudaFuncSetAttribute(filter3_stream_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f3_smem_bytes(FK3_MAXW, true)));
        CUDA_CHECK(cudaFuncSetAttribute(filter3_stream_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        CUDA_CHECK(cudaFuncSetAttribute(filter3_stream_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        CUDA_CHECK(cudaFuncSetAttribute(score_positional_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(POSW_WARPS * posw_warp_bytes(POSW_DOCS, PW_POS))));
        CUDA_CHECK(cudaFuncSetAttribute(score_positional_warp_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        BuildOptions opt;
        opt.want_tables = conf->dense_tables;
        opt.want_positional = conf->positional;
        opt.need_pair_table = conf->query_algorithm != BSP_NAIVE_KMER;
        // positions are segment ordinals compared as int: a segment never holds more than 2^30 tokens (one longer doc gets its own)
        opt.seg_tokens = std::min<i64>(std::max<i64>(env_i64("BSP_SEG_TOKENS", (i64)1 << 28), 1), (i64)1 << 30);
        opt.exc_cap = std::max<i64>(env_i64("BSP_EXC_CAP", (i64)1 << 26), 1024);
        // (kmer_skips > 0 turns every phrase clause into a gap clause; evaluating them all in the pre-pass costs what the
        // positional scorer costs -- measured at C2 -- so those reads overflow this list on purpose and take that scorer)
        opt.gap_cap = std::max<i64>(env_i64("BSP_GAP_CAP", GAP_CAP_DEFAULT), 1);
        if (opt.seg_tokens < 1024) opt.seg_tokens = 1024;
        build_gpu_index(h->ix, *ref, opt, h->st);
        h->global_docs = (i64)h->ix.docs.size();
        h->global_sum_ttf = h->ix.sum_ttf;
        h->doc_base_global = 0;
        upload_scoring_constants(h);
    } catch (...) {
        h->ref = nullptr;      // caller keeps ownership on failure
        delete h;
        throw;
    }
    return h;
}

static int check_query_conf(const bsp_config* conf) {
    if (!conf) return fail(nullptr, BSP_ERR_ARG, "conf is null");                                          // Classifier.java:72-74
    if (conf->kmer_size <= 0) return fail(nullptr, BSP_ERR_ARG, "kmerSize must be larger than 0");         // :80-82
    if (conf->kmer_skips < 0) return fail(nullptr, BSP_ERR_ARG, "kmerSkips must be equal or larger than 0"); // :84-86
    if (conf->kmer_size > 31) return fail(nullptr, BSP_ERR_UNSUPPORTED, "kmerSize > 31 is not supported");
    if (conf->query_algorithm < 0 || conf->query_algorithm > 2) return fail(nullptr, BSP_ERR_ARG, "unknown query_algorithm");
    return BSP_OK;
}

BSP_API int bsp_index_close(bsp_indexer* h) {
    if (!h) return fail(nullptr, BSP_ERR_ARG, "handle is null");
    int rc = BSP_OK;
    {
        std::lock_guard<std::mutex> lk(h->mu);
        try {
            CUDA_CHECK(cudaSetDevice(h->conf.device));
            if (h->ref && !h->index_path.empty()) write_index_file(h->index_path, *h->ref, h->st);
        } catch (const BspError& e) { rc = fail(nullptr, e.code, e.what()); }
        catch (const std::exception& e) { rc = fail(nullptr, BSP_ERR_STATE, e.what()); }
    }
    delete h;
    return rc;
}

BSP_API int bsp_index_close_to_classifier(bsp_indexer* h, const bsp_config* query_conf, bsp_classifier** out) {
    if (out) *out = nullptr;
    if (!h || !out) return fail(nullptr, BSP_ERR_ARG, "handle/out is null");
    int rc = check_query_conf(query_conf);
    if (rc) return rc;
    if (query_conf->kmer_size != h->conf.kmer_size || (query_conf->min_strand_kmer != 0) != (h->conf.min_strand_kmer != 0))
        return fail(&h->last_error, BSP_ERR_ARG, "query configuration disagrees with the index (kmer_size / min_strand_kmer)");
    try {
        CUDA_CHECK(cudaSetDevice(h->conf.device));
        bsp_config qc = *query_conf;
        qc.device = h->conf.device;
        RefStore* ref = h->ref;
        bsp_classifier* c = make_classifier(&qc, ref);
        h->ref = nullptr;
        delete h;
        *out = c;
        return BSP_OK;
    } catch (const BspError& e) { return fail(&h->last_error, e.code, e.what()); }
    catch (const std::exception& e) { return fail(&h->last_error, BSP_ERR_STATE, e.what()); }
}

BSP_API int bsp_classifier_open(const bsp_config* conf, const char* index_path, bsp_classifier** out) {
    if (out) *out = nullptr;
    int rc = check_query_conf(conf);
    if (rc) return rc;
    if (!index_path) return fail(nullptr, BSP_ERR_ARG, "indexPath is null");                               // Classifier.java:76-78
    if (!out) return fail(nullptr, BSP_ERR_ARG, "out is null");
    if (!is_dir(index_path)) return fail(nullptr, BSP_ERR_ARG, "indexPath is not a directory or does not exist"); // :92-94
    RefStore* ref = nullptr;
    cudaStream_t st = nullptr;
    try {
        CUDA_CHECK(cudaSetDevice(conf->device));
        CUDA_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        ref = read_index_file(index_path, conf->part_rank, conf->part_count > 0 ? conf->part_count : 1, st);
        cudaStreamDestroy(st); st = nullptr;
        if (ref->k != conf->kmer_size || ref->min_strand != (conf->min_strand_kmer ? 1 : 0)) {
            delete ref;
            return fail(nullptr, BSP_ERR_ARG, "configuration disagrees with the index (kmer_size / min_strand_kmer)");
        }
        *out = make_classifier(conf, ref);
        return BSP_OK;
    } catch (const BspError& e) { if (st) cudaStreamDestroy(st); delete ref; return fail(nullptr, e.code, e.what()); }
    catch (const std::exception& e) { if (st) cudaStreamDestroy(st); delete ref; return fail(nullptr, BSP_ERR_STATE, e.what()); }
}

BSP_API int bsp_classifier_close(bsp_classifier* h) {
    if (!h) return fail(nullptr, BSP_ERR_ARG, "handle is null");
    { std::lock_guard<std::mutex> lk(h->mu); cudaSetDevice(h->conf.device); cudaStreamSynchronize(h->st); }
    delete h;
    return BSP_OK;
}

// ------------------------------------------------------------------ classification
// counts: [0] exact-dense reads, [1] positional reads, [2] filter reads, [3] filter overflow (appended to list_dense)
__global__ void __launch_bounds__(256) route_compact_kernel(const i32* __restrict__ route, const i32* __restrict__ n_clauses, i32 n,
                                                             i32* __restrict__ list_dense, i32* __restrict__ list_pos,
                                                             i32* __restrict__ list_filter, i32* __restrict__ counts,
                                                             i32* __restrict__ pos_slot) {
    i32 r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    int rt = route[r];
    if (rt == ROUTE_DENSE) list_dense[atomicAdd(&counts[0], 1)] = r;
    else if (rt == ROUTE_POSITIONAL) { const i32 at = atomicAdd(&counts[1], 1); list_pos[at] = r; pos_slot[r] = at; }
    else if (rt == ROUTE_FILTER) {
        list_filter[atomicAdd(&counts[2], 1)] = r;
        atomicMax(&counts[6], n_clauses[r]);          // most clauses of a filter read: sizes the per-class LUT arrays (BM25)
    }
}

__global__ void __launch_bounds__(256) tok_capacity_kernel(const i64* __restrict__ seq_off, i32 n, int k, i64* __restrict__ cap) {
    i32 r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    i64 len = seq_off[r + 1] - seq_off[r];
    i64 w = len - k + 1;
    cap[r] = w > 0 ? w : 0;
}

struct BatchOut {     // device pointers (any may be null except status)
    i32 *status, *label, *n_hits, *n_top, *top_doc; float* top_score;
};

static bool ensure_gap_directory(bsp_classifier* h, cudaStream_t st);

// ---- the device pipeline of one sub-batch, in two phases so that two sub-batches can be in flight:
//   phase 1 (enqueue only): tokenize -> df -> weights/routing -> route lists; the route counts travel to a pinned mailbox
//   phase 2: wait for the counts, launch the scorers (grid sizes depend on the counts), merge; everything on sc.st
static void pipeline_phase1(bsp_classifier* h, Scratch& sc, i32 n, const u8* d_seq, const i64* d_seq_off, i64 total_bytes) {
    GpuIndex& ix = h->ix;
    cudaStream_t st = sc.st;
    const int k = ix.k;
    sc.in_seq = d_seq; sc.in_off = d_seq_off; sc.in_bytes = total_bytes;
    i64 tok_cap_total = total_bytes;        // sum(max(len-k+1,0)) <= total bytes
    sc.tok_off.ensure((size_t)n + 1);
    sc.tok_key.ensure((size_t)tok_cap_total + 1); sc.tok_pos.ensure((size_t)tok_cap_total + 1);
    sc.tok_df.ensure((size_t)tok_cap_total + 1); sc.clause_val.ensure((size_t)tok_cap_total + 1);
    sc.clause_row.ensure((size_t)tok_cap_total + 1);
    sc.n_tok.ensure(n); sc.n_clauses.ensure(n); sc.msm.ensure(n); sc.route.ensure(n); sc.status_in.ensure(n);
    sc.list_dense.ensure(n); sc.list_pos.ensure(n); sc.list_filter.ensure(n); sc.counts.ensure(8); sc.vrange.ensure(n); sc.read_ngap.ensure(n);
    sc.pos_slot.ensure(n);
    if (!sc.fstats.p) sc.fstats.alloc(1);
    sc.n = n; sc.busy = true;
    CUDA_CHECK(cudaEventRecord(sc.ev[0], st));
    // token capacity offsets
    sc.tok_cap.ensure((size_t)n + 1);
    tok_capacity_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(d_seq_off, n, k, sc.tok_cap.p);
    exclusive_scan<i64, i64>(sc.tok_cap.p, sc.tok_off.p, (u64)n, sc.tok_off.p + n, sc.scan_tmp, st);
    QueryParams& qp = sc.qp;
    qp.k = k; qp.skips = h->conf.kmer_skips; qp.min_strand = ix.min_strand; qp.alg = h->conf.query_algorithm;
    qp.sim = h->conf.scoring_algorithm == BSP_BM25 ? SIM_BM25 : SIM_TFIDF;
    qp.msm_ratio = h->conf.query_term_min_should_match;
    qp.max_doc = h->global_docs;
    qp.dense_ok = ix.tables && (qp.alg == ALG_NAIVE || ix.has_pair) && env_i64("BSP_DISABLE_DENSE", 0) == 0;
    qp.positional_ok = ix.positional ? 1 : 0;
    qp.pair_row0 = ix.pair_row0;
    qp.gap_ok = qp.dense_ok && ix.has_pair && ix.positional && env_i64("BSP_DISABLE_GAP", 0) == 0;
    // (the join's position directory is built the first time a batch may carry gap clauses; without it they take the positional scorer)
    if (qp.gap_ok && qp.alg != ALG_NAIVE && !ensure_gap_directory(h, st)) qp.gap_ok = 0;
    qp.temp_row0 = (u32)ix.n_rows + 1;
    // Gap clauses of the sub-batch.  With kmer_skips > 0 every phrase clause is one (about len / (skips + 1) tokens per read);
    // otherwise only the clauses next to an N are.  A list that turns out too short is grown and phase 1 runs again
    // (pipeline_phase2a).
    i64 gc_want = std::max<i64>(ix.gap_cap, sc.gc_min);
    if (h->conf.kmer_skips > 0) {
        const i64 toks = total_bytes / (h->conf.kmer_skips + 1) + n;
        gc_want = std::max<i64>(gc_want, (qp.alg == ALG_CHAIN ? toks : toks / 2 + n) + 1024);
    }
    gc_want = std::min<i64>(gc_want, (i64)1 << 28);
    if (sc.gc < gc_want) { sc.gc = gc_want; sc.gaps.alloc((size_t)sc.gc); sc.gap_off.alloc((size_t)sc.gc + 2); }
    if (!sc.gap_ctr.p) { sc.gap_ctr.alloc(4); sc.gj_bytes.alloc(1); }
    const u32 GC = (u32)sc.gc;
    qp.gap_cap = GC;
    CUDA_CHECK(cudaMemsetAsync(sc.gap_ctr.p, 0, 16, st));
    // fast filter path: tf-idf; CTA per (doc tile, read), 32 docs per thread, <= 512 threads per tile.
    // BSP_FILTER: 0 off, 1 whenever usable, unset = from 256 documents up (at 600 documents the filter kernel takes 2.1 ms per
    // 200 k reads against 6.5 ms for the all-exact kernel; below a few hundred both are noise)
    const u32 D = (u32)ix.docs.size();
    const i64 filter_mode = env_i64("BSP_FILTER", -1);
    // (the filter kernel addresses rows by 32-bit offsets in 16-byte units: table below 64 GiB)
    const bool tab_addressable = ((u64)ix.n_rows + 2) * ix.row_bytes < (64ull << 30);
    // (BM25: one LUT per norm class of the handle; 32 classes span a 256x range of document lengths.  The per-class LUTs share
    //  FK_BM25_LUT_BYTES of shared memory: more classes leave room for fewer clauses, see filter_nc_max)
    qp.filter_ok = qp.dense_ok && tab_addressable && (qp.sim == SIM_TFIDF || (h->n_classes <= FK_BM25_CLASSES && env_i64("BSP_BM25_FILTER", 1) != 0)) &&
                   filter_mode != 0 && (filter_mode == 1 || D >= 256);
    qp.filter_nc_max = FK_MAX_CLAUSES;
    if (qp.sim == SIM_BM25)
        qp.filter_nc_max = (int)std::min<i64>(FK_MAX_CLAUSES, FK_BM25_LUT_BYTES / ((i64)std::max(h->n_classes, 1) * (qp.alg == ALG_NAIVE ? 24 : 16)) - 2);
    const unsigned wgrid = (unsigned)ceil_div((i64)n * 32, 256);
    sc.front_fused = h->conf.kmer_skips == 0 && ix.direct && env_i64("BSP_FRONT_FUSED", 1) != 0;
    if (sc.front_fused) {
        front_kernel<<<wgrid, 256, 0, st>>>(d_seq, d_seq_off, n, sc.tok_off.p, ix.min_strand, ix.df_dense.p, sc.tok_key.p, sc.tok_pos.p,
                                            sc.tok_df.p, sc.n_tok.p, h->idf_by_df.p, qp, sc.clause_val.p, sc.clause_row.p, sc.n_clauses.p,
                                            sc.msm.p, sc.route.p, sc.status_in.p, sc.vrange.p, sc.gaps.p, sc.gap_ctr.p, sc.read_ngap.p);
    } else {
        tokenize_kernel<<<wgrid, 256, 0, st>>>(d_seq, d_seq_off, n, sc.tok_off.p, k, h->conf.kmer_skips, ix.min_strand, sc.tok_key.p,
                                               sc.tok_pos.p, sc.n_tok.p);
        token_df_kernel<<<wgrid, 256, 0, st>>>(sc.tok_key.p, sc.tok_off.p, sc.n_tok.p, n, ix.direct ? ix.df_dense.p : nullptr, ix.d_segs.p,
                                               ix.positional ? (int)ix.segs.size() : 0, sc.tok_df.p);
        weights_kernel<<<wgrid, 256, 0, st>>>(d_seq, d_seq_off, n, sc.tok_off.p, sc.n_tok.p, sc.tok_key.p, sc.tok_pos.p, sc.tok_df.p,
                                              h->idf_by_df.p, qp, sc.clause_val.p, sc.clause_row.p, sc.n_clauses.p, sc.msm.p, sc.route.p,
                                              sc.status_in.p, sc.vrange.p, sc.gaps.p, sc.gap_ctr.p, sc.read_ngap.p);
    }
    CUDA_CHECK(cudaMemsetAsync(sc.counts.p, 0, 32, st));
    route_compact_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(sc.route.p, sc.n_clauses.p, n, sc.list_dense.p, sc.list_pos.p,
                                                                     sc.list_filter.p, sc.counts.p, sc.pos_slot.p);
    KERNEL_CHECK();
    CUDA_CHECK(cudaMemcpyAsync(sc.h_counts, sc.counts.p, 16, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaMemcpyAsync(sc.h_counts + 6, sc.counts.p + 6, 4, cudaMemcpyDeviceToHost, st));      // most clauses of a filter read
    CUDA_CHECK(cudaMemcpyAsync(sc.h_counts + 5, sc.gap_ctr.p, 4, cudaMemcpyDeviceToHost, st));        // # gap clauses
    // partial top-10 buffers
    sc.dense_threads = (int)std::min<i64>(512, round_up(std::max<i64>(ceil_div(ix.d_pad, EX_DPT), 1), 32));
    sc.n_tiles = (int)std::max<i64>(1, ceil_div(ix.d_pad, (i64)sc.dense_threads * EX_DPT));
    sc.parts_pos = ix.positional ? (int)ix.segs.size() : 0;
    // filter tiles: BSP_FILTER_THREADS (multiple of 32, <= 512) caps the CTA size (tests use it to force several tiles)
    const i64 ft_cap = std::min<i64>(FK_MAXT, std::max<i64>(32, round_up(env_i64("BSP_FILTER_THREADS", 512), 32)));
    sc.filter_threads = (int)std::min<i64>(ft_cap, round_up(std::max<i64>(ceil_div(ix.d_pad, FK_DOCS), 1), 32));
    sc.filter_tiles = (int)std::max<i64>(1, ceil_div(ix.d_pad, (i64)sc.filter_threads * FK_DOCS));
    sc.parts_stride = std::max(std::max(sc.n_tiles, sc.filter_tiles), 1);       // (the positional scorer has its own lists: ppart_*)
    sc.part_n.ensure((size_t)n * sc.parts_stride);
    sc.part_doc.ensure((size_t)n * sc.parts_stride * TOPN);
    sc.part_score.ensure((size_t)n * sc.parts_stride * TOPN);
}

// (start, length) directory of the positional index for the gap join; built once per handle, on first use
static bool ensure_gap_directory(bsp_classifier* h, cudaStream_t st) {
    GpuIndex& ix = h->ix;
    if (ix.pse.p) return true;
    if (ix.gap_join_off) return false;
    const u64 nt = 1ull << (2 * ix.k);
    const size_t S = ix.segs.size();
    {
        // 8 B x 4^k x segments: 0.64 GB at C2 (k = 10, 76 segments), 10 GB at k = 12 -- only where it fits comfortably
        size_t free_b = 0, total_b = 0;
        CUDA_CHECK(cudaMemGetInfo(&free_b, &total_b));
        if ((double)nt * (double)std::max<size_t>(S, 1) * 8.0 > 0.4 * (double)free_b) {
            ix.gap_join_off = true;
            ix.notes += "gap join off: its position directory does not fit; ";
            return false;
        }
    }
    ix.pse.alloc((size_t)nt * std::max<size_t>(S, 1));
    CUDA_CHECK(cudaMemsetAsync(ix.pse.p, 0, ix.pse.bytes(), st));
    std::vector<GjSeg> gs(S);
    size_t lut_total = 0;
    for (size_t i = 0; i < S; i++) lut_total += ((size_t)ix.segs[i].n_tok >> GJ_LUT_SHIFT) + 2;
    ix.gj_lut.alloc(lut_total);
    size_t lut_at = 0;
    for (size_t i = 0; i < S; i++) {
        const Segment& sg = ix.segs[i];
        const u32 nb = (u32)(((size_t)sg.n_tok >> GJ_LUT_SHIFT) + 2);
        gs[i].positions = sg.positions.p; gs[i].doc_tok = sg.doc_tok.p; gs[i].doc_lut = ix.gj_lut.p + lut_at;
        gs[i].n_docs = sg.n_docs; gs[i].doc_base = sg.doc_base;
        if (sg.n_terms) pse_build_kernel<<<(unsigned)ceil_div(sg.n_terms, 256), 256, 0, st>>>(sg.view(), ix.pse.p + (size_t)i * nt);
        gj_doc_lut_kernel<<<(unsigned)ceil_div(nb, 256), 256, 0, st>>>(sg.doc_tok.p, sg.n_docs, nb, ix.gj_lut.p + lut_at);
        lut_at += nb;
    }
    KERNEL_CHECK();
    ix.d_gjsegs.alloc(S + 1);
    if (S) CUDA_CHECK(cudaMemcpyAsync(ix.d_gjsegs.p, gs.data(), S * sizeof(GjSeg), cudaMemcpyHostToDevice, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    ix.positional_bytes += ix.pse.bytes();
    CUDA_CHECK(cudaFuncSetAttribute(gap_join_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GJ_SMEM_BYTES));
    CUDA_CHECK(cudaFuncSetAttribute(gap_join_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    return true;
}

// Room for the gap-clause match lists behind the exception arrays (`entries` per pipeline slot).  Called with every slot idle.
static void ensure_gap_tail(bsp_classifier* h, size_t entries) {
    GpuIndex& ix = h->ix;
    if (!ix.tables || entries <= ix.gap_tail) return;
    if ((u64)ix.n_exc + (u64)BSP_SLOTS * entries >= ((u64)1 << 32)) entries = (size_t)((((u64)1 << 32) - 1 - (u64)ix.n_exc) / BSP_SLOTS);
    if (entries <= ix.gap_tail) return;
    for (auto& sc : h->sc) CUDA_CHECK(cudaStreamSynchronize(sc.st));
    DevBuf<u32> nd; DevBuf<float> nf;
    nd.alloc((size_t)ix.n_exc + BSP_SLOTS * entries); nf.alloc((size_t)ix.n_exc + BSP_SLOTS * entries);
    if (ix.n_exc) {
        CUDA_CHECK(cudaMemcpy(nd.p, ix.exc_doc.p, (size_t)ix.n_exc * 4, cudaMemcpyDeviceToDevice));
        CUDA_CHECK(cudaMemcpy(nf.p, ix.exc_freq.p, (size_t)ix.n_exc * 4, cudaMemcpyDeviceToDevice));
    }
    ix.table_bytes += nd.bytes() + nf.bytes() - ix.exc_doc.bytes() - ix.exc_freq.bytes();
    ix.exc_doc = std::move(nd); ix.exc_freq = std::move(nf);
    ix.gap_tail = entries;
}
// before a batch: with kmer_skips > 0 every phrase clause is a gap clause -- about `clauses` of them per sub-batch
static void prepare_gap_tail(bsp_classifier* h, i64 clauses) {
    size_t want = h->gap_tail_want;
    if (h->conf.kmer_skips > 0 && h->conf.query_algorithm != BSP_NAIVE_KMER) want = std::max<size_t>(want, (size_t)clauses * 6);
    ensure_gap_tail(h, want);
}

// The sub-batch's gap clauses against the positional index (gapjoin.cuh): sort by first k-mer, chunks, the join, then the
// matches sorted by (clause, column) into the slot's tail of the exception arrays + per-clause offsets.
static void gap_join(bsp_classifier* h, Scratch& sc, u32 n_g) {
    GpuIndex& ix = h->ix;
    cudaStream_t st = sc.st;
    sc.gj_keys.ensure(n_g); sc.gj_keys_alt.ensure(n_g); sc.gj_vals.ensure(n_g); sc.gj_vals_alt.ensure(n_g); sc.gj_chunks.ensure(n_g);
    u32 *keys = sc.gj_keys.p, *keys_alt = sc.gj_keys_alt.p, *vals = sc.gj_vals.p, *vals_alt = sc.gj_vals_alt.p;
    cudaEvent_t e0, e1, e2;
    CUDA_CHECK(cudaEventCreate(&e0)); CUDA_CHECK(cudaEventCreate(&e1)); CUDA_CHECK(cudaEventCreate(&e2));
    gj_keys_kernel<<<(unsigned)ceil_div(n_g, 256), 256, 0, st>>>(sc.gaps.p, n_g, keys, vals);
    radix_sort_pairs<u32>(&keys, &vals, &keys_alt, &vals_alt, n_g, 2 * ix.k, sc.gj_rst, st);
    sc.gj_sizes.ensure(32);
    CUDA_CHECK(cudaMemsetAsync(sc.gj_sizes.p, 0, 32 * 4, st));
    gj_chunk_kernel<<<(unsigned)ceil_div(n_g, 256), 256, 0, st>>>(keys, n_g, sc.gj_chunks.p, sc.gj_sizes.p, 0);
    gj_chunk_offsets_kernel<<<1, 1, 0, st>>>(sc.gj_sizes.p, sc.gap_ctr.p + 2);
    gj_chunk_kernel<<<(unsigned)ceil_div(n_g, 256), 256, 0, st>>>(keys, n_g, sc.gj_chunks.p, sc.gj_sizes.p, 1);
    KERNEL_CHECK();
    CUDA_CHECK(cudaMemcpyAsync(sc.h_counts + 7, sc.gap_ctr.p + 2, 4, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    const u32 n_chunks = (u32)sc.h_counts[7];
    const u32 cap = (u32)std::min<size_t>(ix.gap_tail, 0xFFFFFFF0u);
    sc.gj_mg.ensure(cap); sc.gj_mcol.ensure(cap); sc.gj_mfreq.ensure(cap); sc.gj_cand.ensure(cap);
    CUDA_CHECK(cudaMemsetAsync(sc.gj_bytes.p, 0, 8, st));
    GjOut out;
    out.g = sc.gj_mg.p; out.col = sc.gj_mcol.p; out.freq = sc.gj_mfreq.p; out.cap = cap;
    out.n_match = sc.gap_ctr.p + 1; out.pos_bytes = sc.gj_bytes.p;
    CUDA_CHECK(cudaEventRecord(e0, st));
    if (n_chunks) {
        // the join appends candidates (clause, segment, occurrence); a thread per candidate then does the per-document work
        gap_join_kernel<<<(unsigned)ceil_div(n_chunks, GJ_WARPS), GJ_WARPS * 32, GJ_SMEM_BYTES, st>>>(
            ix.d_gjsegs.p, (int)ix.segs.size(), ix.pse.p, 1ull << (2 * ix.k), sc.gaps.p, vals, sc.gj_chunks.p, sc.gap_ctr.p + 2,
            sc.gj_cand.p, sc.gap_ctr.p + 3, cap, sc.gj_bytes.p, sc.route.p, sc.status_in.p, sc.qp.positional_ok);
        CUDA_CHECK(cudaEventRecord(e1, st));
        gj_verify_kernel<<<(unsigned)ceil_div(cap, 256), 256, 0, st>>>(sc.gj_cand.p, sc.gap_ctr.p + 3, cap, ix.d_gjsegs.p, ix.pse.p,
                                                                       1ull << (2 * ix.k), sc.gaps.p, ix.colmap.p, out);
    } else {
        CUDA_CHECK(cudaEventRecord(e1, st));
    }
    KERNEL_CHECK();
    CUDA_CHECK(cudaEventRecord(e2, st));
    unsigned long long pos_bytes = 0;
    CUDA_CHECK(cudaMemcpyAsync(sc.h_counts + 8, sc.gap_ctr.p + 1, 4, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaMemcpyAsync(sc.h_counts + 9, sc.gap_ctr.p + 3, 4, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaMemcpyAsync(&pos_bytes, sc.gj_bytes.p, 8, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    float ms = 0, ms_v = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventElapsedTime(&ms_v, e1, e2);
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
    u32 n_m = (u32)sc.h_counts[8];
    const u32 n_cand = (u32)sc.h_counts[9];
    const int overflow = (n_m > cap || n_cand > cap) ? 1 : 0;
    if (overflow) { h->gap_tail_want = std::max<size_t>(h->gap_tail_want, (size_t)std::max(n_m, n_cand) * 2); n_m = 0; }   // more room next batch
    h->gj_cands += n_cand; h->gj_verify_ms += ms_v;
    h->gj_clauses += n_g; h->gj_chunks += n_chunks; h->gj_matches += n_m; h->gj_bytes += (i64)pos_bytes; h->gj_ms += ms;
    const u32 tmp_base = (u32)ix.n_exc + (u32)((size_t)sc.slot_index * ix.gap_tail);
    sc.gj_k64.ensure(std::max<u32>(n_m, 1)); sc.gj_k64_alt.ensure(std::max<u32>(n_m, 1));
    sc.gj_v.ensure(std::max<u32>(n_m, 1)); sc.gj_v_alt.ensure(std::max<u32>(n_m, 1));
    u64 *k64 = sc.gj_k64.p, *k64_alt = sc.gj_k64_alt.p; u32 *v = sc.gj_v.p, *v_alt = sc.gj_v_alt.p;
    if (n_m) {
        exc_keys_kernel<<<(unsigned)ceil_div(n_m, 256), 256, 0, st>>>(sc.gj_mg.p, sc.gj_mcol.p, n_m, k64, v);
        int g_bits = 1;
        while ((1ull << g_bits) < (u64)n_g) g_bits++;
        radix_sort_pairs<u64>(&k64, &v, &k64_alt, &v_alt, n_m, 32 + g_bits, sc.gj_rst, st);
        exc_gather_kernel<<<(unsigned)ceil_div(n_m, 256), 256, 0, st>>>(k64, v, sc.gj_mfreq.p, n_m, ix.exc_doc.p + tmp_base, ix.exc_freq.p + tmp_base);
    }
    exc_offsets_kernel<<<(unsigned)ceil_div((u64)n_g + 1, 256), 256, 0, st>>>(k64, n_m, n_g, sc.gap_off.p);
    gj_overflow_kernel<<<(unsigned)ceil_div(n_g, 256), 256, 0, st>>>(sc.gaps.p, n_g, sc.gap_off.p, overflow, sc.route.p, sc.status_in.p,
                                                                     sc.qp.positional_ok);
    CUDA_CHECK(cudaMemsetAsync(sc.counts.p, 0, 32, st));
    route_compact_kernel<<<(unsigned)ceil_div(sc.n, 256), 256, 0, st>>>(sc.route.p, sc.n_clauses.p, sc.n, sc.list_dense.p, sc.list_pos.p,
                                                                        sc.list_filter.p, sc.counts.p, sc.pos_slot.p);
    KERNEL_CHECK();
    CUDA_CHECK(cudaMemcpyAsync(sc.h_counts, sc.counts.p, 16, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaMemcpyAsync(sc.h_counts + 6, sc.counts.p + 6, 4, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    sc.launches_gap = 8 + (n_m ? 2 + 2 * ((32 + 24 + 7) / 8) : 0) + 2 * ((2 * ix.k + 7) / 8);      // (sort passes: histogram + scatter each)
}

// phase 2a: wait for the route counts, (gap pre-pass), launch the filter scorer and queue the read-back of what it hands
// over -- no wait for it.  phase 2b: wait, then the exact / positional scorers, merge, on the same stream.
static void pipeline_phase2a(bsp_classifier* h, Scratch& sc) {
    GpuIndex& ix = h->ix;
    cudaStream_t st = sc.st;
    const i32 n = sc.n;
    const QueryParams& qp = sc.qp;
    CUDA_CHECK(cudaStreamSynchronize(st));                  // route counts have arrived
    sc.launches_gap = 0;
    if (qp.gap_ok && (i64)(u32)sc.h_counts[5] > sc.gc) {
        // more gap clauses than the list holds: grow it and redo phase 1 (the clauses past the end were not kept)
        sc.gc_min = (i64)(u32)sc.h_counts[5] + (i64)(u32)sc.h_counts[5] / 8 + 1024;
        pipeline_phase1(h, sc, n, sc.in_seq, sc.in_off, sc.in_bytes);
        CUDA_CHECK(cudaStreamSynchronize(st));
    }
    if (qp.gap_ok && sc.h_counts[5] > 0) gap_join(h, sc, (u32)std::min<i64>((u32)sc.h_counts[5], sc.gc));
    i32* counts = sc.cnt;
    counts[0] = sc.h_counts[0]; counts[1] = sc.h_counts[1]; counts[2] = sc.h_counts[2]; counts[3] = 0;
    const u32 D = (u32)ix.docs.size();
    const int parts_stride = sc.parts_stride;
    i32* part_n = sc.part_n.p; u32* part_doc = sc.part_doc.p; float* part_score = sc.part_score.p;
    const u32 gbase = (u32)h->doc_base_global;
    const Tab4Dev T = ix.tab4_view(sc.slot_index, sc.gap_off.p);
    const int full_topn = (h->conf.full_topn || env_i64("BSP_FULL_TOPN", 0)) ? 1 : 0;
    CUDA_CHECK(cudaEventRecord(sc.ev[1], st));
    sc.launches = (sc.front_fused ? 7 : 9) + sc.launches_gap;      // tok_capacity, 4 scan kernels, front (or tokenize, token_df, weights), route_compact
    sc.filter_launched = false;
    if (counts[2] > 0) {
        CUDA_CHECK(cudaMemsetAsync(sc.fstats.p, 0, sizeof(FilterStats), st));
        const int threads = sc.filter_threads;
        const size_t ring_bytes = (size_t)FK_STAGES * threads * 16;
        if ((i64)counts[2] * sc.filter_tiles > 0x7FFFFFFF) throw BspError(BSP_ERR_UNSUPPORTED, "batch too large for the filter path");
        const bool tiled = sc.filter_tiles > 1, naive = qp.alg == ALG_NAIVE;
        const unsigned fgrid = (unsigned)((i64)counts[2] * sc.filter_tiles);
        const i32* fk_list = counts[2] == n ? nullptr : sc.list_filter.p;      // all reads: the list is a permutation of 0..n-1
        const bool bm25 = qp.sim == SIM_BM25;
        // third-generation filter (filter3.cuh): prep -> link -> persistent bulk-copy stream -> rescore.  tf-idf, one doc tile.
        sc.filter_v3 = !bm25 && sc.filter_tiles == 1 && ix.d_pad <= (u64)FK3_MAXW * 1024 && env_i64("BSP_FILTER_V3", 0) != 0;
        if (sc.filter_v3) {
            const i32 nf = counts[2];
            const int cand_cap = full_topn ? FK_CAND_CAP : 32;
            const int W = (int)std::max<i64>(1, ceil_div((i64)ix.d_pad, 1024));
            const size_t smem = f3_smem_bytes(W, naive);
            int per_sm = 1;
            if (naive) CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, filter3_stream_kernel<true>, W * 32, smem));
            else CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, filter3_stream_kernel<false>, W * 32, smem));
            per_sm = (int)std::max<i64>(1, std::min<i64>(per_sm, env_i64("BSP_F3_CTAS", 32)));
            const i32 G = (i32)std::max<i64>(1, std::min<i64>(nf, (i64)h->sm_count * per_sm));
            const size_t n_rec = (size_t)sc.tok_key.n + 2 * (size_t)n + 64;
            sc.f3_hdr.ensure(n); sc.f3_rec.ensure(n_rec);
            if (naive) sc.f3_rec_hi.ensure(n_rec);
            sc.f3_cells.ensure((size_t)n * FK3_CELLS); sc.f3_longs.ensure((size_t)n * FK3_LONG);
            sc.f3_cand.ensure((size_t)n * cand_cap); sc.f3_cand_n.ensure(n); sc.f3_tau.ensure(n); sc.f3_long_over.ensure(n);
            sc.f3_prolog.ensure((size_t)G * FK3_STAGES);
            const u32 zero_off = T.zero_row * (u32)(T.row_bytes >> 4);
            const unsigned pgrid = (unsigned)ceil_div((i64)nf * 32, 256);
#define F3_PREP(AT) filter3_prep_kernel<AT><<<pgrid, 256, 0, st>>>(T, h->tf16.p, fk_list, nf, sc.tok_off.p, sc.n_tok.p, sc.clause_val.p, \
                sc.clause_row.p, sc.n_clauses.p, sc.msm.p, sc.vrange.p, sc.read_ngap.p, sc.gaps.p, qp.alg, sc.f3_hdr.p, sc.f3_rec.p, sc.f3_rec_hi.p, \
                sc.f3_cells.p, sc.f3_longs.p, sc.f3_cand_n.p, sc.f3_tau.p, sc.f3_long_over.p, sc.route.p, sc.list_dense.p + counts[0], \
                sc.counts.p + 3, sc.fstats.p)
            if (naive) F3_PREP(true); else F3_PREP(false);
#undef F3_PREP
            filter3_link_kernel<<<pgrid, 256, 0, st>>>(sc.f3_hdr.p, sc.f3_rec.p, fk_list, nf, sc.tok_off.p, G, zero_off, sc.f3_prolog.p);
#define F3_STREAM(AT) filter3_stream_kernel<AT><<<(unsigned)G, W * 32, smem, st>>>(T, sc.f3_hdr.p, sc.f3_rec.p, sc.f3_rec_hi.p, sc.f3_cells.p, \
                sc.f3_longs.p, sc.f3_prolog.p, fk_list, nf, sc.tok_off.p, h->tf16.p, h->col_w.p, h->group_w.p, cand_cap, sc.f3_cand_n.p, \
                sc.f3_cand.p, sc.f3_tau.p, sc.f3_long_over.p, full_topn, 1u)
            if (naive) F3_STREAM(true); else F3_STREAM(false);
#undef F3_STREAM
            filter3_rescore_kernel<<<pgrid, 256, 0, st>>>(T, sc.f3_hdr.p, sc.f3_cells.p, sc.f3_longs.p, fk_list, nf, sc.tok_off.p,
                sc.clause_val.p, sc.clause_row.p, h->tf16.p, h->col_w.p, cand_cap, sc.f3_cand_n.p, sc.f3_cand.p, sc.f3_tau.p,
                sc.f3_long_over.p, gbase, parts_stride, part_n, part_doc, part_score, sc.route.p, sc.list_dense.p + counts[0],
                sc.counts.p + 3, sc.fstats.p);
            sc.launches += 3;
        } else {
        Bm25Classes bm;
        bm.class_k = h->class_k.p; bm.group_class = h->group_class.p; bm.n_classes = h->n_classes;
        bm.nc_cap = (int)round_up((i64)std::min<i64>(std::max(sc.h_counts[6], 1), qp.filter_nc_max) + 1, 2);
        bm.k_min = h->k_min; bm.k_max = h->k_max; bm.k1p1 = 1.2f + 1.0f;
        const size_t lut_bytes = bm25 ? (size_t)bm.n_classes * bm.nc_cap * (naive ? 24 : 16) : 0;
#define FK_LAUNCH(TI, AT, SIM) filter_score_kernel<TI, AT, SIM><<<fgrid, threads, ring_bytes + lut_bytes, st>>>(T, D, h->tf16.p, bm, fk_list, sc.tok_off.p, sc.n_tok.p, \
            sc.clause_val.p, sc.clause_row.p, sc.n_clauses.p, sc.msm.p, sc.vrange.p, sc.read_ngap.p, h->col_w.p, qp.alg, gbase, parts_stride, part_n, \
            part_doc, part_score, sc.route.p, sc.list_dense.p + counts[0], sc.counts.p + 3, sc.fstats.p, full_topn, 1u, sc.filter_tiles)
#define FK_LAUNCH2(TI, AT) do { if (bm25) FK_LAUNCH(TI, AT, SIM_BM25); else FK_LAUNCH(TI, AT, SIM_TFIDF); } while (0)
        if (tiled) { if (naive) FK_LAUNCH2(true, true); else FK_LAUNCH2(true, false); }
        else { if (naive) FK_LAUNCH2(false, true); else FK_LAUNCH2(false, false); }
#undef FK_LAUNCH2
#undef FK_LAUNCH
        }
        KERNEL_CHECK();
        sc.launches++;
        sc.filter_launched = true;
        // reads the filter could not bound (exception or candidate overflow) were appended to the exact list
        CUDA_CHECK(cudaMemcpyAsync(sc.h_counts + 4, sc.counts.p + 3, 4, cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaMemcpyAsync(sc.h_fstats, sc.fstats.p, sizeof(FilterStats), cudaMemcpyDeviceToHost, st));
    }
}

static void pipeline_phase2b(bsp_classifier* h, Scratch& sc, const BatchOut& out) {
    GpuIndex& ix = h->ix;
    cudaStream_t st = sc.st;
    const i32 n = sc.n;
    const QueryParams& qp = sc.qp;
    i32* counts = sc.cnt;
    const u32 D = (u32)ix.docs.size();
    const int n_tiles = sc.n_tiles, parts_pos = sc.parts_pos, parts_stride = sc.parts_stride;
    i32* part_n = sc.part_n.p; u32* part_doc = sc.part_doc.p; float* part_score = sc.part_score.p;
    SimParams sp; sp.sim = qp.sim; sp.k1p1 = 1.2f + 1.0f;
    const u32 gbase = (u32)h->doc_base_global;
    const Tab4Dev T = ix.tab4_view(sc.slot_index, sc.gap_off.p);
    const int full_topn = (h->conf.full_topn || env_i64("BSP_FULL_TOPN", 0)) ? 1 : 0;
    i64 launches = sc.launches;
    FilterStats fs = {0, 0, 0};
    if (sc.filter_launched) {
        CUDA_CHECK(cudaStreamSynchronize(st));
        counts[3] = sc.h_counts[4];
        fs = *sc.h_fstats;
    }
    const i32 n_exact = counts[0] + counts[3];
    if (n_exact > 0) {
        unsigned grid = (unsigned)((i64)n_exact * n_tiles);
        if (qp.sim == SIM_TFIDF)
            exact_score_kernel<SIM_TFIDF><<<grid, sc.dense_threads, 0, st>>>(T, D, h->tf16.p, sc.list_dense.p, sc.tok_off.p, sc.n_tok.p,
                sc.clause_val.p, sc.clause_row.p, sc.n_clauses.p, sc.msm.p, h->col_w.p, sp.k1p1, gbase, n_tiles, parts_stride,
                part_n, part_doc, part_score);
        else
            exact_score_kernel<SIM_BM25><<<grid, sc.dense_threads, 0, st>>>(T, D, h->tf16.p, sc.list_dense.p, sc.tok_off.p, sc.n_tok.p,
                sc.clause_val.p, sc.clause_row.p, sc.n_clauses.p, sc.msm.p, h->col_w.p, sp.k1p1, gbase, n_tiles, parts_stride,
                part_n, part_doc, part_score);
        launches++;
    }
    if (counts[1] > 0) {
        size_t smem = (size_t)ix.max_seg_docs * 10 + 16;
        i64 blocks = (i64)counts[1] * parts_pos;
        // partial top-10 lists of the positional scorer: one per (list entry, segment)
        sc.ppart_n.ensure((size_t)blocks); sc.ppart_doc.ensure((size_t)blocks * TOPN); sc.ppart_score.ensure((size_t)blocks * TOPN);
        const i64 MAXG = 0x7FFFFFFF;
        if (blocks > MAXG) throw BspError(BSP_ERR_UNSUPPORTED, "batch too large for the positional path");
        if (ix.max_seg_docs <= POSW_DOCS && env_i64("BSP_POS_BLOCK_KERNEL", 0) == 0) {
            // few documents per segment (large genomes): one warp per (segment, read)
            const i64 wblocks = ceil_div(blocks, POSW_WARPS);
            const int dcap = (int)round_up(std::max<i64>(ix.max_seg_docs, 1), 32);
            score_positional_warp_kernel<<<(unsigned)wblocks, POSW_WARPS * 32, POSW_WARPS * posw_warp_bytes(dcap, PW_POS), st>>>(ix.d_segs.p, parts_pos, sc.list_pos.p, counts[1],
                sc.tok_off.p, sc.n_tok.p, sc.tok_key.p, sc.tok_pos.p, sc.clause_val.p, sc.n_clauses.p, sc.msm.p, h->doc_w.p, qp, sp, gbase,
                sc.ppart_n.p, sc.ppart_doc.p, sc.ppart_score.p, (int)env_i64("BSP_POS_STAGE", 1), dcap, PW_POS);
        } else {
            score_positional_kernel<<<(unsigned)blocks, POS_THREADS, smem, st>>>(ix.d_segs.p, parts_pos, sc.list_pos.p, counts[1], sc.tok_off.p,
                sc.n_tok.p, sc.tok_key.p, sc.tok_pos.p, sc.clause_val.p, sc.n_clauses.p, sc.msm.p, h->doc_w.p, qp, sp, gbase,
                sc.ppart_n.p, sc.ppart_doc.p, sc.ppart_score.p);
        }
        launches++;
    }
    KERNEL_CHECK();
    CUDA_CHECK(cudaEventRecord(sc.ev[2], st));
    merge_kernel<<<(unsigned)ceil_div(n, 128), 128, 0, st>>>(n, sc.route.p, sc.status_in.p, n_tiles, parts_pos, sc.filter_v3 ? 1 : sc.filter_tiles, parts_stride, part_n,
                                                             part_doc, part_score, sc.pos_slot.p, sc.ppart_n.p, sc.ppart_doc.p, sc.ppart_score.p,
                                                             out.status, out.label, out.n_hits, out.n_top, out.top_doc, out.top_score, full_topn);
    KERNEL_CHECK();
    CUDA_CHECK(cudaEventRecord(sc.ev[3], st));
    h->counters[0] += n; h->counters[1] += n_exact; h->counters[2] += counts[1]; h->counters[3] += counts[2] - counts[3];
    h->counters[4] += launches + 1; h->counters[5] += (i64)fs.candidates; h->counters[6] += (i64)fs.forced;
    h->counters[7] += counts[3];
}

// after the slot's stream has been synchronised: add its CUDA-event timings to the handle
static void pipeline_collect(bsp_classifier* h, Scratch& sc) {
    if (!sc.busy) return;
    float ms_score = 0, ms_pipe = 0;
    cudaEventElapsedTime(&ms_score, sc.ev[1], sc.ev[2]);
    cudaEventElapsedTime(&ms_pipe, sc.ev[0], sc.ev[3]);
    h->score_ms += ms_score; h->pipe_ms += ms_pipe;
    if (sc.routes_pending && (size_t)sc.r0 + (size_t)sc.n <= h->last_routes.size())
        memcpy(h->last_routes.data() + sc.r0, sc.pin_routes.p, (size_t)sc.n * 4);
    sc.routes_pending = false;
    sc.busy = false;
}

// Runs the device pipeline on reads already resident in device memory (one sub-batch, slot 0).
static void classify_on_device(bsp_classifier* h, i32 n, const u8* d_seq, const i64* d_seq_off, i64 total_bytes,
                               const BatchOut& out) {
    Scratch& sc = h->sc[0];
    prepare_gap_tail(h, total_bytes / (h->conf.kmer_skips + 1) + 2 * (i64)n);
    h->gj_clauses = h->gj_chunks = h->gj_matches = h->gj_bytes = h->gj_cands = 0; h->gj_ms = h->gj_verify_ms = 0;
    pipeline_phase1(h, sc, n, d_seq, d_seq_off, total_bytes);
    pipeline_phase2a(h, sc);
    pipeline_phase2b(h, sc, out);
    CUDA_CHECK(cudaStreamSynchronize(sc.st));
    pipeline_collect(h, sc);
}

// D2H of a finished sub-batch into the caller's arrays (async on the slot's stream; pinned destinations overlap
// with the other slot's kernels, pageable ones are staged by the driver)
static void enqueue_outputs(bsp_classifier* h, Scratch& sc, i32* status, i32* label, i32* n_hits, i32* n_top, i32* top_doc,
                            float* top_score) {
    const i32 r0 = sc.r0, m = sc.n;
    cudaStream_t st = sc.st;
    CUDA_CHECK(cudaMemcpyAsync(status + r0, sc.o_status.p, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
    if (label) CUDA_CHECK(cudaMemcpyAsync(label + r0, sc.o_label.p, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
    if (n_hits) CUDA_CHECK(cudaMemcpyAsync(n_hits + r0, sc.o_nhits.p, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
    if (n_top) CUDA_CHECK(cudaMemcpyAsync(n_top + r0, sc.o_ntop.p, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
    if (top_doc) CUDA_CHECK(cudaMemcpyAsync(top_doc + (i64)r0 * TOPN, sc.o_doc.p, (size_t)m * TOPN * 4, cudaMemcpyDeviceToHost, st));
    if (top_score) CUDA_CHECK(cudaMemcpyAsync(top_score + (i64)r0 * TOPN, sc.o_score.p, (size_t)m * TOPN * 4, cudaMemcpyDeviceToHost, st));
    sc.pin_routes.ensure((size_t)m * 4);
    CUDA_CHECK(cudaMemcpyAsync(sc.pin_routes.p, sc.route.p, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
    sc.routes_pending = true;
}

// taxonomy labels of a finished sub-batch (its merged hit lists are on the device): one more small kernel + two copies
static void enqueue_taxonomy(bsp_classifier* h, Scratch& sc, i32* tax_label, i32* tax_index) {
    const i32 m = sc.n;
    sc.o_taxlabel.ensure(m); sc.o_taxidx.ensure(m);
    taxonomy_label_kernel<<<(unsigned)ceil_div(m, 128), 128, 0, sc.st>>>(m, sc.o_nhits.p, sc.o_doc.p, h->lin_off.p, h->lin_tax.p, h->lin_flags.p,
                                                                         h->lin_docs, sc.o_taxlabel.p, sc.o_taxidx.p);
    KERNEL_CHECK();
    CUDA_CHECK(cudaMemcpyAsync(tax_label + sc.r0, sc.o_taxlabel.p, (size_t)m * 4, cudaMemcpyDeviceToHost, sc.st));
    if (tax_index) CUDA_CHECK(cudaMemcpyAsync(tax_index + sc.r0, sc.o_taxidx.p, (size_t)m * 4, cudaMemcpyDeviceToHost, sc.st));
}

static int classify_packed_impl(bsp_classifier* h, i32 n, const char* seq_bytes, const i64* seq_offsets,
                                i32* status, i32* label, i32* n_hits, i32* n_top, i32* top_doc, float* top_score,
                                i32* tax_label = nullptr, i32* tax_index = nullptr) {
    if (!h) return fail(nullptr, BSP_ERR_ARG, "handle is null");
    std::lock_guard<std::mutex> lk(h->mu);
    API_TRY(h)
    if (n < 0 || (n > 0 && (!seq_bytes || !seq_offsets || !status))) return fail(&h->last_error, BSP_ERR_ARG, "null argument");
    if (tax_label && h->lin_docs == 0) return fail(&h->last_error, BSP_ERR_STATE, "taxonomy labels requested before bsp_classifier_set_lineages");
    CUDA_CHECK(cudaSetDevice(h->conf.device));
    for (auto& c : h->counters) c = 0;
    h->score_ms = 0; h->pipe_ms = 0;
    h->last_routes.assign(n, 0);
    // (kmer_skips > 0: the gap join shares a first k-mer's positions among the clauses of the sub-batch that start with it --
    //  4^k terms want a sub-batch of a million reads)
    const bool all_gap = h->conf.kmer_skips > 0 && h->conf.query_algorithm != BSP_NAIVE_KMER;
    const i32 SUB = (i32)std::max<i64>(1, all_gap ? env_i64("BSP_SUB_BATCH_GAP", 1048576) : env_i64("BSP_SUB_BATCH", 131072));
    {
        const i32 m0 = std::min(SUB, n);
        const i64 bytes0 = n > 0 ? (seq_offsets[n] - seq_offsets[0]) / std::max(1, n) * m0 : 0;
        prepare_gap_tail(h, bytes0 / (h->conf.kmer_skips + 1) + 2 * (i64)m0);
    }
    h->gj_clauses = h->gj_chunks = h->gj_matches = h->gj_bytes = h->gj_cands = 0; h->gj_ms = h->gj_verify_ms = 0;
    // Software pipeline over sub-batches, BSP_SLOTS slots.  Iteration i copies sub-batch i in and queues its tokenizer /
    // weights (phase 1), launches the scorer of sub-batch i-1 (phase 2a: its route counts arrived while sub-batch i-2
    // was being scored) and only then waits for the scorer of sub-batch i-2 to finish, merges and copies it out (phase
    // 2b): the GPU always has the next scoring kernel queued behind the running one.
    auto finish = [&](Scratch& sc) {
        BatchOut bo{sc.o_status.p, sc.o_label.p, sc.o_nhits.p, sc.o_ntop.p, sc.o_doc.p, sc.o_score.p};
        pipeline_phase2b(h, sc, bo);
        enqueue_outputs(h, sc, status, label, n_hits, n_top, top_doc, top_score);
        if (tax_label) enqueue_taxonomy(h, sc, tax_label, tax_index);
    };
    try {
        i32 i = 0;                                        // sub-batch index
        for (i32 r0 = 0; r0 < n; r0 += SUB, i++) {
            Scratch& sc = h->sc[i % BSP_SLOTS];
            if (sc.busy) {          // the slot's previous sub-batch (BSP_SLOTS iterations ago) must be completely done
                CUDA_CHECK(cudaStreamSynchronize(sc.st));
                pipeline_collect(h, sc);
            }
            const i32 m = std::min(SUB, n - r0);
            const i64 b0 = seq_offsets[r0], bytes = seq_offsets[r0 + m] - b0;
            sc.seq.ensure((size_t)bytes + 64);
            sc.seq_off.ensure((size_t)m + 1);
            const size_t in_need = (size_t)(m + 1) * 8;           // offsets rebased to the sub-batch, staged in pinned memory
            if (in_need > sc.pin_in_bytes) {
                if (sc.pin_in) cudaFreeHost(sc.pin_in);
                sc.pin_in = nullptr; sc.pin_in_bytes = 0;
                CUDA_CHECK(cudaMallocHost((void**)&sc.pin_in, in_need + in_need / 4));
                sc.pin_in_bytes = in_need + in_need / 4;
            }
            i64* po = (i64*)sc.pin_in;
            for (i32 j = 0; j <= m; j++) po[j] = seq_offsets[r0 + j] - b0;
            CUDA_CHECK(cudaMemcpyAsync(sc.seq_off.p, po, (size_t)(m + 1) * 8, cudaMemcpyHostToDevice, sc.st));
            if (bytes) CUDA_CHECK(cudaMemcpyAsync(sc.seq.p, seq_bytes + b0, (size_t)bytes, cudaMemcpyHostToDevice, sc.st));
            sc.o_status.ensure(m); sc.o_label.ensure(m); sc.o_nhits.ensure(m); sc.o_ntop.ensure(m);
            sc.o_doc.ensure((size_t)m * TOPN); sc.o_score.ensure((size_t)m * TOPN);
            sc.r0 = r0;
            pipeline_phase1(h, sc, m, sc.seq.p, sc.seq_off.p, bytes);
            if (i >= 1) pipeline_phase2a(h, h->sc[(i - 1) % BSP_SLOTS]);
            if (i >= 2) finish(h->sc[(i - 2) % BSP_SLOTS]);
        }
        if (i >= 1) pipeline_phase2a(h, h->sc[(i - 1) % BSP_SLOTS]);
        if (i >= 2) finish(h->sc[(i - 2) % BSP_SLOTS]);
        if (i >= 1) finish(h->sc[(i - 1) % BSP_SLOTS]);
        for (auto& sc : h->sc) {
            CUDA_CHECK(cudaStreamSynchronize(sc.st));
            pipeline_collect(h, sc);
        }
    } catch (...) {
        for (auto& sc : h->sc) { cudaStreamSynchronize(sc.st); sc.busy = false; }
        throw;
    }
    return BSP_OK;
    API_CATCH(h)
}

BSP_API int bsp_classify_packed(bsp_classifier* h, int32_t n, const char* seq_bytes, const int64_t* seq_offsets,
                                int32_t* status, int32_t* label, int32_t* n_hits, int32_t* n_top,
                                int32_t* top_doc, float* top_score) {
    return classify_packed_impl(h, n, seq_bytes, seq_offsets, status, label, n_hits, n_top, top_doc, top_score);
}

BSP_API int bsp_classify_packed_tax(bsp_classifier* h, int32_t n, const char* seq_bytes, const int64_t* seq_offsets,
                                    int32_t* status, int32_t* label, int32_t* n_hits, int32_t* n_top,
                                    int32_t* top_doc, float* top_score, int32_t* tax_label, int32_t* tax_index) {
    if (h && !tax_label) return fail(&h->last_error, BSP_ERR_ARG, "tax_label is null");
    return classify_packed_impl(h, n, seq_bytes, seq_offsets, status, label, n_hits, n_top, top_doc, top_score, tax_label, tax_index);
}

BSP_API int bsp_classifier_set_lineages(bsp_classifier* h, int32_t n_docs, const int64_t* lineage_off, const int32_t* taxid,
                                        const uint8_t* flags) {
    if (!h) return fail(nullptr, BSP_ERR_ARG, "handle is null");
    std::lock_guard<std::mutex> lk(h->mu);
    API_TRY(h)
    const i64 expect = h->global_docs > 0 ? h->global_docs : (i64)h->ix.docs.size();
    if (n_docs <= 0 || !lineage_off || !flags || (lineage_off[n_docs] > 0 && !taxid)) return fail(&h->last_error, BSP_ERR_ARG, "null argument");
    if ((i64)n_docs != expect) return fail(&h->last_error, BSP_ERR_ARG, "lineages must cover every document of the index (global ids)");
    if (lineage_off[0] != 0) return fail(&h->last_error, BSP_ERR_ARG, "lineage_off[0] must be 0");
    for (i32 d = 0; d < n_docs; d++) if (lineage_off[d + 1] < lineage_off[d]) return fail(&h->last_error, BSP_ERR_ARG, "lineage_off must be non-decreasing");
    CUDA_CHECK(cudaSetDevice(h->conf.device));
    const i64 total = lineage_off[n_docs];
    h->lin_off.alloc((size_t)n_docs + 1);
    h->lin_tax.alloc((size_t)std::max<i64>(total, 1));
    h->lin_flags.alloc((size_t)n_docs);
    CUDA_CHECK(cudaMemcpyAsync(h->lin_off.p, lineage_off, ((size_t)n_docs + 1) * 8, cudaMemcpyHostToDevice, h->st));
    if (total) CUDA_CHECK(cudaMemcpyAsync(h->lin_tax.p, taxid, (size_t)total * 4, cudaMemcpyHostToDevice, h->st));
    CUDA_CHECK(cudaMemcpyAsync(h->lin_flags.p, flags, (size_t)n_docs, cudaMemcpyHostToDevice, h->st));
    CUDA_CHECK(cudaStreamSynchronize(h->st));
    h->lin_docs = n_docs;
    return BSP_OK;
    API_CATCH(h)
}

BSP_API int bsp_classify_batch(bsp_classifier* h, int32_t n, const char* const* seqs, const int32_t* seq_lens,
                               int32_t* status, int32_t* label, int32_t* n_hits, int32_t* n_top,
                               int32_t* top_doc, float* top_score) {
    if (!h) return fail(nullptr, BSP_ERR_ARG, "handle is null");
    if (n < 0 || (n > 0 && (!seqs || !seq_lens))) return fail(&h->last_error, BSP_ERR_ARG, "null argument");
    std::vector<i64> off((size_t)n + 1, 0);
    for (i32 i = 0; i < n; i++) {
        if (seq_lens[i] < 0 || (seq_lens[i] > 0 && !seqs[i])) return fail(&h->last_error, BSP_ERR_ARG, "sequence is null or empty");
        off[i + 1] = off[i] + seq_lens[i];
    }
    std::string bytes((size_t)off[n], '\0');
    for (i32 i = 0; i < n; i++) if (seq_lens[i]) memcpy(&bytes[off[i]], seqs[i], (size_t)seq_lens[i]);
    return classify_packed_impl(h, n, bytes.data(), off.data(), status, label, n_hits, n_top, top_doc, top_score);
}

BSP_API int bsp_classifier_set_microbatch(bsp_classifier* h, int32_t max_reads, int64_t max_wait_us) {
    if (!h) return fail(nullptr, BSP_ERR_ARG, "handle is null");
    if (max_reads < 1 || max_wait_us < 0) return fail(&h->last_error, BSP_ERR_ARG, "micro-batch: max_reads >= 1 and max_wait_us >= 0");
    std::lock_guard<std::mutex> lk(h->mb.mu);
    h->mb.max_reads = max_reads;
    h->mb.wait_us = max_wait_us;
    return BSP_OK;
}

BSP_API int bsp_microbatch_stats(bsp_classifier* h, int64_t* batches, int64_t* reads) {
    if (!h) return fail(nullptr, BSP_ERR_ARG, "handle is null");
    std::lock_guard<std::mutex> lk(h->mb.mu);
    if (batches) *batches = h->mb.batches;
    if (reads) *reads = h->mb.reads;
    return BSP_OK;
}

// the leader's part: pack the batch into page-locked staging, one bsp_classify_packed, scatter the results
static void microbatch_run(bsp_classifier* h, std::vector<OneRead*>& batch) {
    MicroBatcher& mb = h->mb;
    std::lock_guard<std::mutex> run(mb.run_mu);
    const i32 n = (i32)batch.size();
    int rc = BSP_OK;
    std::string err;
    try {
        CUDA_CHECK(cudaSetDevice(h->conf.device));
        i64 total = 0;
        for (i32 i = 0; i < n; i++) total += batch[i]->len;
        mb.pin_off.ensure(((size_t)n + 1) * 8);
        mb.pin_seq.ensure((size_t)total + 1);
        mb.pin_i32.ensure((size_t)n * (4 + TOPN) * 4);
        mb.pin_f32.ensure((size_t)n * TOPN * 4);
        i64* off = (i64*)mb.pin_off.p;
        char* bytes = (char*)mb.pin_seq.p;
        off[0] = 0;
        for (i32 i = 0; i < n; i++) {
            memcpy(bytes + off[i], batch[i]->seq, (size_t)batch[i]->len);
            off[i + 1] = off[i] + batch[i]->len;
        }
        i32* st = (i32*)mb.pin_i32.p; i32* label = st + n; i32* nh = label + n; i32* nt = nh + n; i32* doc = nt + n;
        float* score = (float*)mb.pin_f32.p;
        rc = classify_packed_impl(h, n, bytes, off, st, label, nh, nt, doc, score);
        if (rc != BSP_OK) err = g_last_error;
        else
            for (i32 i = 0; i < n; i++) {
                OneRead& q = *batch[i];
                q.status = st[i]; q.label = label[i]; q.n_hits = nh[i]; q.n_top = nt[i];
                memcpy(q.doc, doc + (size_t)i * TOPN, sizeof q.doc);
                memcpy(q.score, score + (size_t)i * TOPN, sizeof q.score);
            }
    } catch (const BspError& e) { rc = e.code; err = e.what(); }
    for (i32 i = 0; i < n; i++) { batch[i]->rc = rc; batch[i]->err = err; }
}

BSP_API int bsp_classify_one(bsp_classifier* h, const char* seq, int32_t seq_len, int32_t* status, int32_t* label,
                             int32_t* n_hits, int32_t* n_top, int32_t* top_doc, float* top_score) {
    if (!h) return fail(nullptr, BSP_ERR_ARG, "handle is null");
    if (!seq || seq_len <= 0) return fail(&h->last_error, BSP_ERR_ARG, "sequence is null or empty");     // Classifier.java:342-344
    if (!status) return fail(&h->last_error, BSP_ERR_ARG, "null argument");
    MicroBatcher& mb = h->mb;
    OneRead rq;
    rq.seq = seq; rq.len = seq_len;
    std::unique_lock<std::mutex> lk(mb.mu);
    mb.pending.push_back(&rq);
    if (!mb.leader) {
        // Leader of this batch: wait for company until the window closes, and for as long as the previous batch
        // still owns the GPU (those reads would only queue on the handle's lock one by one otherwise).
        mb.leader = true;
        const auto deadline = std::chrono::steady_clock::now() + std::chrono::microseconds(mb.wait_us);
        while ((i32)mb.pending.size() < mb.max_reads) {
            // a closed loop of T callers comes back T at a time: once as many reads wait as the previous batch held,
            // the window would only add latency
            if (!mb.gpu_busy && mb.last_size > 0 && (i32)mb.pending.size() >= mb.last_size) break;
            if (mb.gpu_busy) mb.cv_collect.wait(lk);
            else if (mb.cv_collect.wait_until(lk, deadline) == std::cv_status::timeout) { if (!mb.gpu_busy) break; }
        }
        std::vector<OneRead*> batch;
        batch.swap(mb.pending);
        mb.last_size = (i32)batch.size();
        mb.leader = false;                 // the next arrival leads the next batch and collects while this one runs
        mb.gpu_busy++;
        lk.unlock();
        try { microbatch_run(h, batch); }
        catch (const std::exception& e) { for (OneRead* q : batch) { q->rc = BSP_ERR_NOMEM; q->err = e.what(); } }
        catch (...) { for (OneRead* q : batch) { q->rc = BSP_ERR_STATE; q->err = "micro-batch failed"; } }    // followers must wake up
        lk.lock();
        mb.gpu_busy--;
        mb.batches++;
        mb.reads += (i64)batch.size();
        for (OneRead* q : batch) q->done = true;
        mb.cv_done.notify_all();
        mb.cv_collect.notify_all();        // a leader that waited for the GPU may go now
    } else {
        if ((i32)mb.pending.size() >= mb.max_reads || (i32)mb.pending.size() == mb.last_size) mb.cv_collect.notify_all();
        mb.cv_done.wait(lk, [&] { return rq.done; });
    }
    lk.unlock();
    if (rq.rc != BSP_OK) return fail(&h->last_error, rq.rc, rq.err);
    *status = rq.status;
    if (label) *label = rq.label;
    if (n_hits) *n_hits = rq.n_hits;
    if (n_top) *n_top = rq.n_top;
    if (top_doc) memcpy(top_doc, rq.doc, sizeof rq.doc);
    if (top_score) memcpy(top_score, rq.score, sizeof rq.score);
    return BSP_OK;
}

BSP_API int bsp_classify_device(bsp_classifier* h, int32_t n, const void* d_seq_bytes, const void* d_seq_offsets,
                                int64_t total_bytes, int32_t max_len, void* d_status, void* d_label, void* d_n_hits,
                                void* d_n_top, void* d_top_doc, void* d_top_score) {
    (void)max_len;
    if (!h) return fail(nullptr, BSP_ERR_ARG, "handle is null");
    std::lock_guard<std::mutex> lk(h->mu);
    API_TRY(h)
    if (n < 0 || (n > 0 && (!d_seq_bytes || !d_seq_offsets || !d_status))) return fail(&h->last_error, BSP_ERR_ARG, "null argument");
    CUDA_CHECK(cudaSetDevice(h->conf.device));
    for (auto& c : h->counters) c = 0;
    h->score_ms = 0; h->pipe_ms = 0;
    BatchOut bo{(i32*)d_status, (i32*)d_label, (i32*)d_n_hits, (i32*)d_n_top, (i32*)d_top_doc, (float*)d_top_score};
    if (n) classify_on_device(h, n, (const u8*)d_seq_bytes, (const i64*)d_seq_offsets, total_bytes, bo);
    return BSP_OK;
    API_CATCH(h)
}

// ------------------------------------------------------------------ introspection
BSP_API int bsp_doc_fields(bsp_classifier* h, int32_t doc, const char** filename, const char** header,
                           const char** direction, const char** taxonomy) {
    if (!h) return fail(nullptr, BSP_ERR_ARG, "handle is null");
    // ids are the ones classify/merge report: global in doc-partitioned mode (bsp_partition_set_global), local otherwise
    const i64 ldoc = (i64)doc - h->doc_base_global;
    if (ldoc < 0 || ldoc >= (i64)h->ix.docs.size())
        return fail(&h->last_error, BSP_ERR_ARG, "doc id out of range (in doc-partitioned mode: not owned by this partition)");
    const DocMeta& d = h->ix.docs[ldoc];
    const EntryMeta& m = h->ref->entries[d.entry];
    static const char* DIR[3] = {"forward", "reverse", "min_strand"};       // Indexer.java:211,216,221
    if (filename) *filename = m.filename.c_str();
    if (header) *header = m.header.c_str();
    if (direction) *direction = DIR[d.strand];
    if (taxonomy) *taxonomy = m.taxonomy.c_str();
    return BSP_OK;
}

BSP_API int bsp_index_stats(bsp_classifier* h, int64_t* n_docs, int64_t* n_terms, int64_t* n_postings, int64_t* n_positions) {
    if (!h) return fail(nullptr, BSP_ERR_ARG, "handle is null");
    if (n_docs) *n_docs = (i64)h->ix.docs.size();
    if (n_terms) *n_terms = h->ix.n_terms;
    if (n_postings) *n_postings = h->ix.n_postings;
    if (n_positions) *n_positions = h->ix.n_positions;
    return BSP_OK;
}

BSP_API int bsp_doc_info(bsp_classifier* h, int32_t doc, int32_t* num_tokens, int32_t* norm_byte) {
    if (!h) return fail(nullptr, BSP_ERR_ARG, "handle is null");
    const i64 ldoc = (i64)doc - h->doc_base_global;          // same id space as bsp_doc_fields
    if (ldoc < 0 || ldoc >= (i64)h->ix.docs.size())
        return fail(&h->last_error, BSP_ERR_ARG, "doc id out of range (in doc-partitioned mode: not owned by this partition)");
    if (num_tokens) *num_tokens = (i32)h->ix.doc_ntok[ldoc];
    if (norm_byte) *norm_byte = h->ix.doc_norm[ldoc];
    return BSP_OK;
}

__global__ void lookup_term_kernel(const SegDev* segs, int n_segs, u64 key, u32* out /* [n_segs*3]: t, a, b */) {
    int s = threadIdx.x;
    if (s >= n_segs) return;
    u32 t = seg_find_term(segs[s], key);
    out[3 * s] = t;
    out[3 * s + 1] = t == NO_TERM ? 0 : segs[s].term_off[t];
    out[3 * s + 2] = t == NO_TERM ? 0 : segs[s].term_off[t + 1];
}

BSP_API int bsp_term_postings(bsp_classifier* h, uint64_t packed_kmer, int64_t* df, int64_t* ttf,
                              int32_t* docs, int32_t* tfs, int32_t* positions) {
    if (!h) return fail(nullptr, BSP_ERR_ARG, "handle is null");
    std::lock_guard<std::mutex> lk(h->mu);
    API_TRY(h)
    GpuIndex& ix = h->ix;
    if (!ix.positional) return fail(&h->last_error, BSP_ERR_STATE, "positional index is not resident");
    CUDA_CHECK(cudaSetDevice(h->conf.device));
    if (2 * ix.k < 64 && (packed_kmer >> (2 * ix.k))) { if (df) *df = 0; if (ttf) *ttf = 0; return BSP_OK; }
    i64 ndf = 0, nttf = 0;
    for (size_t si = 0; si < ix.segs.size(); si++) {
        Segment& seg = ix.segs[si];
        DevBuf<u32> d_out; d_out.alloc(3);
        lookup_term_kernel<<<1, 32, 0, h->st>>>(ix.d_segs.p + si, 1, packed_kmer, d_out.p);
        u32 o[3];
        CUDA_CHECK(cudaMemcpyAsync(o, d_out.p, 12, cudaMemcpyDeviceToHost, h->st));
        CUDA_CHECK(cudaStreamSynchronize(h->st));
        if (o[0] == NO_TERM) continue;
        u32 a = o[1], b = o[2], np = b - a;
        std::vector<u32> pd(np), pp(np + 1);
        CUDA_CHECK(cudaMemcpy(pd.data(), seg.post_doc.p + a, np * 4, cudaMemcpyDeviceToHost));
        CUDA_CHECK(cudaMemcpy(pp.data(), seg.post_pos.p + a, (np + 1) * 4, cudaMemcpyDeviceToHost));
        u32 npos = pp[np] - pp[0];
        if (docs && tfs) for (u32 i = 0; i < np; i++) { docs[ndf + i] = (i32)(seg.doc_base + pd[i]); tfs[ndf + i] = (i32)(pp[i + 1] - pp[i]); }
        if (positions) {
            // the device keeps segment ordinals: Lucene's position = ordinal - first ordinal of the doc
            CUDA_CHECK(cudaMemcpy(positions + nttf, seg.positions.p + pp[0], (size_t)npos * 4, cudaMemcpyDeviceToHost));
            std::vector<u32> dtok((size_t)seg.n_docs + 1);
            CUDA_CHECK(cudaMemcpy(dtok.data(), seg.doc_tok.p, dtok.size() * 4, cudaMemcpyDeviceToHost));
            for (u32 i = 0; i < np; i++)
                for (u32 q = pp[i] - pp[0]; q < pp[i + 1] - pp[0]; q++) positions[nttf + q] -= (i32)dtok[pd[i]];
        }
        ndf += np; nttf += npos;
    }
    if (df) *df = ndf;
    if (ttf) *ttf = nttf;
    return BSP_OK;
    API_CATCH(h)
}

BSP_API int64_t bsp_tokenize(bsp_classifier* h, const char* seq, int32_t len, uint64_t* keys, int32_t* offsets, int64_t cap) {
    if (!h) return fail(nullptr, BSP_ERR_ARG, "handle is null");
    std::lock_guard<std::mutex> lk(h->mu);
    API_TRY(h)
    if (len < 0 || (len > 0 && !seq)) return fail(&h->last_error, BSP_ERR_ARG, "sequence is null");
    CUDA_CHECK(cudaSetDevice(h->conf.device));
    cudaStream_t st = h->st;
    DevBuf<u8> d_seq; d_seq.alloc((size_t)len + 64);
    DevBuf<i64> d_off; d_off.alloc(2);
    DevBuf<i64> d_tok_off; d_tok_off.alloc(2);
    DevBuf<u64> d_keys; d_keys.alloc((size_t)len + 1);
    DevBuf<u32> d_pos; d_pos.alloc((size_t)len + 1);
    DevBuf<i32> d_n; d_n.alloc(1);
    i64 off[2] = {0, len}, tz[2] = {0, 0};
    CUDA_CHECK(cudaMemcpyAsync(d_seq.p, seq, (size_t)len, cudaMemcpyHostToDevice, st));
    CUDA_CHECK(cudaMemcpyAsync(d_off.p, off, 16, cudaMemcpyHostToDevice, st));
    CUDA_CHECK(cudaMemcpyAsync(d_tok_off.p, tz, 16, cudaMemcpyHostToDevice, st));
    tokenize_kernel<<<1, 32, 0, st>>>(d_seq.p, d_off.p, 1, d_tok_off.p, h->ix.k, h->conf.kmer_skips, h->ix.min_strand, d_keys.p, d_pos.p, d_n.p);
    KERNEL_CHECK();
    i32 n = 0;
    CUDA_CHECK(cudaMemcpyAsync(&n, d_n.p, 4, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    i64 m = std::min<i64>(n, cap);
    if (keys && m > 0) CUDA_CHECK(cudaMemcpy(keys, d_keys.p, (size_t)m * 8, cudaMemcpyDeviceToHost));
    if (offsets && m > 0) CUDA_CHECK(cudaMemcpy(offsets, d_pos.p, (size_t)m * 4, cudaMemcpyDeviceToHost));
    return n;
    API_CATCH(h)
}

BSP_API int bsp_last_routes(bsp_classifier* h, int32_t n, int32_t* routes) {
    if (!h || !routes) return fail(nullptr, BSP_ERR_ARG, "null argument");
    std::lock_guard<std::mutex> lk(h->mu);
    for (i32 i = 0; i < n; i++) routes[i] = i < (i32)h->last_routes.size() ? h->last_routes[i] : 0;
    return BSP_OK;
}

BSP_API int bsp_last_counters(bsp_classifier* h, int64_t counters[8]) {
    if (!h || !counters) return fail(nullptr, BSP_ERR_ARG, "null argument");
    std::lock_guard<std::mutex> lk(h->mu);
    for (int i = 0; i < 8; i++) counters[i] = h->counters[i];
    return BSP_OK;
}

BSP_API int bsp_last_timings(bsp_classifier* h, double* score_kernel_ms, double* pipeline_ms) {
    if (!h) return fail(nullptr, BSP_ERR_ARG, "null argument");
    std::lock_guard<std::mutex> lk(h->mu);
    if (score_kernel_ms) *score_kernel_ms = h->score_ms;
    if (pipeline_ms) *pipeline_ms = h->pipe_ms;
    return BSP_OK;
}

BSP_API int bsp_last_gap_join(bsp_classifier* h, int64_t out[4], double* join_kernel_ms, double* verify_kernel_ms, int64_t* candidates) {
    if (!h || !out) return fail(nullptr, BSP_ERR_ARG, "null argument");
    std::lock_guard<std::mutex> lk(h->mu);
    out[0] = h->gj_clauses; out[1] = h->gj_chunks; out[2] = h->gj_matches; out[3] = h->gj_bytes;
    if (join_kernel_ms) *join_kernel_ms = h->gj_ms;
    if (verify_kernel_ms) *verify_kernel_ms = h->gj_verify_ms;
    if (candidates) *candidates = h->gj_cands;
    return BSP_OK;
}

BSP_API int bsp_memory_info(bsp_classifier* h, int64_t* positional_bytes, int64_t* table_bytes, int64_t* staged_bytes) {
    if (!h) return fail(nullptr, BSP_ERR_ARG, "null argument");
    if (positional_bytes) *positional_bytes = h->ix.positional ? (i64)h->ix.positional_bytes : 0;
    if (table_bytes) *table_bytes = (i64)h->ix.table_bytes;
    if (staged_bytes) *staged_bytes = (i64)(h->ref->packed.bytes() + h->ref->mask.bytes());
    return BSP_OK;
}

// Cost model of SURVEY.md 8(d) on the actual index, for a read set (host buffers).
BSP_API int bsp_read_cost(bsp_classifier* h, int32_t n, const char* seq_bytes, const int64_t* seq_offsets, int64_t out[8]) {
    if (!h || !out) return fail(nullptr, BSP_ERR_ARG, "null argument");
    std::lock_guard<std::mutex> lk(h->mu);
    API_TRY(h)
    for (int i = 0; i < 8; i++) out[i] = 0;
    if (n <= 0) return BSP_OK;
    if (!seq_bytes || !seq_offsets) return fail(&h->last_error, BSP_ERR_ARG, "null argument");
    CUDA_CHECK(cudaSetDevice(h->conf.device));
    GpuIndex& ix = h->ix;
    Scratch& sc = h->sc[0];
    cudaStream_t st = h->st;
    if (!ix.direct && !ix.positional) return fail(&h->last_error, BSP_ERR_STATE, "no term statistics resident");
    DevBuf<unsigned long long> d_out; d_out.alloc(8);
    CUDA_CHECK(cudaMemsetAsync(d_out.p, 0, 64, st));
    const i32 SUB = 262144;
    for (i32 r0 = 0; r0 < n; r0 += SUB) {
        i32 m = std::min(SUB, n - r0);
        i64 b0 = seq_offsets[r0], bytes = seq_offsets[r0 + m] - b0;
        std::vector<i64> off((size_t)m + 1);
        for (i32 i = 0; i <= m; i++) off[i] = seq_offsets[r0 + i] - b0;
        sc.seq.ensure((size_t)bytes + 64); sc.seq_off.ensure((size_t)m + 1); sc.tok_off.ensure((size_t)m + 1);
        sc.tok_cap.ensure((size_t)m + 1); sc.n_tok.ensure(m);
        sc.tok_key.ensure((size_t)bytes + 1); sc.tok_pos.ensure((size_t)bytes + 1); sc.tok_df.ensure((size_t)bytes + 1);
        CUDA_CHECK(cudaMemcpyAsync(sc.seq_off.p, off.data(), (size_t)(m + 1) * 8, cudaMemcpyHostToDevice, st));
        if (bytes) CUDA_CHECK(cudaMemcpyAsync(sc.seq.p, seq_bytes + b0, (size_t)bytes, cudaMemcpyHostToDevice, st));
        tok_capacity_kernel<<<(unsigned)ceil_div(m, 256), 256, 0, st>>>(sc.seq_off.p, m, ix.k, sc.tok_cap.p);
        exclusive_scan<i64, i64>(sc.tok_cap.p, sc.tok_off.p, (u64)m, sc.tok_off.p + m, sc.scan_tmp, st);
        tokenize_kernel<<<(unsigned)ceil_div((i64)m * 32, 256), 256, 0, st>>>(sc.seq.p, sc.seq_off.p, m, sc.tok_off.p, ix.k,
            h->conf.kmer_skips, ix.min_strand, sc.tok_key.p, sc.tok_pos.p, sc.n_tok.p);
        token_df_kernel<<<(unsigned)ceil_div((i64)m * 32, 256), 256, 0, st>>>(sc.tok_key.p, sc.tok_off.p, sc.n_tok.p, m,
            ix.direct ? ix.df_dense.p : nullptr, ix.d_segs.p, ix.positional ? (int)ix.segs.size() : 0, sc.tok_df.p);
        read_cost_kernel<<<(unsigned)ceil_div((i64)m * 32, 256), 256, 0, st>>>(sc.seq_off.p, m, sc.tok_off.p, sc.n_tok.p, sc.tok_key.p,
            sc.tok_df.p, ix.direct ? ix.ttf_dense.p : nullptr, ix.d_segs.p, ix.positional ? (int)ix.segs.size() : 0,
            h->conf.query_algorithm, ix.row_bytes, (u64)ix.docs.size(), d_out.p);
        KERNEL_CHECK();
        CUDA_CHECK(cudaStreamSynchronize(st));
    }
    unsigned long long ho[8];
    CUDA_CHECK(cudaMemcpy(ho, d_out.p, 64, cudaMemcpyDeviceToHost));
    for (int i = 0; i < 8; i++) out[i] = (i64)ho[i];
    return BSP_OK;
    API_CATCH(h)
}

// ------------------------------------------------------------------ doc-partitioned mode
BSP_API int bsp_partition_df_buffer(bsp_classifier* h, void** d_df, int64_t* n_entries) {
    if (!h || !d_df || !n_entries) return fail(nullptr, BSP_ERR_ARG, "null argument");
    if (!h->ix.direct) return fail(&h->last_error, BSP_ERR_UNSUPPORTED, "doc-partitioned mode needs the dense term space (k <= 12)");
    *d_df = h->ix.df_dense.p;
    *n_entries = (i64)1 << (2 * h->ix.k);
    return BSP_OK;
}

BSP_API int bsp_partition_local_stats(bsp_classifier* h, int64_t* n_docs, int64_t* sum_total_term_freq) {
    if (!h) return fail(nullptr, BSP_ERR_ARG, "null argument");
    if (n_docs) *n_docs = (i64)h->ix.docs.size();
    if (sum_total_term_freq) *sum_total_term_freq = h->ix.sum_ttf;
    return BSP_OK;
}

BSP_API int bsp_partition_set_global(bsp_classifier* h, int64_t doc_base, int64_t global_docs, int64_t global_sum_total_term_freq) {
    if (!h) return fail(nullptr, BSP_ERR_ARG, "null argument");
    std::lock_guard<std::mutex> lk(h->mu);
    API_TRY(h)
    CUDA_CHECK(cudaSetDevice(h->conf.device));
    h->doc_base_global = doc_base;
    h->global_docs = global_docs;
    h->global_sum_ttf = global_sum_total_term_freq;
    upload_scoring_constants(h);
    return BSP_OK;
    API_CATCH(h)
}

BSP_API int bsp_index_layout(bsp_classifier* h, int64_t out[8]) {
    if (!h || !out) return fail(nullptr, BSP_ERR_ARG, "null argument");
    const GpuIndex& ix = h->ix;
    out[0] = (i64)ix.row_bytes; out[1] = (i64)ix.d_pad; out[2] = (i64)ix.segs.size(); out[3] = (i64)ix.max_seg_docs;
    out[4] = ix.tables ? 1 : 0; out[5] = ix.positional ? 1 : 0; out[6] = ix.has_pair ? 1 : 0; out[7] = ix.direct ? 0 : 1;
    return BSP_OK;
}

BSP_API int bsp_pack_topk_device(bsp_classifier* h, int32_t n, const void* d_n_top, const void* d_top_doc, const void* d_top_score,
                                 void* d_records) {
    if (!h || !d_n_top || !d_top_doc || !d_top_score || !d_records) return fail(nullptr, BSP_ERR_ARG, "null argument");
    std::lock_guard<std::mutex> lk(h->mu);
    API_TRY(h)
    CUDA_CHECK(cudaSetDevice(h->conf.device));
    pack_lists_kernel<<<(unsigned)ceil_div(std::max<i64>((i64)n * LIST_REC_WORDS, 1), 256), 256, 0, h->st>>>(
        n, (const i32*)d_n_top, (const i32*)d_top_doc, (const float*)d_top_score, (i32*)d_records);
    KERNEL_CHECK();
    CUDA_CHECK(cudaStreamSynchronize(h->st));
    return BSP_OK;
    API_CATCH(h)
}

BSP_API int bsp_merge_topk_records_device(bsp_classifier* h, int32_t n, int32_t parts, const void* d_records, const void* d_part_status,
                                          void* d_status, void* d_label, void* d_n_hits, void* d_n_top, void* d_top_doc,
                                          void* d_top_score) {
    if (!h || !d_records || !d_part_status) return fail(nullptr, BSP_ERR_ARG, "null argument");
    std::lock_guard<std::mutex> lk(h->mu);
    API_TRY(h)
    CUDA_CHECK(cudaSetDevice(h->conf.device));
    merge_records_kernel<<<(unsigned)ceil_div(std::max(n, 1), 128), 128, 0, h->st>>>(n, parts, (const i32*)d_records,
        (const i32*)d_part_status, (i32*)d_status, (i32*)d_label, (i32*)d_n_hits, (i32*)d_n_top, (i32*)d_top_doc, (float*)d_top_score,
        (h->conf.full_topn || env_i64("BSP_FULL_TOPN", 0)) ? 1 : 0);
    KERNEL_CHECK();
    CUDA_CHECK(cudaStreamSynchronize(h->st));
    return BSP_OK;
    API_CATCH(h)
}

BSP_API int bsp_classifier_set_query(bsp_classifier* h, const bsp_config* query_conf) {
    if (!h || !query_conf) return fail(nullptr, BSP_ERR_ARG, "null argument");
    const int rc = check_query_conf(query_conf);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(h->mu);
    API_TRY(h)
    if (query_conf->kmer_size != h->conf.kmer_size || (query_conf->min_strand_kmer != 0) != (h->conf.min_strand_kmer != 0))
        throw BspError(BSP_ERR_ARG, "kmer_size / min_strand_kmer belong to the index: open another classifier for them");
    if (query_conf->query_algorithm != BSP_NAIVE_KMER && h->ix.tables && !h->ix.has_pair && !h->ix.positional)
        throw BspError(BSP_ERR_STATE, "the index was opened for NAIVE_KMER without pair tables or positional postings");
    CUDA_CHECK(cudaSetDevice(h->conf.device));
    for (int i = 0; i < BSP_SLOTS; i++) CUDA_CHECK(cudaStreamSynchronize(h->sc[i].st));
    h->conf.kmer_skips = query_conf->kmer_skips;
    h->conf.query_algorithm = query_conf->query_algorithm;
    h->conf.scoring_algorithm = query_conf->scoring_algorithm;
    h->conf.query_term_min_should_match = query_conf->query_term_min_should_match;
    h->conf.full_topn = query_conf->full_topn;
    upload_scoring_constants(h);
    return BSP_OK;
    API_CATCH(h)
}

BSP_API int bsp_merge_topk_device(bsp_classifier* h, int32_t n, int32_t parts, const void* d_part_n_top, const void* d_part_doc,
                                  const void* d_part_score, const void* d_part_status, void* d_status, void* d_label,
                                  void* d_n_hits, void* d_n_top, void* d_top_doc, void* d_top_score) {
    if (!h) return fail(nullptr, BSP_ERR_ARG, "null argument");
    std::lock_guard<std::mutex> lk(h->mu);
    API_TRY(h)
    CUDA_CHECK(cudaSetDevice(h->conf.device));
    // layout: [n][parts] lists; status is taken from part 0 (identical on every rank: it depends on the read only)
    merge_kernel<<<(unsigned)ceil_div(std::max(n, 1), 128), 128, 0, h->st>>>(n, nullptr, (const i32*)d_part_status, parts, parts, parts, parts,
        (const i32*)d_part_n_top, (const u32*)d_part_doc, (const float*)d_part_score, nullptr, (const i32*)d_part_n_top,
        (const u32*)d_part_doc, (const float*)d_part_score, (i32*)d_status, (i32*)d_label, (i32*)d_n_hits,
        (i32*)d_n_top, (i32*)d_top_doc, (float*)d_top_score, (h->conf.full_topn || env_i64("BSP_FULL_TOPN", 0)) ? 1 : 0);
    KERNEL_CHECK();
    CUDA_CHECK(cudaStreamSynchronize(h->st));
    return BSP_OK;
    API_CATCH(h)
}
