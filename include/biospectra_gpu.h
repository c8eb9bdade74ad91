/*
 * biospectra_gpu.h -- C ABI of libbiospectra_gpu.so (CUDA, sm_100a).
 *
 * Drop-in boundary for BioSpectra's k-mer index construction + read classification
 * hot path.  The reference has no FFI; the seam is the two Java classes
 *   biospectra.index.Indexer          (src/biospectra/index/Indexer.java:64,131,141,150,239)
 *   biospectra.classify.Classifier    (src/biospectra/classify/Classifier.java:71,341,377)
 * whose bodies (Lucene IndexWriter / IndexSearcher) this library replaces.  The JNI
 * stub a maintainer would add is shown in INTEGRATION.md; tests drive the same
 * functions through ctypes.
 *
 * Conventions: opaque handles; every function returns 0 on success, <0 on error
 * (bsp_last_error(handle) -> UTF-8 message, bsp_last_error(NULL) -> last error of
 * a failed create/open on this thread); inputs are caller-owned and only read during
 * the call; outputs go to caller-provided arrays; handles are thread-safe (calls are
 * serialised per handle).  No torch/CUDA types appear in any signature: "device
 * pointers" are passed as void* / uint64 addresses.
 */
#ifndef BIOSPECTRA_GPU_H
#define BIOSPECTRA_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BSP_OK 0
#define BSP_ERR_ARG (-1)      /* IllegalArgumentException in the reference */
#define BSP_ERR_IO (-2)       /* IOException */
#define BSP_ERR_CUDA (-3)     /* CUDA runtime failure */
#define BSP_ERR_STATE (-4)    /* call not valid in the handle's state */
#define BSP_ERR_UNSUPPORTED (-5)
#define BSP_ERR_NOMEM (-6)

#define BSP_TOPN 10           /* hitsPerPage, Classifier.java:349 */

/* query_algorithm (src/biospectra/classify/QueryGenerationAlgorithm.java) */
#define BSP_NAIVE_KMER 0
#define BSP_CHAIN_PROXIMITY 1
#define BSP_PAIRED_PROXIMITY 2
/* scoring_algorithm (src/biospectra/Configuration.java:169-179) */
#define BSP_TFIDF 0           /* "default" | "tfidf" | "vectorspace" | anything unknown */
#define BSP_BM25 1

/* per-read status */
#define BSP_READ_OK 0
#define BSP_READ_DROPPED 1    /* < 2 tokens or > 10000 clauses: the reference throws and
                                 LocalClassifier.java:112-114 logs + skips the read */
#define BSP_READ_ERROR 2      /* could not be evaluated (see bsp_last_error) */
/* per-read label, score tier only (taxonomy/LCA is string work done by the host) */
#define BSP_UNKNOWN 0
#define BSP_VAGUE 1           /* >= 2 hits tie at the max score */
#define BSP_CLASSIFIED 2      /* exactly 1 hit at the max score */

/* Mirrors configuration.json 1:1 (src/biospectra/Configuration.java:37-44,78-166),
 * plus GPU placement, which the reference has no notion of. */
typedef struct bsp_config {
    int32_t kmer_size;                  /* "kmer_size", default 10 */
    int32_t kmer_skips;                 /* "kmer_skips" (queries only), code default 5, shipped json 0 */
    int32_t min_strand_kmer;            /* "min_strand_kmer" 0/1 */
    int32_t query_algorithm;            /* "query_algorithm" BSP_* */
    int32_t scoring_algorithm;          /* "scoring_algorithm" BSP_TFIDF / BSP_BM25 */
    int32_t worker_threads;             /* accepted (must be > 0 for the indexer), advisory */
    int32_t index_ram_buffer;           /* accepted, ignored */
    int32_t device;                     /* CUDA device ordinal of this handle */
    double  query_term_min_should_match;/* "query_term_min_should_match", default 0.5 */
    /* doc-partitioned mode: this handle holds the part_rank-th of part_count contiguous
     * slices of the reference entries (0/1 or 0/0 = whole index) */
    int32_t part_rank;
    int32_t part_count;
    /* index residency: 0 auto, 1 force on, 2 force off */
    int32_t dense_tables;               /* dense tf / pair-frequency tables (small k) */
    int32_t positional;                 /* term-major positional postings (general path) */
    /* 0: report the reference's result set only -- the hits whose score equals the maximum (what
     *    Classifier.classify returns, Classifier.java:358-366); n_top == n_hits.
     * 1: also report the rest of the top-10 list Lucene's collector held (verification, diagnostics). */
    int32_t full_topn;
} bsp_config;

typedef struct bsp_indexer bsp_indexer;
typedef struct bsp_classifier bsp_classifier;

void bsp_config_default(bsp_config* c);         /* Configuration.java:37-44 defaults */
const char* bsp_last_error(const void* handle);
const char* bsp_version(void);

/* ------------------------------------------------------------------ indexing
 * Indexer(Configuration) -> bsp_index_create  (Indexer.java:64-129: validates k>0,
 *   worker_threads>0, index_path!=NULL; creates index_path and WIPES its contents)
 * Indexer.index(File,File) worker (Indexer.java:178-232) -> bsp_index_add_entry,
 *   one call per FASTA entry: 2 documents (forward, reverse) or 1 (min_strand)
 * Indexer.close() (Indexer.java:239-251) -> bsp_index_close: builds on the GPU,
 *   writes <index_path>/biospectra.idx atomically, frees the handle. */
int bsp_index_create(const bsp_config* conf, const char* index_path, bsp_indexer** out);
/* seq: the entry's sequence as FASTAReader returns it (upper-cased, trimmed; A.1).
 * header: header line without its leading '>' (Indexer.java:167-170). taxonomy: whole
 * text of the .taxd sidecar, "" if none.  Errors here are per-entry: the reference
 * logs and skips (Indexer.java:227-229). */
int bsp_index_add_entry(bsp_indexer* h, const char* filename, const char* header, const char* taxonomy,
                        const char* seq, int64_t seq_len);
/* Indexer.index(File fastaDoc, File taxonDoc) (Indexer.java:150-236) in one call: the library's own FASTA reader
 * (org.yeastrc.fasta.FASTAReader#readNext semantics, gzip by file name) + one bsp_index_add_entry per entry with
 * filename = the file's base name; taxd_path may be NULL or missing on disk ("" taxonomy).  A malformed file returns
 * BSP_ERR_IO after the entries before the error were added (message: bsp_host_last_error()). */
int bsp_index_add_fasta(bsp_indexer* h, const char* fasta_path, const char* taxd_path_or_null);
/* same as bsp_index_add_entry, with the sequence already resident in device memory (uint8 ASCII) */
int bsp_index_add_entry_device(bsp_indexer* h, const char* filename, const char* header, const char* taxonomy,
                               const void* d_seq, int64_t seq_len);
int bsp_index_close(bsp_indexer* h);
/* close without writing the index file and hand the staged reference to a classifier
 * (used by benchmarks to avoid a disk round trip); frees the indexer */
int bsp_index_close_to_classifier(bsp_indexer* h, const bsp_config* query_conf, bsp_classifier** out);

/* ------------------------------------------------------------- classification
 * Classifier(Configuration) -> bsp_classifier_open (Classifier.java:71-111: validates
 *   k>0, skips>=0, index dir exists); loads biospectra.idx and builds the GPU index.
 * Classifier.classify(header, sequence) (Classifier.java:341-374), batched:
 *   bsp_classify_batch.  Per read i:
 *     status[i]  BSP_READ_*; label[i] BSP_UNKNOWN/VAGUE/CLASSIFIED;
 *     n_hits[i]  #hits whose score equals the max (the reference's result list, <= 10);
 *     n_top[i]   #entries reported in the list below: n_hits[i], or with conf.full_topn the whole top-10 list
 *                (score desc, doc asc);
 *     top_doc[i*10+j], top_score[i*10+j]  the list; rank of entry j is j.
 *   Any of the output pointers except status may be NULL. */
int bsp_classifier_open(const bsp_config* conf, const char* index_path, bsp_classifier** out);
int bsp_classify_batch(bsp_classifier* h, int32_t n, const char* const* seqs, const int32_t* seq_lens,
                       int32_t* status, int32_t* label, int32_t* n_hits, int32_t* n_top,
                       int32_t* top_doc, float* top_score);
/* Same with reads packed back to back: seq_offsets[n+1] into seq_bytes (host memory,
 * ideally pinned).  This is the call the batch driver (LocalClassifier) uses. */
int bsp_classify_packed(bsp_classifier* h, int32_t n, const char* seq_bytes, const int64_t* seq_offsets,
                        int32_t* status, int32_t* label, int32_t* n_hits, int32_t* n_top,
                        int32_t* top_doc, float* top_score);
/* Taxonomy-aware labels on the device: Classifier.makeClassificationResult (classify/Classifier.java:263-339) with
 * the lineages interned once.  The host parses each document's taxon_hierarchy JSON and passes the taxids of
 * TaxonTreeDescription#getClassifiableTaxonomyTree (beans/TaxonTreeDescription.java:33-98: the entries with a rank,
 * leaf -> root): lineage_off[n_docs + 1] into taxid[]; flags[d] bit0 = the document has a taxonomy string, bit1 = the
 * string is malformed (the reference throws while parsing it and its caller skips the read).  n_docs must equal the
 * document count of the index (the global count in doc-partitioned mode). */
int bsp_classifier_set_lineages(bsp_classifier* h, int32_t n_docs, const int64_t* lineage_off,
                                const int32_t* taxid, const uint8_t* flags);
/* bsp_classify_packed plus, per read: tax_label = BSP_UNKNOWN / BSP_VAGUE / BSP_CLASSIFIED as the reference labels the
 * read with taxonomy (-1: a hit's taxonomy is malformed, the reference throws), tax_index = position in hit 0's ranked
 * lineage of the taxon the read is classified at (-1: taxon ("unknown", "")).  tax_index may be NULL. */
int bsp_classify_packed_tax(bsp_classifier* h, int32_t n, const char* seq_bytes, const int64_t* seq_offsets,
                            int32_t* status, int32_t* label, int32_t* n_hits, int32_t* n_top,
                            int32_t* top_doc, float* top_score, int32_t* tax_label, int32_t* tax_index);
/* One read, blocking, for concurrent callers: replaces the body of
 * Classifier.classify(header, sequence) (classify/Classifier.java:341-366), which the
 * reference calls from `worker_threads` pool threads at once (LocalClassifier.java:96-118,
 * ClassifierServer.java:70-98).  Calls that arrive within the micro-batch window of each
 * other -- or while the GPU is busy with the previous micro-batch -- are coalesced into one
 * bsp_classify_packed; every caller gets exactly its own read's result.  top_doc / top_score
 * point at 10 entries each (may be NULL).  seq == NULL or seq_len <= 0 returns BSP_ERR_ARG
 * (IllegalArgumentException, Classifier.java:342-344). */
int bsp_classify_one(bsp_classifier* h, const char* seq, int32_t seq_len,
                     int32_t* status, int32_t* label, int32_t* n_hits, int32_t* n_top,
                     int32_t* top_doc, float* top_score);
/* micro-batch window of bsp_classify_one: at most max_reads per GPU batch (default 4096),
 * the first caller of a batch waits at most max_wait_us for company (default 200; 0 = no
 * waiting, reads still coalesce while the GPU is busy).  Environment BSP_MICROBATCH_MAX /
 * BSP_MICROBATCH_US give the defaults. */
int bsp_classifier_set_microbatch(bsp_classifier* h, int32_t max_reads, int64_t max_wait_us);
/* counters since open: GPU batches run and reads served by bsp_classify_one */
int bsp_microbatch_stats(bsp_classifier* h, int64_t* batches, int64_t* reads);
/* Device-resident variant: d_seq_bytes / d_seq_offsets and all outputs are device
 * pointers; nothing crosses PCIe.  total_bytes = seq_offsets[n]. */
int bsp_classify_device(bsp_classifier* h, int32_t n, const void* d_seq_bytes, const void* d_seq_offsets,
                        int64_t total_bytes, int32_t max_len,
                        void* d_status, void* d_label, void* d_n_hits, void* d_n_top,
                        void* d_top_doc, void* d_top_score);
/* New query-side settings on an open classifier (kmer_skips, query_algorithm, scoring_algorithm,
 * query_term_min_should_match, full_topn): what constructing a second Classifier(Configuration) on the same index
 * directory is in the reference (Classifier.java:71-111 -- the IndexReader is reopened, the index is not rebuilt).
 * kmer_size / min_strand_kmer belong to the index and must match (BSP_ERR_ARG).  Waits for running batches. */
int bsp_classifier_set_query(bsp_classifier* h, const bsp_config* query_conf);
/* stored fields of a document (IndexSearcher.doc(), Classifier.java:361-363);
 * borrowed pointers, valid until close.  `doc` is an id as classify / merge report it: after
 * bsp_partition_set_global that is the GLOBAL id, and the call returns BSP_ERR_ARG for a document another
 * partition owns (bsp_doc_info follows the same rule). */
int bsp_doc_fields(bsp_classifier* h, int32_t doc, const char** filename, const char** header,
                   const char** direction, const char** taxonomy);
int bsp_classifier_close(bsp_classifier* h);

/* ------------------------------------------------ introspection (parity tests) */
int bsp_index_stats(bsp_classifier* h, int64_t* n_docs, int64_t* n_terms, int64_t* n_postings,
                    int64_t* n_positions);
int bsp_doc_info(bsp_classifier* h, int32_t doc, int32_t* num_tokens, int32_t* norm_byte);
/* postings of one term (packed 2-bit k-mer, first base most significant, canonical if
 * min_strand): pass NULL arrays to query sizes first.  docs/tfs hold *df entries,
 * positions holds sum(tfs) token ordinals, grouped by doc. Needs the positional index. */
int bsp_term_postings(bsp_classifier* h, uint64_t packed_kmer, int64_t* df, int64_t* ttf,
                      int32_t* docs, int32_t* tfs, int32_t* positions);
/* tokens of a read exactly as the query analyzer emits them (KmerQueryAnalyzer):
 * returns the count; keys/offsets may be NULL */
int64_t bsp_tokenize(bsp_classifier* h, const char* seq, int32_t len, uint64_t* keys, int32_t* offsets,
                     int64_t cap);
/* which scoring path handled read i of the last batch: 1 dense tables scored exactly for every doc,
 * 2 positional postings, 3 dense tables through the bound-and-rescore filter */
int bsp_last_routes(bsp_classifier* h, int32_t n, int32_t* routes);
/* counters of the last batch: [0] reads, [1] reads scored by the exact dense kernel, [2] positional-path reads,
 * [3] reads finished by the filter kernel, [4] kernels launched, [5] docs the filter rescored exactly,
 * [6] exception cells met by the filter, [7] reads the filter handed to the exact kernel */
int bsp_last_counters(bsp_classifier* h, int64_t counters[8]);
/* kernel time (ms, CUDA events on the handle's stream) of the dominant scoring kernel
 * and of the whole device pipeline in the last batch */
int bsp_last_timings(bsp_classifier* h, double* score_kernel_ms, double* pipeline_ms);
/* gap join of the last batch (phrase clauses whose k-mers are not adjacent in the read -- every phrase clause with
 * kmer_skips > 0, Configuration.java:37-44 / Classifier.java:160-200 -- evaluated against the positional index, sorted by
 * first k-mer): out[0] gap clauses, [1] chunks (clauses sharing a first k-mer, one warp each), [2] (clause, document)
 * matches, [3] position bytes the join kernel fetched; the join kernel's and the per-candidate verification kernel's time
 * in ms (CUDA events on the handle's streams); candidates = occurrences of the second k-mer with an occurrence of the
 * first one inside the window (before the per-document check).  The last three pointers may be null. */
int bsp_last_gap_join(bsp_classifier* h, int64_t out[4], double* join_kernel_ms, double* verify_kernel_ms, int64_t* candidates);
/* Cost model of SURVEY.md 8(d) evaluated on the actual index for a read set (host
 * buffers as in bsp_classify_packed): out[0] reads that are scored, [1] tokens, [2] unique
 * query terms (= dictionary probes), [3] sum df, [4] sum ttf over unique terms,
 * [5] A_read bytes (CSR formula), [6] bytes the dense-table layout reads, [7] clauses */
int bsp_read_cost(bsp_classifier* h, int32_t n, const char* seq_bytes, const int64_t* seq_offsets, int64_t out[8]);
/* layout of the resident index: [0] bytes per dense-table row, [1] padded doc count of a row, [2] positional segments,
 * [3] most docs in one segment (<= 256: the warp-per-(segment, read) positional scorer runs, else the CTA one),
 * [4] dense tables resident, [5] positional postings resident, [6] pair rows built, [7] dictionary: 0 direct table, 1 hash */
int bsp_index_layout(bsp_classifier* h, int64_t out[8]);
int bsp_memory_info(bsp_classifier* h, int64_t* positional_bytes, int64_t* table_bytes, int64_t* staged_bytes);

/* --------------------------------------------- doc-partitioned mode plumbing
 * Global statistics must be identical on every rank (idf / avgdl are global in
 * Lucene: IndexSearcher#termStatistics,#collectionStatistics).  The per-term df array
 * is exposed as a device pointer so the host can all-reduce it in place over NCCL
 * (dense term space only, 4^k entries of uint32). */
int bsp_partition_df_buffer(bsp_classifier* h, void** d_df, int64_t* n_entries);
int bsp_partition_local_stats(bsp_classifier* h, int64_t* n_docs, int64_t* sum_total_term_freq);
int bsp_partition_set_global(bsp_classifier* h, int64_t doc_base, int64_t global_docs,
                             int64_t global_sum_total_term_freq);
/* merge per-rank top-10 lists (n reads x parts x 10, device pointers; docs are global
 * ids) into final outputs (device pointers, same meaning as bsp_classify_device) */
int bsp_merge_topk_device(bsp_classifier* h, int32_t n, int32_t parts,
                          const void* d_part_n_top, const void* d_part_doc, const void* d_part_score,
                          const void* d_part_status,
                          void* d_status, void* d_label, void* d_n_hits, void* d_n_top,
                          void* d_top_doc, void* d_top_score);

/* The same exchange with ONE collective: bsp_pack_topk_device turns a rank's lists into one 84-byte record per read
 * (int32 n_top | int32 doc[10], global ids | float score[10]); the host all-gathers the record arrays rank-major
 * (records[(rank * n + read) * 21], e.g. ncclAllGather / all_gather_into_tensor) and bsp_merge_topk_records_device merges
 * them where the collective left them.  All pointers are device pointers. */
int bsp_pack_topk_device(bsp_classifier* h, int32_t n, const void* d_n_top, const void* d_top_doc, const void* d_top_score,
                         void* d_records);
int bsp_merge_topk_records_device(bsp_classifier* h, int32_t n, int32_t parts, const void* d_records, const void* d_part_status,
                                  void* d_status, void* d_label, void* d_n_hits, void* d_n_top, void* d_top_doc,
                                  void* d_top_score);

/* ------------------------------------------------ host-side entry points
 * The reference's drivers above the two classes, restated in C++ (no JVM here):
 *   BioSpectra.index         (src/biospectra/BioSpectra.java:110-134)  -> bsp_host_index
 *   BioSpectra.classifyLocal (src/biospectra/BioSpectra.java:136-161)  -> bsp_host_classify_local
 * json_conf: path of a configuration.json, or NULL to use defaults + index_dir.  They find
 * FASTA files like FastaFileHelper.findFastaDocs (sorted by path), read them with
 * FASTAReader semantics, pick up "<fasta>.taxd" sidecars, and write "<name>.result" /
 * "<name>.result.sum" in the reference's JSON formats. */
int bsp_host_index(const char* json_conf, const char* index_dir, const char* reference_dir, int32_t device);
int bsp_host_classify_local(const char* json_conf, const char* index_dir, const char* input_dir,
                            const char* output_dir, int32_t device);
/* bsp_host_classify_local keeps the files but batches reads itself (reader thread -> GPU ->
 * formatter threads, lines in read order).  The _per_read variant keeps the reference's control
 * flow too: `threads` pool threads (0 = conf.worker_threads) call Classifier.classify(header,
 * sequence) once per read (LocalClassifier.java:88-120) and write lines in completion order;
 * the concurrent calls meet in bsp_classify_one's micro-batches. */
int bsp_host_classify_local_per_read(const char* json_conf, const char* index_dir, const char* input_dir,
                                     const char* output_dir, int32_t device, int32_t threads);
/* LocalClassifier.classify(File inputFasta, File classifyOutput, File summaryOutput) (LocalClassifier.java:60-131) over an
 * index that is already open: reader thread -> GPU batches -> `worker_threads` formatter threads -> in-order writer.  The
 * handle stays open.  summary_path may be NULL. */
int bsp_host_classify_file(bsp_classifier* h, const char* input_fasta, const char* result_path, const char* summary_path,
                           int32_t worker_threads);
const char* bsp_host_last_error(void);
/* pure host helpers (no GPU): used by the CPU-side tests */
int bsp_host_label_json(const char* const* taxon_hierarchies, int32_t n, char* out, int32_t cap);
int bsp_host_read_fasta(const char* path, char* out, int64_t cap, int64_t* needed);
/* the same file read in stripes of ~stripe_bytes cut at entry starts, the way the batch driver's parser threads read it:
 * byte-identical output and return code */
int bsp_host_read_fasta_striped(const char* path, int64_t stripe_bytes, char* out, int64_t cap, int64_t* needed);
int bsp_host_double_to_string(double d, char* out, int32_t cap);
int bsp_host_parse_config(const char* json, bsp_config* out, char* index_path, int32_t cap);

#ifdef __cplusplus
}
#endif
#endif /* BIOSPECTRA_GPU_H */
