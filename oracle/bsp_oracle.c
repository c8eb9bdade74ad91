/*
 * bsp_oracle.c -- fast CPU oracle.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C (OpenMP) restatement of the BioSpectra k-mer index + classify hot
 * path and of the Lucene 5.3.1 arithmetic it delegates to.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product (biospectra_b200/) never does.
 *
 * PARITY PINNED BY EXECUTION: the reference ships no tests or golden outputs and no JVM
 * exists here; oracle/minijvm.py executes the reference's own lucene-core-5.3.1.jar /
 * fasta-utils-1.1.jar and tests/golden/lucene_*.json holds what they returned.  This file
 * must reproduce those vectors bit for bit (tests/test_lucene_golden.py), agree bit for bit
 * with the literal Python spec oracle (oracle/spec_oracle.py) on randomized corpora
 * (tests/test_oracle_cross.py) and with SURVEY.md Appendix B (tests/test_oracle_kat.py).
 *
 * Citations: paths are relative to /root/reference; "lucene!X#m" means class X,
 * method m of libs/lucene-core-5.3.1.jar (restated from the class files,
 * SURVEY.md Appendix A/D).
 *
 * Build: see oracle/Makefile (-O2 -fopenmp -ffp-contract=off: no FMA contraction,
 * every float op rounds like a Java float).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORA_API __attribute__((visibility("default")))

enum { ALG_NAIVE = 0, ALG_CHAIN = 1, ALG_PAIRED = 2 };
enum { SIM_TFIDF = 0, SIM_BM25 = 1 };
#define MAX_CLAUSES 10000 /* classify/Classifier.java:110 BooleanQuery.setMaxClauseCount(10000) */
#define TOPN 10           /* classify/Classifier.java:349 hitsPerPage */

/* ------------------------------------------------------------------ encoding */
/* utils/SequenceHelper.java:29-31 convCharToBitLUT: A0 C1 G2 T3 */
static inline int base_code(unsigned char c) {
    switch (c) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; default: return -1; }
}

/* reverse complement of a packed k-mer (first base most significant) */
static inline uint64_t rc_packed(uint64_t v, int k) {
    uint64_t r = 0;
    for (int i = 0; i < k; i++) { r = (r << 2) | (3 - (v & 3)); v >>= 2; }
    return r;
}

/* lucene/SequenceCompressFilter.java:56-68: String.compareTo over A<C<G<T equals the
 * numeric order of the MSB-first 2-bit code. */
static inline uint64_t canon(uint64_t v, int k, int min_strand) {
    if (!min_strand) return v;
    uint64_t r = rc_packed(v, k);
    return v > r ? r : v;
}

/* lucene/KmerSequenceTokenizer.java:115-154.  Returns number of tokens (counting
 * beyond cap without storing). */
ORA_API int64_t ora_tokenize(const char* seq, int64_t len, int k, int skips, int min_strand,
                             uint64_t* keys, int32_t* offs, int64_t cap) {
    int64_t n = 0, pos = 0;
    if (k < 1 || k > 32 || skips < 0) return -1;
    int64_t bad_until = -1; /* index of the last invalid char seen so far in [.., pos+k) */
    int64_t scanned = 0;    /* chars [0, scanned) have been examined */
    for (;;) {
        int cur_skip = skips;
        int emitted = 0;
        for (;;) {
            if (pos + k > len) break;
            while (scanned < pos + k) {
                if (base_code((unsigned char)seq[scanned]) < 0) bad_until = scanned;
                scanned++;
            }
            if (bad_until < pos) {
                if (n < cap) {
                    uint64_t v = 0;
                    for (int i = 0; i < k; i++) v = (v << 2) | (uint64_t)base_code((unsigned char)seq[pos + i]);
                    keys[n] = canon(v, k, min_strand);
                    offs[n] = (int32_t)pos;
                }
                n++;
                pos += 1 + cur_skip;
                emitted = 1;
                break;
            }
            cur_skip = 0;
            pos += 1;
        }
        if (!emitted) return n;
    }
}

/* ---------------------------------------------------------------- SmallFloat */
static inline int32_t f2bits(float f) { int32_t b; memcpy(&b, &f, 4); return b; }
static inline float bits2f(int32_t b) { float f; memcpy(&f, &b, 4); return f; }

/* lucene!util/SmallFloat#floatToByte315 */
ORA_API int ora_float_to_byte315(float f) {
    int32_t bits = f2bits(f);
    int32_t s = bits >> 21;
    if (s <= ((63 - 15) << 3)) return bits <= 0 ? 0 : 1;
    if (s >= ((63 - 15) << 3) + 0x100) return 0xFF;
    return (s - ((63 - 15) << 3)) & 0xFF;
}
/* lucene!util/SmallFloat#byte315ToFloat */
ORA_API float ora_byte315_to_float(int b) {
    if (b == 0) return 0.0f;
    return bits2f(((b & 0xFF) << 21) + ((63 - 15) << 24));
}
/* lucene!similarities/DefaultSimilarity#lengthNorm + #encodeNormValue */
ORA_API int ora_norm_byte(int64_t num_tokens) {
    float ln = 1.0f * (float)(1.0 / sqrt((double)num_tokens));
    return ora_float_to_byte315(ln);
}
/* lucene!similarities/DefaultSimilarity#idf */
ORA_API float ora_idf_tfidf(int64_t df, int64_t max_doc) {
    return (float)(log((double)max_doc / (double)(df + 1)) + 1.0);
}
/* lucene!similarities/BM25Similarity#idf */
ORA_API float ora_idf_bm25(int64_t df, int64_t max_doc) {
    return (float)log(1.0 + ((double)(max_doc - df) + 0.5) / ((double)df + 0.5));
}

/* ------------------------------------------------- sloppy phrase (A.7), literal */
typedef struct {
    const int32_t* pl; int n, idx, count;
    int position, offset, ord, rptGroup, rptInd;
} PP; /* lucene!search/PhrasePositions */

static inline int pp_next(PP* p) {
    if (p->count-- > 0) { p->position = p->pl[p->idx++] - p->offset; return 1; }
    return 0;
}
static inline void pp_first(PP* p) { p->count = p->n; p->idx = 0; pp_next(p); }
/* lucene!search/PhraseQueue#lessThan */
static inline int pp_less(const PP* a, const PP* b) {
    if (a->position == b->position) {
        if (a->offset == b->offset) return a->ord < b->ord;
        return a->offset < b->offset;
    }
    return a->position < b->position;
}

typedef struct { PP pp[2]; PP* q[2]; int qn; int end; int slop; int hasRpts; } Sloppy;

static inline void q_add(Sloppy* s, PP* p) { s->q[s->qn++] = p; }
static inline PP* q_top(Sloppy* s) {
    if (s->qn == 1) return s->q[0];
    return pp_less(s->q[0], s->q[1]) ? s->q[0] : s->q[1];
}
static inline PP* q_pop(Sloppy* s) {
    PP* t = q_top(s);
    if (s->qn == 2) { if (t == s->q[0]) s->q[0] = s->q[1]; }
    s->qn--;
    return t;
}
/* #advancePP */
static inline int advance_pp(Sloppy* s, PP* p) {
    if (!pp_next(p)) return 0;
    if (p->position > s->end) s->end = p->position;
    return 1;
}
/* #collide (one repeat group holding both pps) */
static inline int collide(Sloppy* s, PP* p) {
    int tp = p->position + p->offset;
    for (int i = 0; i < 2; i++) {
        PP* o = &s->pp[i];
        if (o != p && o->position + o->offset == tp) return o->rptInd;
    }
    return -1;
}
/* #lesser */
static inline PP* lesser(PP* a, PP* b) {
    if (a->position < b->position || (a->position == b->position && a->offset < b->offset)) return a;
    return b;
}
/* #advanceRpts: the heap-repair tail is a no-op for a 2-element pick-the-min queue */
static int advance_rpts(Sloppy* s, PP* p) {
    if (p->rptGroup < 0) return 1;
    int k;
    while ((k = collide(s, p)) >= 0) {
        p = lesser(p, &s->pp[k]); /* rptInd == index in pp[] (group sorted by offset) */
        if (!advance_pp(s, p)) return 0;
    }
    return 1;
}

/* lucene!search/SloppyPhraseScorer#phraseFreq for phrase (t0@0, t1@1) */
ORA_API float ora_sloppy_freq(const int32_t* p0, int n0, const int32_t* p1, int n1, int slop, int same_term) {
    Sloppy s;
    memset(&s, 0, sizeof s);
    s.slop = slop; s.hasRpts = same_term;
    s.pp[0].pl = p0; s.pp[0].n = n0; s.pp[0].offset = 0; s.pp[0].ord = 0; s.pp[0].rptGroup = -1;
    s.pp[1].pl = p1; s.pp[1].n = n1; s.pp[1].offset = 1; s.pp[1].ord = 1; s.pp[1].rptGroup = -1;
    /* #initPhrasePositions */
    s.end = INT32_MIN;
    pp_first(&s.pp[0]); pp_first(&s.pp[1]);              /* placeFirstPositions */
    if (same_term) {
        s.pp[0].rptGroup = s.pp[1].rptGroup = 0;         /* gatherRptGroups/sortRptGroups */
        s.pp[0].rptInd = 0; s.pp[1].rptInd = 1;
        if (!pp_next(&s.pp[1])) return 0.0f;             /* advanceRepeatGroups: pp[1] once */
    }
    s.qn = 0;                                            /* fillQueue */
    for (int i = 0; i < 2; i++) {
        if (s.pp[i].position > s.end) s.end = s.pp[i].position;
        q_add(&s, &s.pp[i]);
    }
    float freq = 0.0f;
    PP* pp = q_pop(&s);
    int matchLength = s.end - pp->position;
    int next = q_top(&s)->position;
    while (advance_pp(&s, pp)) {
        if (s.hasRpts && !advance_rpts(&s, pp)) break;
        if (pp->position > next) {
            if (matchLength <= slop) freq += 1.0f / (float)(matchLength + 1);
            q_add(&s, pp);
            pp = q_pop(&s);
            next = q_top(&s)->position;
            matchLength = s.end - pp->position;
        } else {
            int ml2 = s.end - pp->position;
            if (ml2 < matchLength) matchLength = ml2;
        }
    }
    if (matchLength <= slop) freq += 1.0f / (float)(matchLength + 1);
    return freq;
}

/* ----------------------------------------------------------------- the index */
typedef struct {
    int k, min_strand;
    /* staging (until finish) */
    uint64_t* t_key; uint32_t* t_doc; uint32_t* t_pos; int64_t nt, cap_t;
    /* docs */
    int32_t n_docs, cap_docs; int32_t* doc_ntok; uint8_t* doc_norm;
    /* CSR */
    int64_t n_terms, n_post;
    uint64_t* term_key; int64_t* term_off;      /* [T], [T+1] */
    int32_t* post_doc; int64_t* post_pos;       /* [P], [P+1] */
    int32_t* positions;                          /* [nt] */
    int32_t* direct;                             /* key -> term idx (or -1) when 4^k <= 2^24 */
    int finished;
    /* doc-partitioned checks: statistics of the WHOLE collection (lucene!search/IndexSearcher#termStatistics,
     * #collectionStatistics are global); unset (g_max_doc == 0) = this index is the whole collection */
    int64_t g_max_doc, g_sum_ttf; uint32_t* g_df;   /* g_df: dense [4^k] */
} ora_index;

ORA_API ora_index* ora_index_new(int k, int min_strand) {
    if (k < 1 || k > 32) return NULL;
    ora_index* ix = (ora_index*)calloc(1, sizeof *ix);
    ix->k = k; ix->min_strand = min_strand ? 1 : 0;
    return ix;
}

static void reserve_tuples(ora_index* ix, int64_t extra) {
    if (ix->nt + extra <= ix->cap_t) return;
    int64_t nc = ix->cap_t ? ix->cap_t : (1 << 20);
    while (nc < ix->nt + extra) nc = nc + nc / 2;
    ix->t_key = (uint64_t*)realloc(ix->t_key, nc * sizeof(uint64_t));
    ix->t_doc = (uint32_t*)realloc(ix->t_doc, nc * sizeof(uint32_t));
    ix->t_pos = (uint32_t*)realloc(ix->t_pos, nc * sizeof(uint32_t));
    ix->cap_t = nc;
}

static int32_t new_doc(ora_index* ix) {
    if (ix->n_docs == ix->cap_docs) {
        ix->cap_docs = ix->cap_docs ? ix->cap_docs * 2 : 64;
        ix->doc_ntok = (int32_t*)realloc(ix->doc_ntok, ix->cap_docs * sizeof(int32_t));
        ix->doc_norm = (uint8_t*)realloc(ix->doc_norm, ix->cap_docs);
    }
    return ix->n_docs++;
}

/* One Lucene document = token stream of one sequence (KmerIndexAnalyzer: skips 0),
 * positions = token ordinals (lucene!index/DefaultIndexingChain$PerField#invert). */
static void add_doc_fwd(ora_index* ix, const char* seq, int64_t len, int reverse) {
    int k = ix->k;
    int32_t d = new_doc(ix);
    int64_t nwin = len - k + 1;
    if (nwin < 0) nwin = 0;
    reserve_tuples(ix, nwin);
    uint64_t mask = k == 32 ? ~0ull : ((1ull << (2 * k)) - 1);
    uint64_t v = 0; int run = 0; uint32_t pos = 0;
    for (int64_t i = 0; i < len; i++) {
        /* reverse doc = SequenceHelper.getReverseComplement(sequence) (Indexer.java:221):
         * char i of the reverse doc is the complement of char len-1-i. */
        unsigned char c = (unsigned char)(reverse ? seq[len - 1 - i] : seq[i]);
        int b = base_code(c);
        if (b < 0) { run = 0; v = 0; continue; }
        if (reverse) b = 3 - b;
        v = ((v << 2) | (uint64_t)b) & mask;
        if (++run >= k) {
            ix->t_key[ix->nt] = canon(v, k, ix->min_strand);
            ix->t_doc[ix->nt] = (uint32_t)d;
            ix->t_pos[ix->nt] = pos++;
            ix->nt++;
        }
    }
    ix->doc_ntok[d] = (int32_t)pos;
    ix->doc_norm[d] = (uint8_t)ora_norm_byte(pos);
}

/* index/Indexer.java:178-232.  seq is the FASTA sequence as FASTAReader returns it
 * (upper-cased, trimmed).  Returns the number of documents added (1 or 2). */
ORA_API int ora_index_add_entry(ora_index* ix, const char* seq, int64_t len) {
    if (ix->finished) return -1;
    if (ix->min_strand) { add_doc_fwd(ix, seq, len, 0); return 1; }
    add_doc_fwd(ix, seq, len, 0);
    /* utils/SequenceHelper.java:33-35: chars outside 'A'..'Z' throw -> reverse doc lost */
    for (int64_t i = 0; i < len; i++)
        if (seq[i] < 'A' || seq[i] > 'Z') return 1;
    add_doc_fwd(ix, seq, len, 1);
    return 2;
}

/* stable LSD radix sort of (key, doc, pos) by key, 16-bit digits */
static void sort_tuples(ora_index* ix) {
    int64_t n = ix->nt;
    int bits = 2 * ix->k;
    uint64_t* k2 = (uint64_t*)malloc((n ? n : 1) * sizeof(uint64_t));
    uint32_t* d2 = (uint32_t*)malloc((n ? n : 1) * sizeof(uint32_t));
    uint32_t* p2 = (uint32_t*)malloc((n ? n : 1) * sizeof(uint32_t));
    int64_t* cnt = (int64_t*)malloc(65537 * sizeof(int64_t));
    for (int sh = 0; sh < bits; sh += 16) {
        memset(cnt, 0, 65537 * sizeof(int64_t));
        for (int64_t i = 0; i < n; i++) cnt[((ix->t_key[i] >> sh) & 0xFFFF) + 1]++;
        for (int i = 0; i < 65536; i++) cnt[i + 1] += cnt[i];
        for (int64_t i = 0; i < n; i++) {
            int64_t o = cnt[(ix->t_key[i] >> sh) & 0xFFFF]++;
            k2[o] = ix->t_key[i]; d2[o] = ix->t_doc[i]; p2[o] = ix->t_pos[i];
        }
        uint64_t* tk = ix->t_key; ix->t_key = k2; k2 = tk;
        uint32_t* td = ix->t_doc; ix->t_doc = d2; d2 = td;
        uint32_t* tp = ix->t_pos; ix->t_pos = p2; p2 = tp;
    }
    free(k2); free(d2); free(p2); free(cnt);
}

ORA_API int ora_index_finish(ora_index* ix) {
    if (ix->finished) return 0;
    sort_tuples(ix);
    int64_t n = ix->nt, T = 0, P = 0;
    for (int64_t i = 0; i < n; i++) {
        int nk = (i == 0 || ix->t_key[i] != ix->t_key[i - 1]);
        if (nk) T++;
        if (nk || ix->t_doc[i] != ix->t_doc[i - 1]) P++;
    }
    ix->n_terms = T; ix->n_post = P;
    ix->term_key = (uint64_t*)malloc((T + 1) * sizeof(uint64_t));
    ix->term_off = (int64_t*)malloc((T + 1) * sizeof(int64_t));
    ix->post_doc = (int32_t*)malloc((P + 1) * sizeof(int32_t));
    ix->post_pos = (int64_t*)malloc((P + 1) * sizeof(int64_t));
    ix->positions = (int32_t*)malloc((n + 1) * sizeof(int32_t));
    int64_t t = -1, p = -1;
    for (int64_t i = 0; i < n; i++) {
        int nk = (i == 0 || ix->t_key[i] != ix->t_key[i - 1]);
        if (nk) { t++; ix->term_key[t] = ix->t_key[i]; ix->term_off[t] = p + 1; }
        if (nk || ix->t_doc[i] != ix->t_doc[i - 1]) { p++; ix->post_doc[p] = (int32_t)ix->t_doc[i]; ix->post_pos[p] = i; }
        ix->positions[i] = (int32_t)ix->t_pos[i];
    }
    ix->term_off[T] = P; ix->post_pos[P] = n;
    free(ix->t_key); free(ix->t_doc); free(ix->t_pos);
    ix->t_key = NULL; ix->t_doc = NULL; ix->t_pos = NULL;
    if (2 * ix->k <= 24) {
        int64_t sz = 1ll << (2 * ix->k);
        ix->direct = (int32_t*)malloc(sz * sizeof(int32_t));
        for (int64_t i = 0; i < sz; i++) ix->direct[i] = -1;
        for (int64_t i = 0; i < T; i++) ix->direct[ix->term_key[i]] = (int32_t)i;
    }
    ix->finished = 1;
    return 0;
}

ORA_API void ora_index_free(ora_index* ix) {
    if (!ix) return;
    free(ix->t_key); free(ix->t_doc); free(ix->t_pos); free(ix->doc_ntok); free(ix->doc_norm);
    free(ix->term_key); free(ix->term_off); free(ix->post_doc); free(ix->post_pos);
    free(ix->positions); free(ix->direct); free(ix->g_df); free(ix);
}

ORA_API int ora_index_stats(const ora_index* ix, int64_t* n_docs, int64_t* n_terms, int64_t* n_postings, int64_t* n_positions) {
    if (!ix->finished) return -1;
    *n_docs = ix->n_docs; *n_terms = ix->n_terms; *n_postings = ix->n_post; *n_positions = ix->nt;
    return 0;
}

ORA_API int ora_doc_info(const ora_index* ix, int32_t doc, int32_t* ntok, int32_t* norm) {
    if (doc < 0 || doc >= ix->n_docs) return -1;
    *ntok = ix->doc_ntok[doc]; *norm = ix->doc_norm[doc];
    return 0;
}

static int64_t find_term(const ora_index* ix, uint64_t key) {
    if (ix->direct) {
        if (2 * ix->k < 64 && key >> (2 * ix->k)) return -1;
        return ix->direct[key];
    }
    int64_t lo = 0, hi = ix->n_terms - 1;
    while (lo <= hi) {
        int64_t m = (lo + hi) >> 1;
        if (ix->term_key[m] == key) return m;
        if (ix->term_key[m] < key) lo = m + 1; else hi = m - 1;
    }
    return -1;
}

/* postings of one term: df, docs[df], pos_off[df+1] (into positions[]) */
ORA_API int ora_term_postings(const ora_index* ix, uint64_t key, int64_t* df, const int32_t** docs,
                              const int64_t** pos_off, const int32_t** positions) {
    int64_t t = find_term(ix, key);
    if (t < 0) { *df = 0; *docs = NULL; *pos_off = NULL; *positions = ix->positions; return 0; }
    int64_t a = ix->term_off[t], b = ix->term_off[t + 1];
    *df = b - a; *docs = ix->post_doc + a; *pos_off = ix->post_pos + a; *positions = ix->positions;
    return 0;
}

/* i-th term in key order (for exhaustive postings comparison) */
ORA_API int ora_term_at(const ora_index* ix, int64_t i, uint64_t* key) {
    if (i < 0 || i >= ix->n_terms) return -1;
    *key = ix->term_key[i];
    return 0;
}

/* -------------------------------------------------------------- classification */
typedef struct {
    int32_t kmer_skips;
    int32_t query_algorithm;    /* 0 NAIVE_KMER 1 CHAIN_PROXIMITY 2 PAIRED_PROXIMITY */
    int32_t scoring_algorithm;  /* 0 tfidf 1 bm25 */
    int32_t threads;
    double query_term_min_should_match;
} ora_query_conf;

typedef struct { int kind; uint64_t t0, t1; int slop; } Clause; /* kind 0 Term, 1 Phrase */

typedef struct {
    double* acc; int32_t* cnt; int32_t* touched; int32_t n_touched;
    uint64_t* keys; int32_t* offs; Clause* cl; float* idf; float* val; int64_t tok_cap;
} Scratch;

static void scratch_init(Scratch* s, int32_t n_docs) {
    memset(s, 0, sizeof *s);
    s->acc = (double*)calloc(n_docs + 1, sizeof(double));
    s->cnt = (int32_t*)calloc(n_docs + 1, sizeof(int32_t));
    s->touched = (int32_t*)malloc((n_docs + 1) * sizeof(int32_t));
    s->cl = (Clause*)malloc((MAX_CLAUSES + 2) * sizeof(Clause));
    s->idf = (float*)malloc((MAX_CLAUSES + 2) * sizeof(float));
    s->val = (float*)malloc((MAX_CLAUSES + 2) * sizeof(float));
}
static void scratch_free(Scratch* s) {
    free(s->acc); free(s->cnt); free(s->touched); free(s->keys); free(s->offs); free(s->cl); free(s->idf); free(s->val);
}

static inline void add_score(Scratch* s, int32_t d, float sc) {
    if (s->cnt[d] == 0) s->touched[s->n_touched++] = d;
    s->acc[d] += (double)sc;  /* lucene!search/BooleanScorer$OrCollector#collect: double sum */
    s->cnt[d] += 1;
}

typedef struct {
    int sim; float cache[256]; float norm_table[256];
} SimCtx;

/* lucene!similarities/TFIDFSimilarity$TFIDFSimScorer#score / BM25Similarity$BM25DocScorer#score */
static inline float doc_score(const SimCtx* c, float freq, float value, int normb) {
    if (c->sim == SIM_TFIDF) {
        float tf = (float)sqrt((double)freq);
        float raw = tf * value;
        return raw * c->norm_table[normb];
    }
    float wv = value * (1.2f + 1.0f);
    return (wv * freq) / (freq + c->cache[normb]);
}

static int64_t df_of(const ora_index* ix, uint64_t key) {
    if (ix->g_df) return (int64_t)ix->g_df[key];
    int64_t t = find_term(ix, key);
    return t < 0 ? 0 : ix->term_off[t + 1] - ix->term_off[t];
}

/* classify one read.  Returns status 0 ok / 1 dropped.
 * top_doc/top_score: up to TOPN (score desc, doc asc); n_hits = entries equal to max. */
static int classify_one(const ora_index* ix, const ora_query_conf* qc, const SimCtx* sc, Scratch* s,
                        const char* seq, int32_t len, int32_t* n_clauses_out, int32_t* msm_out,
                        int32_t* n_top, int32_t* n_hits, int32_t* top_doc, float* top_score) {
    *n_top = 0; *n_hits = 0; *n_clauses_out = 0; *msm_out = 0;
    int k = ix->k;
    if (len + 1 > s->tok_cap) {
        s->tok_cap = len + 1;
        s->keys = (uint64_t*)realloc(s->keys, s->tok_cap * sizeof(uint64_t));
        s->offs = (int32_t*)realloc(s->offs, s->tok_cap * sizeof(int32_t));
    }
    int64_t n = ora_tokenize(seq, len, k, qc->kmer_skips, ix->min_strand, s->keys, s->offs, s->tok_cap);
    if (n <= 1) return 1;                      /* classify/Classifier.java:223-227 */
    /* clause construction (classify/Classifier.java:113-200) */
    int64_t nc = 0;
    if (qc->query_algorithm == ALG_NAIVE) {
        if (n > MAX_CLAUSES) return 1;
        for (int64_t i = 0; i < n; i++) { s->cl[nc].kind = 0; s->cl[nc].t0 = s->keys[i]; nc++; }
    } else if (qc->query_algorithm == ALG_CHAIN) {
        if (n - 1 > MAX_CLAUSES) return 1;
        for (int64_t i = 1; i < n; i++) {
            int d = s->offs[i] - s->offs[i - 1];
            if (d > 0) { s->cl[nc].kind = 1; s->cl[nc].t0 = s->keys[i - 1]; s->cl[nc].t1 = s->keys[i]; s->cl[nc].slop = d + 1; nc++; }
        }
    } else {
        if ((n + 1) / 2 > MAX_CLAUSES) return 1;
        for (int64_t j = 0; j + 1 < n; j += 2) {
            int d = s->offs[j + 1] - s->offs[j];
            if (d > 0) { s->cl[nc].kind = 1; s->cl[nc].t0 = s->keys[j]; s->cl[nc].t1 = s->keys[j + 1]; s->cl[nc].slop = d + 1; nc++; }
        }
        if (n & 1) { s->cl[nc].kind = 0; s->cl[nc].t0 = s->keys[n - 1]; nc++; }
    }
    int32_t msm = (int32_t)(qc->query_term_min_should_match * (double)nc); /* Classifier.java:257 */
    *n_clauses_out = (int32_t)nc; *msm_out = msm;
    int64_t max_doc = ix->g_max_doc ? ix->g_max_doc : ix->n_docs;
    /* weights (A.5) */
    for (int64_t c = 0; c < nc; c++) {
        float a = sc->sim == SIM_TFIDF ? ora_idf_tfidf(df_of(ix, s->cl[c].t0), max_doc)
                                       : ora_idf_bm25(df_of(ix, s->cl[c].t0), max_doc);
        if (s->cl[c].kind == 1) {
            float b = sc->sim == SIM_TFIDF ? ora_idf_tfidf(df_of(ix, s->cl[c].t1), max_doc)
                                           : ora_idf_bm25(df_of(ix, s->cl[c].t1), max_doc);
            float acc = 0.0f; acc += a; acc += b; a = acc;  /* TFIDFSimilarity#idfExplain(stats, TermStatistics[]) */
        }
        s->idf[c] = a;
    }
    if (sc->sim == SIM_TFIDF) {
        float sum_sq = 0.0f;
        for (int64_t c = 0; c < nc; c++) { float qw = s->idf[c] * 1.0f; sum_sq += qw * qw; }
        sum_sq *= 1.0f * 1.0f;                                  /* BooleanWeight#getValueForNormalization */
        float qn = (float)(1.0 / sqrt((double)sum_sq));         /* DefaultSimilarity#queryNorm */
        if (isinf(qn) || isnan(qn)) qn = 1.0f;                  /* IndexSearcher#createNormalizedWeight */
        for (int64_t c = 0; c < nc; c++) {
            float qw = s->idf[c] * 1.0f;
            qw = qw * (qn * 1.0f);                              /* IDFStats#normalize */
            s->val[c] = qw * s->idf[c];
        }
    } else {
        for (int64_t c = 0; c < nc; c++) s->val[c] = (s->idf[c] * 1.0f) * 1.0f; /* BM25Stats#normalize */
    }
    /* term-at-a-time scoring into per-doc accumulators */
    s->n_touched = 0;
    for (int64_t c = 0; c < nc; c++) {
        const Clause* cl = &s->cl[c];
        int64_t t0 = find_term(ix, cl->t0);
        if (t0 < 0) continue;
        int64_t a0 = ix->term_off[t0], b0 = ix->term_off[t0 + 1];
        if (cl->kind == 0) {
            for (int64_t p = a0; p < b0; p++) {
                int32_t d = ix->post_doc[p];
                float freq = (float)(ix->post_pos[p + 1] - ix->post_pos[p]);  /* TermScorer#score */
                add_score(s, d, doc_score(sc, freq, s->val[c], ix->doc_norm[d]));
            }
        } else {
            int64_t t1 = find_term(ix, cl->t1);
            if (t1 < 0) continue;
            int64_t a1 = ix->term_off[t1], b1 = ix->term_off[t1 + 1];
            int same = cl->t0 == cl->t1;
            int64_t p = a0, q = a1;
            while (p < b0 && q < b1) {                          /* ConjunctionDISI */
                int32_t d0 = ix->post_doc[p], d1 = ix->post_doc[q];
                if (d0 < d1) p++;
                else if (d1 < d0) q++;
                else {
                    float fr = ora_sloppy_freq(ix->positions + ix->post_pos[p], (int)(ix->post_pos[p + 1] - ix->post_pos[p]),
                                               ix->positions + ix->post_pos[q], (int)(ix->post_pos[q + 1] - ix->post_pos[q]),
                                               cl->slop, same);
                    if (fr != 0.0f) add_score(s, d0, doc_score(sc, fr, s->val[c], ix->doc_norm[d0]));
                    p++; q++;
                }
            }
        }
    }
    /* collect: msm, coord, top-N (lucene!search/TopScoreDocCollector: score desc, doc asc) */
    int32_t need = msm > 1 ? msm : 1;
    int nt = 0;
    for (int32_t i = 0; i < s->n_touched; i++) {
        int32_t d = s->touched[i];
        int32_t m = s->cnt[d];
        double sum = s->acc[d];
        s->cnt[d] = 0; s->acc[d] = 0.0;
        if (m < need) continue;
        float coord;
        if (sc->sim == SIM_BM25) coord = 1.0f;
        else coord = nc == 1 ? 1.0f : (float)m / (float)nc;       /* BooleanWeight#coord, DefaultSimilarity#coord */
        float score = (float)sum * coord;
        /* insert into sorted top list */
        int pos = nt;
        while (pos > 0 && (top_score[pos - 1] < score || (top_score[pos - 1] == score && top_doc[pos - 1] > d))) pos--;
        if (pos >= TOPN) continue;
        int last = nt < TOPN ? nt : TOPN - 1;
        for (int j = last; j > pos; j--) { top_score[j] = top_score[j - 1]; top_doc[j] = top_doc[j - 1]; }
        top_score[pos] = score; top_doc[pos] = d;
        if (nt < TOPN) nt++;
    }
    *n_top = nt;
    int nh = 0;
    for (int i = 0; i < nt; i++)
        if ((double)top_score[0] - (double)top_score[i] == 0) nh++;   /* Classifier.java:358-366 */
    *n_hits = nh;
    return 0;
}

static void sim_init(SimCtx* c, const ora_index* ix, int sim) {
    c->sim = sim;
    for (int b = 0; b < 256; b++) c->norm_table[b] = ora_byte315_to_float(b);
    if (sim == SIM_BM25) {
        int64_t stt = 0, nd = ix->n_docs;
        for (int32_t d = 0; d < ix->n_docs; d++) stt += ix->doc_ntok[d];
        if (ix->g_max_doc) { stt = ix->g_sum_ttf; nd = ix->g_max_doc; }
        float avgdl = stt <= 0 ? 1.0f : (float)((double)stt / (double)nd);           /* BM25Similarity#avgFieldLength */
        const float k1 = 1.2f, b = 0.75f;
        for (int x = 0; x < 256; x++) {
            float f = c->norm_table[x];
            float ln = 1.0f / (f * f);                                               /* BM25Similarity.<clinit> */
            c->cache[x] = k1 * ((1.0f - b) + (b * ln) / avgdl);                      /* #computeWeight */
        }
    }
}

/* Batch classify.  Outputs per read i: status[i] (0 ok, 1 dropped), label[i]
 * (0 UNKNOWN, 1 VAGUE, 2 CLASSIFIED; score tier only, no taxonomy), n_hits[i],
 * n_top[i], top_doc[i*10..], top_score[i*10..], n_clauses[i], msm[i]. */
ORA_API int ora_classify_batch(const ora_index* ix, const ora_query_conf* qc, int32_t n,
                               const char* const* seqs, const int32_t* lens,
                               int32_t* status, int32_t* label, int32_t* n_hits, int32_t* n_top,
                               int32_t* top_doc, float* top_score, int32_t* n_clauses, int32_t* msm) {
    if (!ix->finished) return -1;
    SimCtx sc;
    sim_init(&sc, ix, qc->scoring_algorithm);
    int threads = qc->threads > 0 ? qc->threads : 1;
#ifdef _OPENMP
#pragma omp parallel num_threads(threads)
#endif
    {
        Scratch s;
        scratch_init(&s, ix->n_docs);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 16)
#endif
        for (int32_t i = 0; i < n; i++) {
            status[i] = classify_one(ix, qc, &sc, &s, seqs[i], lens[i], &n_clauses[i], &msm[i],
                                     &n_top[i], &n_hits[i], top_doc + (int64_t)i * TOPN, top_score + (int64_t)i * TOPN);
            label[i] = n_hits[i] == 0 ? 0 : (n_hits[i] == 1 ? 2 : 1);
        }
        scratch_free(&s);
    }
    (void)threads;
    return 0;
}

/* dense per-term document frequencies of this index (k <= 12): out[4^k] */
ORA_API int ora_index_dense_df(const ora_index* ix, uint32_t* out) {
    if (!ix->finished || ix->k > 12) return -1;
    memset(out, 0, ((size_t)1 << (2 * ix->k)) * sizeof(uint32_t));
    for (int64_t t = 0; t < ix->n_terms; t++) out[ix->term_key[t]] = (uint32_t)(ix->term_off[t + 1] - ix->term_off[t]);
    return 0;
}

/* score against collection-wide statistics (doc-partitioned mode): max_doc, sum of doc lengths, dense df[4^k] (copied) */
ORA_API int ora_index_set_global(ora_index* ix, int64_t max_doc, int64_t sum_ttf, const uint32_t* dense_df) {
    if (!ix->finished || ix->k > 12) return -1;
    size_t n = (size_t)1 << (2 * ix->k);
    free(ix->g_df);
    ix->g_df = (uint32_t*)malloc(n * sizeof(uint32_t));
    memcpy(ix->g_df, dense_df, n * sizeof(uint32_t));
    ix->g_max_doc = max_doc; ix->g_sum_ttf = sum_ttf;
    return 0;
}

ORA_API int64_t ora_index_sum_ttf(const ora_index* ix) {
    int64_t s = 0;
    for (int32_t d = 0; d < ix->n_docs; d++) s += ix->doc_ntok[d];
    return s;
}

ORA_API int ora_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
