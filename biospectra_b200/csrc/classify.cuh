// classify.cuh -- query-side kernels (K5-K8 of SURVEY.md 2.1).
//
//   K5 tokenize_kernel / weights_kernel   read -> tokens -> implicit clauses + Lucene weights
//   K6/K7 score_positional_kernel         general path: term-major postings + positions,
//                                         TermQuery and sloppy PhraseQuery scoring
//   K6' dense4.cuh                        dense path (small k): 4-bit frequency tables, packed-SIMD
//                                         filter + exact rescoring, and the all-exact scorer
//   K8 block top-10 (fused into both scorers) + merge_kernel (labels)
//
// Float semantics follow Lucene 5.3.1 operation by operation: explicit
// __fmul_rn/__fadd_rn/__fdiv_rn (no FMA contraction), double accumulation of the per-clause
// float scores, one (float) cast, then coord.
#pragma once
#include "common.cuh"
#include "index_build.cuh"

#define ALG_NAIVE 0
#define ALG_CHAIN 1
#define ALG_PAIRED 2
#define SIM_TFIDF 0
#define SIM_BM25 1
#define TOPN 10
#define MAX_CLAUSES 10000        // Classifier.java:110 BooleanQuery.setMaxClauseCount(10000)

#define ROUTE_NONE 0             // dropped / error: no scoring
#define ROUTE_DENSE 1            // exact_score_kernel over the 4-bit tables
#define ROUTE_POSITIONAL 2
#define ROUTE_FILTER 3           // filter_score_kernel (<= 255 clauses; BM25: <= 32 norm classes in the handle)

#define ST_OK 0
#define ST_DROPPED 1
#define ST_ERROR 2

// ------------------------------------------------------- sloppy phrase (A.7)
// lucene!search/SloppyPhraseScorer#phraseFreq for the phrase (t0@0, t1@1), t0 != t1.
// A = positions of t0, B = positions of t1 (ascending, non-empty).  With two phrase
// positions the priority queue degenerates to "the other one": after pp passes `next`
// the other pp is strictly smaller, so pop() always returns it.
__device__ __forceinline__ float sloppy_freq_distinct(const u32* __restrict__ A, int nA, const u32* __restrict__ B, int nB,
                                                      int slop) {
    int pa = (int)A[0], pb = (int)B[0] - 1;
    int end = pa > pb ? pa : pb;
    // PhraseQueue#lessThan: smaller position first, ties -> smaller offset (t0)
    bool cur_is_a = pa <= pb;
    const u32* cl = cur_is_a ? A : B; int cn = cur_is_a ? nA : nB; int co = cur_is_a ? 0 : 1; int cp = cur_is_a ? pa : pb;
    const u32* ol = cur_is_a ? B : A; int on = cur_is_a ? nB : nA; int oo = cur_is_a ? 1 : 0; int op = cur_is_a ? pb : pa;
    int ci = 1, oi = 1;
    int ml = end - cp;
    int next = op;
    float freq = 0.0f;
    while (ci < cn) {                                   // advancePP
        cp = (int)cl[ci++] - co;
        if (cp > end) end = cp;
        if (cp > next) {
            if (ml <= slop) freq = __fadd_rn(freq, __fdiv_rn(1.0f, (float)(ml + 1)));
            // pq.add(pp); pp = pq.pop()  -> swap roles
            const u32* tl = cl; cl = ol; ol = tl;
            int t;
            t = cn; cn = on; on = t;
            t = co; co = oo; oo = t;
            t = ci; ci = oi; oi = t;
            t = cp; cp = op; op = t;
            next = op;
            ml = end - cp;
        } else {
            int ml2 = end - cp;
            if (ml2 < ml) ml = ml2;
        }
    }
    if (ml <= slop) freq = __fadd_rn(freq, __fdiv_rn(1.0f, (float)(ml + 1)));
    return freq;
}

// Same-term phrase (t0 == t1): the repeat-group path, restated literally
// (#initComplex/#advanceRepeatGroups/#advanceRpts/#collide/#lesser).
struct PPd { const u32* pl; int n, idx, count, position, offset; };
__device__ __forceinline__ bool ppd_next(PPd& p) {
    if (p.count-- > 0) { p.position = (int)p.pl[p.idx++] - p.offset; return true; }
    return false;
}
__device__ __forceinline__ bool ppd_less(const PPd& a, const PPd& b) {   // PhraseQueue#lessThan (ord == offset here)
    if (a.position == b.position) return a.offset < b.offset;
    return a.position < b.position;
}
__device__ __noinline__ float sloppy_freq_same(const u32* __restrict__ P, int n, int slop) {
    PPd pp[2];
    pp[0].pl = P; pp[0].n = n; pp[0].offset = 0;
    pp[1].pl = P; pp[1].n = n; pp[1].offset = 1;
    int end = INT_MIN;
    for (int i = 0; i < 2; i++) { pp[i].count = n; pp[i].idx = 0; ppd_next(pp[i]); }   // placeFirstPositions
    if (!ppd_next(pp[1])) return 0.0f;                                               // advanceRepeatGroups
    for (int i = 0; i < 2; i++) if (pp[i].position > end) end = pp[i].position;       // fillQueue
    int cur = ppd_less(pp[0], pp[1]) ? 0 : 1;                                         // pq.pop()
    int ml = end - pp[cur].position;
    int next = pp[1 - cur].position;                                                  // pq.top()
    float freq = 0.0f;
    while (true) {
        // advancePP(pp)
        if (!ppd_next(pp[cur])) break;
        if (pp[cur].position > end) end = pp[cur].position;
        // advanceRpts(pp): `p` is the local variable that #lesser rebinds
        {
            int p = cur;
            bool exhausted = false;
            while (true) {
                int o = 1 - p;                                                        // #collide
                if (pp[o].position + pp[o].offset != pp[p].position + pp[p].offset) break;
                // #lesser(pp, rg[k])
                int l = (pp[p].position < pp[o].position ||
                         (pp[p].position == pp[o].position && pp[p].offset < pp[o].offset)) ? p : o;
                p = l;
                if (!ppd_next(pp[p])) { exhausted = true; break; }
                if (pp[p].position > end) end = pp[p].position;
            }
            if (exhausted) break;
        }
        if (pp[cur].position > next) {
            if (ml <= slop) freq = __fadd_rn(freq, __fdiv_rn(1.0f, (float)(ml + 1)));
            // pq.add(pp); pp = pq.pop(): the queue holds both -> take the lesser
            cur = ppd_less(pp[0], pp[1]) ? 0 : 1;
            next = pp[1 - cur].position;
            ml = end - pp[cur].position;
        } else {
            int ml2 = end - pp[cur].position;
            if (ml2 < ml) ml = ml2;
        }
    }
    if (ml <= slop) freq = __fadd_rn(freq, __fdiv_rn(1.0f, (float)(ml + 1)));
    return freq;
}

// lower_bound of doc in post_doc[a,b); returns index or b
__device__ __forceinline__ u32 find_posting(const u32* __restrict__ post_doc, u32 a, u32 b, u32 doc, u32 guess) {
    if (guess >= a && guess < b && post_doc[guess] == doc) return guess;
    u32 lo = a, hi = b;
    while (lo < hi) {
        u32 mid = (lo + hi) >> 1;
        if (post_doc[mid] < doc) lo = mid + 1; else hi = mid;
    }
    return (lo < b && post_doc[lo] == doc) ? lo : b;
}

// phrase frequency of (t0, t1, slop) in the doc of posting i0 of t0; 0 if t1 is absent
__device__ __forceinline__ float phrase_freq_at(const SegDev& s, u32 i0, u32 a1, u32 b1, u32 guess, int slop, bool same) {
    const u32 p0 = s.post_pos[i0], n0 = s.post_pos[i0 + 1] - p0;
    if (same) return sloppy_freq_same(s.positions + p0, (int)n0, slop);
    u32 i1 = find_posting(s.post_doc, a1, b1, s.post_doc[i0], guess);
    if (i1 == b1) return 0.0f;
    const u32 p1 = s.post_pos[i1], n1 = s.post_pos[i1 + 1] - p1;
    return sloppy_freq_distinct(s.positions + p0, (int)n0, s.positions + p1, (int)n1, slop);
}

// ------------------------------------------------------------------------ K5
struct QueryParams {
    int k, skips, min_strand, alg, sim;
    double msm_ratio;
    i64 max_doc;            // global maxDoc (all partitions)
    int dense_ok;           // dense tables resident
    int positional_ok;      // positional segments resident
    int filter_ok;          // fast filter path usable
    int filter_nc_max;      // most clauses of a read the filter takes (255 count lanes; BM25: the per-class LUT arrays must fit)
    u64 pair_row0;          // first pair row of the dense table (= 4^k)
    int gap_ok;             // gap clauses can be evaluated sparsely (dense tables AND positional postings resident)
    u32 temp_row0;          // row id of gap clause 0 of the sub-batch
    u32 gap_cap;            // capacity of the sub-batch's gap clause list
};

// Gap clauses: a phrase clause whose two k-mers are not adjacent in the read (an N in between, or kmer_skips > 0) is
// not a function of a (k+1)-mer, so no table row exists for it.  Its matches are sparse (two independent k-mers must
// fall within the slop of each other: ~1 doc per clause at C2), so a pre-pass evaluates it against the positional
// postings and stores the matches as the exception list of a virtual all-zero row (dense4.cuh).
#define GAP_CAP_DEFAULT 4096    // gap clauses per sub-batch (reads with an N); BSP_GAP_CAP overrides
#define GAP_LIST 64             // matches kept per gap clause (more: the read falls back to the positional scorer)
struct GapClause { u64 k0, k1; i32 slop; i32 read; };

__device__ __forceinline__ int base_code_dev(u8 c) {
    return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : -1;
}

// Warp per read: lucene/KmerSequenceTokenizer.java:115-154 + SequenceCompressFilter.
// tok_off[r] = capacity offset of read r's token slots (max(len-k+1,0) each).
// `codes` (optional): 64 bytes of shared memory private to the warp.  With it the kmer_skips = 0 path stages the 2-bit codes
// of the 32 + k - 1 bases a round of 32 windows covers (two coalesced byte loads per lane) and builds the windows from
// shared memory instead of k byte loads from global memory per window; `df_dense` (optional) looks each token's document
// frequency up where the token is emitted (tok_df), which saves token_df_kernel's pass over the keys.
__device__ __forceinline__ void tokenize_read(const u8* __restrict__ seq, const i64* __restrict__ seq_off, i32 r, int lane,
                                              const i64* __restrict__ tok_off, int k, int skips, int min_strand,
                                              u64* __restrict__ tok_key, u32* __restrict__ tok_pos, i32* __restrict__ n_tok,
                                              u8* codes, const u32* __restrict__ df_dense, u32* __restrict__ tok_df) {
    const u8* s = seq + seq_off[r];
    const i64 len = seq_off[r + 1] - seq_off[r];
    const i64 nwin = len - k + 1;
    u64* keys = tok_key + tok_off[r];
    u32* offs = tok_pos + tok_off[r];
    i32 cnt = 0;
    if (nwin > 0) {
        if (skips == 0 && codes != nullptr) {
            u32* dfs = tok_df + tok_off[r];
            for (i64 w0 = 0; w0 < nwin; w0 += 32) {
                __syncwarp();
                codes[lane] = w0 + lane < len ? (u8)(base_code_dev(s[w0 + lane]) & 7) : (u8)4;           // invalid base: bit 2 set
                if (lane < k - 1) codes[32 + lane] = w0 + 32 + lane < len ? (u8)(base_code_dev(s[w0 + 32 + lane]) & 7) : (u8)4;
                __syncwarp();
                const i64 w = w0 + lane;
                u64 v = 0;
                u32 bad = 0;
                for (int i = 0; i < k; i++) { const u32 b = codes[lane + i]; bad |= b; v = (v << 2) | (u64)(b & 3u); }
                const bool ok = w < nwin && !(bad & 4u);
                const u32 bal = __ballot_sync(0xFFFFFFFFu, ok);
                if (ok) {
                    const i32 o = cnt + __popc(bal & ((1u << lane) - 1u));
                    const u64 key = canon_kmer(v, k, min_strand);
                    keys[o] = key;
                    offs[o] = (u32)w;
                    if (df_dense) dfs[o] = df_dense[key];
                }
                cnt += __popc(bal);
            }
        } else if (skips == 0) {
            // every valid window is a token: ordered compaction by ballot
            for (i64 w0 = 0; w0 < nwin; w0 += 32) {
                i64 w = w0 + lane;
                bool ok = w < nwin;
                u64 v = 0;
                if (ok) {
                    for (int i = 0; i < k; i++) {
                        int b = base_code_dev(s[w + i]);
                        if (b < 0) { ok = false; break; }
                        v = (v << 2) | (u64)b;
                    }
                }
                u32 bal = __ballot_sync(0xFFFFFFFFu, ok);
                if (ok) {
                    i32 o = cnt + __popc(bal & ((1u << lane) - 1u));
                    keys[o] = canon_kmer(v, k, min_strand);
                    offs[o] = (u32)w;
                }
                cnt += __popc(bal);
            }
        } else {
            // stride rule: 1+skips after an emit, 1 while dropping and for the emit that ends a drop run
            if (lane == 0) {
                i64 pos = 0, scanned = 0, bad = -1;
                while (true) {
                    int cur_skip = skips;
                    bool emitted = false;
                    while (pos + k <= len) {
                        while (scanned < pos + k) { if (base_code_dev(s[scanned]) < 0) bad = scanned; scanned++; }
                        if (bad < pos) { offs[cnt++] = (u32)pos; pos += 1 + cur_skip; emitted = true; break; }
                        cur_skip = 0;
                        pos += 1;
                    }
                    if (!emitted) break;
                }
            }
            cnt = __shfl_sync(0xFFFFFFFFu, cnt, 0);
            __syncwarp();
            for (i32 j = lane; j < cnt; j += 32) {
                u32 w = offs[j];
                u64 v = 0;
                for (int i = 0; i < k; i++) v = (v << 2) | (u64)base_code_dev(s[w + i]);
                keys[j] = canon_kmer(v, k, min_strand);
            }
        }
    }
    if (lane == 0) n_tok[r] = cnt;
}
__global__ void __launch_bounds__(256) tokenize_kernel(const u8* __restrict__ seq, const i64* __restrict__ seq_off, i32 n_reads,
                                                        const i64* __restrict__ tok_off, int k, int skips, int min_strand,
                                                        u64* __restrict__ tok_key, u32* __restrict__ tok_pos,
                                                        i32* __restrict__ n_tok) {
    const int lane = threadIdx.x & 31;
    const i32 r = (i32)((blockIdx.x * (u64)blockDim.x + threadIdx.x) >> 5);
    if (r >= n_reads) return;
    tokenize_read(seq, seq_off, r, lane, tok_off, k, skips, min_strand, tok_key, tok_pos, n_tok, nullptr, nullptr, nullptr);
}

// number of clauses for n tokens (Classifier.java:113-200; offsetDiff > 0 always holds)
__host__ __device__ __forceinline__ i32 clause_count(int alg, i32 n) {
    if (n <= 1) return 0;
    if (alg == ALG_NAIVE) return n;
    if (alg == ALG_CHAIN) return n - 1;
    return (n + 1) / 2;
}
// clause c -> token indices (j0, j1); j1 < 0 for a Term clause
__device__ __forceinline__ void clause_tokens(int alg, i32 n, i32 c, i32& j0, i32& j1) {
    if (alg == ALG_NAIVE) { j0 = c; j1 = -1; }
    else if (alg == ALG_CHAIN) { j0 = c; j1 = c + 1; }
    else { j0 = 2 * c; j1 = (2 * c + 1 < n) ? 2 * c + 1 : -1; }
}

// local df of every token: sum over this handle's segments (or the dense df array)
__global__ void __launch_bounds__(256) token_df_kernel(const u64* __restrict__ tok_key, const i64* __restrict__ tok_off,
                                                        const i32* __restrict__ n_tok, i32 n_reads,
                                                        const u32* __restrict__ df_dense, const SegDev* __restrict__ segs,
                                                        int n_segs, u32* __restrict__ tok_df) {
    const int lane = threadIdx.x & 31;
    const i32 r = (i32)((blockIdx.x * (u64)blockDim.x + threadIdx.x) >> 5);
    if (r >= n_reads) return;
    const i64 base = tok_off[r];
    const i32 n = n_tok[r];
    for (i32 j = lane; j < n; j += 32) {
        u64 key = tok_key[base + j];
        u32 df = 0;
        if (df_dense) df = df_dense[key];
        else {
            for (int sidx = 0; sidx < n_segs; sidx++) {
                const SegDev& s = segs[sidx];
                u32 t = seg_find_term(s, key);
                if (t != NO_TERM) df += s.term_off[t + 1] - s.term_off[t];
            }
        }
        tok_df[base + j] = df;
    }
}

// Warp per read: msm, clause weights (A.5) and routing.  Lanes work on clauses in parallel; the only order-dependent
// step -- BooleanWeight#getValueForNormalization's float sum over the clauses -- is replayed in clause order.
//   clause_val[tok_off[r] + c] = value_c ; clause_row[...] = dense-table row (dense route)
__device__ __forceinline__ void weights_read(const u8* __restrict__ seq, const i64* __restrict__ seq_off, i32 r, int lane,
                                             const i64* __restrict__ tok_off, i32 n,
                                             // (no __restrict__: front_kernel wrote these arrays a moment ago -- they must not
                                             //  be read through the non-coherent path)
                                             const u64* tok_key, const u32* tok_pos, const u32* tok_df,
                                             const float* __restrict__ idf_by_df, const QueryParams& qp,
                                             float* __restrict__ clause_val, u32* __restrict__ clause_row,
                                             i32* __restrict__ n_clauses, i32* __restrict__ msm_out,
                                             i32* __restrict__ route, i32* __restrict__ status,
                                             float2* __restrict__ vrange /* per read: (max, min) clause value */,
                                             GapClause* __restrict__ gaps, u32* __restrict__ gap_count,
                                             i32* __restrict__ read_ngap /* per read: # gap clauses registered */) {
    const i64 base = tok_off[r];
    const i32 nc = clause_count(qp.alg, n);
    if (n <= 1 || nc > MAX_CLAUSES) {               // Classifier.java:223-227 / TooManyClauses
        if (lane == 0) { n_clauses[r] = nc; msm_out[r] = 0; route[r] = ROUTE_NONE; status[r] = ST_DROPPED; }
        return;
    }
    bool dense = qp.dense_ok != 0;
    float sum_sq = 0.0f;
    i32 n_gap = 0;
    for (i32 c0 = 0; c0 < nc; c0 += 32) {
        const i32 c = c0 + lane;
        float sq = 0.0f;
        bool bad_slop = false, is_gap = false;
        if (c < nc) {
            i32 j0, j1;
            clause_tokens(qp.alg, n, c, j0, j1);
            float idf = idf_by_df[tok_df[base + j0]];
            if (j1 >= 0) {
                // TFIDFSimilarity#idfExplain(stats, TermStatistics[]): float sum, t0 first
                idf = __fadd_rn(__fadd_rn(0.0f, idf), idf_by_df[tok_df[base + j1]]);
                const u32 delta = tok_pos[base + j1] - tok_pos[base + j0];
                if (delta != 1u) {                                                  // slop != 2: not a dense-table row
                    bad_slop = true;
                    if (qp.gap_ok) {                                                // a gap clause: sparse pre-pass (dense4.cuh)
                        const u32 g = atomicAdd(gap_count, 1u);
                        if (g < qp.gap_cap) {
                            GapClause gc;
                            gc.k0 = tok_key[base + j0]; gc.k1 = tok_key[base + j1]; gc.slop = (int)delta + 1; gc.read = r;
                            gaps[g] = gc;
                            clause_row[base + c] = qp.temp_row0 + g;
                            bad_slop = false;
                            is_gap = true;
                        }
                    }
                }
            }
            clause_val[base + c] = idf;
            const float qw = __fmul_rn(idf, 1.0f);
            sq = __fmul_rn(qw, qw);
        }
        if (__any_sync(0xFFFFFFFFu, bad_slop)) dense = false;
        n_gap += __popc(__ballot_sync(0xFFFFFFFFu, is_gap));
        const i32 lim = min(32, nc - c0);
        for (i32 i = 0; i < lim; i++) sum_sq = __fadd_rn(sum_sq, __shfl_sync(0xFFFFFFFFu, sq, i));   // clause order
    }
    if (qp.sim == SIM_TFIDF) {
        sum_sq = __fmul_rn(sum_sq, 1.0f);                                           // BooleanWeight#getValueForNormalization
        float qn = __double2float_rn(__ddiv_rn(1.0, __dsqrt_rn((double)sum_sq)));   // DefaultSimilarity#queryNorm
        if (isinf(qn) || isnan(qn)) qn = 1.0f;                                      // IndexSearcher#createNormalizedWeight
        float vmax = 0.0f, vmin = 3.0e38f;
        for (i32 c = lane; c < nc; c += 32) {
            const float idf = clause_val[base + c];
            const float qw = __fmul_rn(__fmul_rn(idf, 1.0f), __fmul_rn(qn, 1.0f));  // IDFStats#normalize
            const float v = __fmul_rn(qw, idf);
            clause_val[base + c] = v;
            vmax = fmaxf(vmax, v); vmin = fminf(vmin, v);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            vmax = fmaxf(vmax, __shfl_xor_sync(0xFFFFFFFFu, vmax, o));
            vmin = fminf(vmin, __shfl_xor_sync(0xFFFFFFFFu, vmin, o));
        }
        if (lane == 0) vrange[r] = make_float2(vmax, vmin);
    } else {   // BM25Stats#normalize: weight = idf * 1 * 1
        float vmax = 0.0f, vmin = 3.0e38f;
        for (i32 c = lane; c < nc; c += 32) { const float v = clause_val[base + c]; vmax = fmaxf(vmax, v); vmin = fminf(vmin, v); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            vmax = fmaxf(vmax, __shfl_xor_sync(0xFFFFFFFFu, vmax, o));
            vmin = fminf(vmin, __shfl_xor_sync(0xFFFFFFFFu, vmin, o));
        }
        if (lane == 0) vrange[r] = make_float2(vmax, vmin);
    }
    if (dense) {
        const u8* s = seq + seq_off[r];
        for (i32 c = lane; c < nc; c += 32) {
            i32 j0, j1;
            clause_tokens(qp.alg, n, c, j0, j1);
            if (j1 >= 0 && tok_pos[base + j1] - tok_pos[base + j0] != 1u) {
                // gap clause: its temporary row id was stored above
            } else if (j1 >= 0) {      // pair row: the raw (k+1)-mer at the first token's offset
                const u32 off = tok_pos[base + j0];
                u64 u = 0;
                for (int i = 0; i <= qp.k; i++) u = (u << 2) | (u64)base_code_dev(s[off + i]);
                clause_row[base + c] = (u32)(qp.pair_row0 + u);
            } else {
                clause_row[base + c] = (u32)tok_key[base + j0];
            }
        }
    }
    if (lane == 0) {
        n_clauses[r] = nc;
        read_ngap[r] = n_gap;
        msm_out[r] = (i32)(qp.msm_ratio * (double)nc);   // Classifier.java:257
        // u8 count lanes of the filter kernel: <= 255 clauses (the u16 weight lanes then hold 255 * 255 at most)
        const bool fast = dense && qp.filter_ok && nc <= qp.filter_nc_max;
        if (fast) { route[r] = ROUTE_FILTER; status[r] = ST_OK; }
        else if (dense) { route[r] = ROUTE_DENSE; status[r] = ST_OK; }
        else if (qp.positional_ok) { route[r] = ROUTE_POSITIONAL; status[r] = ST_OK; }
        else { route[r] = ROUTE_NONE; status[r] = ST_ERROR; }
    }
}
__global__ void __launch_bounds__(256) weights_kernel(const u8* __restrict__ seq, const i64* __restrict__ seq_off, i32 n_reads,
                                                       const i64* __restrict__ tok_off, const i32* __restrict__ n_tok,
                                                       const u64* __restrict__ tok_key, const u32* __restrict__ tok_pos,
                                                       const u32* __restrict__ tok_df, const float* __restrict__ idf_by_df,
                                                       QueryParams qp,
                                                       float* __restrict__ clause_val, u32* __restrict__ clause_row,
                                                       i32* __restrict__ n_clauses, i32* __restrict__ msm_out,
                                                       i32* __restrict__ route, i32* __restrict__ status, float2* __restrict__ vrange,
                                                       GapClause* __restrict__ gaps, u32* __restrict__ gap_count,
                                                       i32* __restrict__ read_ngap) {
    const int lane = threadIdx.x & 31;
    const i32 r = (i32)((blockIdx.x * (u64)blockDim.x + threadIdx.x) >> 5);
    if (r >= n_reads) return;
    weights_read(seq, seq_off, r, lane, tok_off, n_tok[r], tok_key, tok_pos, tok_df, idf_by_df, qp, clause_val, clause_row, n_clauses,
                 msm_out, route, status, vrange, gaps, gap_count, read_ngap);
}

// The three kernels in front of the scorer as one (kmer_skips = 0, direct term table): a warp tokenises its read out of
// shared-memory staged codes, looks each token's df up as it emits it and goes on to the clause weights and the route
// while its tokens are still in L1 -- one launch and one pass over the keys instead of three.
__global__ void __launch_bounds__(256) front_kernel(const u8* __restrict__ seq, const i64* __restrict__ seq_off, i32 n_reads,
                                                     const i64* __restrict__ tok_off, int min_strand, const u32* __restrict__ df_dense,
                                                     u64* __restrict__ tok_key, u32* __restrict__ tok_pos, u32* __restrict__ tok_df,
                                                     i32* __restrict__ n_tok, const float* __restrict__ idf_by_df, QueryParams qp,
                                                     float* __restrict__ clause_val, u32* __restrict__ clause_row,
                                                     i32* __restrict__ n_clauses, i32* __restrict__ msm_out,
                                                     i32* __restrict__ route, i32* __restrict__ status, float2* __restrict__ vrange,
                                                     GapClause* __restrict__ gaps, u32* __restrict__ gap_count,
                                                     i32* __restrict__ read_ngap) {
    __shared__ u8 s_codes[8][64];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const i32 r = (i32)((blockIdx.x * (u64)blockDim.x + threadIdx.x) >> 5);
    if (r >= n_reads) return;
    tokenize_read(seq, seq_off, r, lane, tok_off, qp.k, 0, min_strand, tok_key, tok_pos, n_tok, s_codes[warp], df_dense, tok_df);
    __syncwarp();                                    // the warp's own token arrays (global memory, same SM) are complete
    i32 n = 0;
    if (lane == 0) n = n_tok[r];
    n = __shfl_sync(0xFFFFFFFFu, n, 0);
    weights_read(seq, seq_off, r, lane, tok_off, n, tok_key, tok_pos, tok_df, idf_by_df, qp, clause_val, clause_row, n_clauses,
                 msm_out, route, status, vrange, gaps, gap_count, read_ngap);
}

// ------------------------------------------------------------------ scoring math
struct SimParams {
    int sim;
    float k1p1;              // BM25: k1 + 1f
};
// doc_w[d]: tf-idf -> decoded norm (DefaultSimilarity#decodeNormValue); BM25 -> cache[normByte]
__device__ __forceinline__ float clause_doc_score(int sim, float freq_or_tf, float value, float doc_w, float k1p1) {
    if (sim == SIM_TFIDF)      // TFIDFSimScorer#score: (tf(freq) * value) * norm ; freq_or_tf = tf(freq)
        return __fmul_rn(__fmul_rn(freq_or_tf, value), doc_w);
    // BM25DocScorer#score: weightValue * freq / (freq + norm), weightValue = weight * (k1 + 1)
    return __fdiv_rn(__fmul_rn(__fmul_rn(value, k1p1), freq_or_tf), __fadd_rn(freq_or_tf, doc_w));
}
__device__ __forceinline__ float tf_of(float freq) {    // DefaultSimilarity#tf
    return __double2float_rn(__dsqrt_rn((double)freq));
}
__device__ __forceinline__ float final_score(int sim, double sum, i32 m, i32 nc) {
    float coord = 1.0f;      // BooleanWeight#coord / DefaultSimilarity#coord ; BM25: Similarity#coord = 1
    if (sim == SIM_TFIDF && nc != 1) coord = __fdiv_rn((float)m, (float)nc);
    return __fmul_rn(__double2float_rn(sum), coord);
}

// order key: larger = better (score desc, doc asc); scores are >= 0; 0 = no candidate
__device__ __forceinline__ u64 cand_key(float score, u32 doc) {
    return ((u64)__float_as_uint(score) << 32) | (u64)(0xFFFFFFFFu - doc);
}

__device__ __forceinline__ u64 block_max_u64(u64 v, u64* sm /* [33] */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        u64 o = __shfl_xor_sync(0xFFFFFFFFu, v, d);
        if (o > v) v = o;
    }
    __syncthreads();
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    if (warp == 0) {
        u64 w = lane < (int)((blockDim.x + 31) >> 5) ? sm[lane] : 0ull;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            u64 o = __shfl_xor_sync(0xFFFFFFFFu, w, d);
            if (o > w) w = o;
        }
        if (lane == 0) sm[32] = w;
    }
    __syncthreads();
    return sm[32];
}

// ------------------------------------------------------------------- K6/K7
// General path.  CTA per (segment, read): per-doc accumulators for the segment's docs
// live in shared memory (double sum + u16 match count); clauses are walked one at a
// time (each doc occurs at most once per clause -> no atomics), threads stride over
// the postings of the clause's first term; then the fused block top-10.
#define POS_THREADS 256
#define POS_CHUNK 128            // clauses staged per round
struct ClauseStage { u32 a0, b0, a1, b1; float val; int slop; };   // slop 0 = Term, <0 = same-term phrase (-slop)
// the warp kernel also keeps each term's position range of the segment: with it every copy of a clause can be issued at once
struct ClauseStageW { u32 a0, b0, a1, b1; float val; int slop; u32 p0, m0, p1, m1; };
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
    const u32 sa = (u32)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gmem) : "memory");
}
// 16-byte copy global -> shared, bypassing L1 (both addresses 16-byte aligned)
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const u32 sa = (u32)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

__global__ void __launch_bounds__(POS_THREADS) score_positional_kernel(
        const SegDev* __restrict__ segs, int n_segs, const i32* __restrict__ read_list, i32 n_list,
        const i64* __restrict__ tok_off, const i32* __restrict__ n_tok, const u64* __restrict__ tok_key,
        const u32* __restrict__ tok_pos, const float* __restrict__ clause_val, const i32* __restrict__ n_clauses,
        const i32* __restrict__ msm, const float* __restrict__ doc_w, QueryParams qp, SimParams sp, u32 global_doc_base,
        i32* __restrict__ part_n, u32* __restrict__ part_doc, float* __restrict__ part_score) {
    extern __shared__ __align__(16) u8 smem_raw[];
    __shared__ ClauseStage stage[POS_CHUNK];
    __shared__ u64 red[33];
    // segment-major: the CTAs resident at any moment work on one or two segments, so their postings, term tables and
    // page translations stay hot (read-major order touched all of the 100+ GB index at once)
    const int sidx = blockIdx.x / n_list;
    const i32 r = read_list[blockIdx.x % n_list];
    const SegDev s = segs[sidx];
    double* acc = reinterpret_cast<double*>(smem_raw);
    u16* cnt = reinterpret_cast<u16*>(smem_raw + (size_t)s.n_docs * 8);
    for (u32 d = threadIdx.x; d < s.n_docs; d += blockDim.x) { acc[d] = 0.0; cnt[d] = 0; }
    const i64 base = tok_off[r];
    const i32 n = n_tok[r], nc = n_clauses[r];
    const float* dw = doc_w + s.doc_base;
    for (i32 c0 = 0; c0 < nc; c0 += POS_CHUNK) {
        __syncthreads();
        i32 c = c0 + threadIdx.x;
        if (threadIdx.x < POS_CHUNK && c < nc) {
            i32 j0, j1;
            clause_tokens(qp.alg, n, c, j0, j1);
            ClauseStage st;
            st.val = clause_val[base + c];
            u64 k0 = tok_key[base + j0];
            u32 t0 = seg_find_term(s, k0);
            st.a0 = st.b0 = st.a1 = st.b1 = 0; st.slop = 0;
            if (t0 != NO_TERM) { st.a0 = s.term_off[t0]; st.b0 = s.term_off[t0 + 1]; }
            if (j1 >= 0) {
                u64 k1 = tok_key[base + j1];
                int slop = (int)(tok_pos[base + j1] - tok_pos[base + j0]) + 1;       // Classifier.java:146,182
                if (k1 == k0) { st.slop = -slop; st.a1 = st.a0; st.b1 = st.b0; }
                else {
                    st.slop = slop;
                    u32 t1 = seg_find_term(s, k1);
                    if (t1 != NO_TERM) { st.a1 = s.term_off[t1]; st.b1 = s.term_off[t1 + 1]; }
                    else st.b0 = st.a0;                                              // conjunction is empty
                }
            }
            stage[threadIdx.x] = st;
        }
        __syncthreads();
        const i32 lim = min(POS_CHUNK, nc - c0);
        for (i32 ci = 0; ci < lim; ci++) {
            const ClauseStage st = stage[ci];
            for (u32 i = st.a0 + threadIdx.x; i < st.b0; i += blockDim.x) {
                const u32 d = s.post_doc[i];
                float f;
                if (st.slop == 0) {
                    f = (float)(s.post_pos[i + 1] - s.post_pos[i]);                  // TermScorer: freq as float
                } else {
                    const bool same = st.slop < 0;
                    f = phrase_freq_at(s, i, st.a1, st.b1, st.a1 + (i - st.a0), same ? -st.slop : st.slop, same);
                }
                if (f != 0.0f) {
                    float sc = clause_doc_score(sp.sim, sp.sim == SIM_TFIDF ? tf_of(f) : f, st.val, dw[d], sp.k1p1);
                    acc[d] += (double)sc;                                            // BooleanScorer: double sum
                    cnt[d] += 1;
                }
            }
            __syncthreads();
        }
    }
    __syncthreads();
    // final scores (in place: float in the low half of the double slot), msm filter
    const i32 need = max(1, msm[r]);
    for (u32 d = threadIdx.x; d < s.n_docs; d += blockDim.x) {
        i32 m = cnt[d];
        float sc = -1.0f;
        if (m >= need) sc = final_score(sp.sim, acc[d], m, nc);
        reinterpret_cast<float*>(&acc[d])[0] = sc;
    }
    __syncthreads();
    // fused top-10: rounds of block arg-max
    const i64 pbase = (i64)(blockIdx.x % n_list) * n_segs + sidx;        // (entry of the read list, segment)
    i32 found = 0;
    for (int round = 0; round < TOPN; round++) {
        u64 best = 0;
        for (u32 d = threadIdx.x; d < s.n_docs; d += blockDim.x) {
            float sc = reinterpret_cast<float*>(&acc[d])[0];
            if (sc >= 0.0f) { u64 kk = cand_key(sc, d); if (kk > best) best = kk; }
        }
        best = block_max_u64(best, red);
        if (best == 0) break;
        u32 d = 0xFFFFFFFFu - (u32)(best & 0xFFFFFFFFu);
        if (threadIdx.x == 0) {
            part_doc[pbase * TOPN + round] = global_doc_base + s.doc_base + d;
            part_score[pbase * TOPN + round] = __uint_as_float((u32)(best >> 32));
            reinterpret_cast<float*>(&acc[d])[0] = -1.0f;
        }
        found++;
        __syncthreads();
    }
    if (threadIdx.x == 0) part_n[pbase] = found;
}

// Same scorer, one WARP per (segment, read), for segments of at most POSW_DOCS documents (large genomes: a 2^28-token
// segment holds ~80 docs at 3 Mbp each).  No block barriers: the warp stages 32 clauses at a time, walks each clause's
// postings 32 at a time (a doc occurs once per clause -> plain shared-memory updates), and selects its top 10 with
// shuffles.  POSW_WARPS pairs per CTA.
#define POSW_DOCS 256
#define POSW_WARPS 8
#ifndef POSW_MINB
#define POSW_MINB 4                // 64 registers, 4 CTAs per SM (3 CTAs at 80 registers: 0.75 M reads/s against 0.87 M at C2, skips 5)
#endif
// Phrase clauses are staged: the warp copies both terms' posting slices of the segment (doc ids, position offsets)
// and their position ranges -- each one contiguous in the term-major layout -- into shared memory as one batch of
// asynchronous 4-byte copies (the ranges are looked up when the clauses are set up), and the conjunction + sloppy sweep
// run out of shared memory: one memory round trip per clause instead of a dozen dependent scalar loads per
// (clause, doc).  Slices beyond the position cap take the direct path.
// Shared memory per warp (all dynamic, sized by the host from the largest segment): double acc[dcap] | PhraseStage
// {u32 pos[2][pcap] | u32 doc[2][dcap] | u32 pp[2][dcap + 4] | u8 inv[dcap] | u8 hit[dcap]} | ClauseStageW stage[32] |
// u16 cnt[dcap];  dcap = docs per segment rounded up to 32 (a term has at most one posting per doc), pcap = PW_POS
// positions per term and segment.  inv[doc] = index of the doc's posting in the second term's slice (possibly stale: always verified);
// hit[i] = posting i of the first term has a position within the slop of one of the second term's.
//
// Which docs can match at all?  Every matchLength the SloppyPhraseScorer sweep tests is |(b - 1) - a| of an actual
// pair (a in A, b in B), so freq = 0 unless some pair is within the slop.  Positions are segment ordinals
// (index_build.cuh), hence both staged slices ascend across docs and ONE scan per clause finds all close pairs: each
// lane takes a contiguous chunk of the first term's positions and walks the second term's from a binary-searched start.
// No false negatives; a pair straddling two docs is a false positive the exact sweep discards.
#define PW_POS 384               // 2^28 tokens / 4^10 terms = 256 on average
// per-warp staging buffers of one phrase clause
struct PhraseStage { u32 *pos0, *pos1, *doc0, *doc1, *pp0, *pp1; u8 *inv, *hit; };
__host__ __device__ inline size_t phrase_stage_bytes(int dcap, int pcap) {
    return (size_t)pcap * 8 + (size_t)(dcap + 4) * 8 + (size_t)(dcap + 4) * 8 + (size_t)dcap * 2;
}
__device__ __forceinline__ PhraseStage phrase_stage_at(u8* base /* 4-byte aligned */, int dcap, int pcap) {
    PhraseStage W;
    W.pos0 = reinterpret_cast<u32*>(base);
    W.pos1 = W.pos0 + pcap;
    W.doc0 = W.pos1 + pcap;
    W.doc1 = W.doc0 + dcap + 4;
    W.pp0 = W.doc1 + dcap + 4;
    W.pp1 = W.pp0 + dcap + 4;
    W.inv = reinterpret_cast<u8*>(W.pp1 + dcap + 4);
    W.hit = W.inv + dcap;
    return W;
}
// The docs of the segment in which the phrase (t0@0, t1@1, slop) with t0 != t1 has a non-zero sloppy frequency:
// on_match(local doc, freq) is called once per such doc (by whichever lane owns the posting).  Postings [a0, a0+n0) /
// [a1, a1+n1) with position ranges [p0, p0+m0) / [p1, p1+m1); n0, n1 <= dcap, m0 + 3, m1 + 3 <= pcap, inv[] and hit[]
// hold anything / zeros on entry and are left that way.  All 32 lanes call it.  The position ranges travel as 16-byte
// copies from the enclosing aligned range (the arrays are 256-byte aligned and padded), so they start up to 3 elements
// into their buffers.
template <typename F>
__device__ __forceinline__ void phrase_matches_staged(const SegDev& s, const PhraseStage& W, u32 a0, u32 n0, u32 p0, u32 m0,
                                                      u32 a1, u32 n1, u32 p1, u32 m1, int slop, int lane, F&& on_match) {
    // one batch of asynchronous 16-byte copies (LDGSTS) for the whole clause: one memory round trip.  Every slice is
    // copied from its enclosing 16-byte aligned range, so it starts up to 3 elements into its buffer.
    const u32 s0 = a0 & 3u, s1 = a1 & 3u;
    for (u32 i = 4u * (u32)lane; i < n0 + 1 + s0; i += 128) {
        cp_async16(W.doc0 + i, s.post_doc + (a0 - s0) + i);
        cp_async16(W.pp0 + i, s.post_pos + (a0 - s0) + i);
    }
    for (u32 i = 4u * (u32)lane; i < n1 + 1 + s1; i += 128) {
        cp_async16(W.doc1 + i, s.post_doc + (a1 - s1) + i);
        cp_async16(W.pp1 + i, s.post_pos + (a1 - s1) + i);
    }
    const u32* doc0 = W.doc0 + s0; const u32* pp0 = W.pp0 + s0;
    const u32* doc1 = W.doc1 + s1; const u32* pp1 = W.pp1 + s1;
    const u32 o0 = p0 & 3u, o1 = p1 & 3u;
    for (u32 i = 4u * (u32)lane; i < m0 + o0; i += 128) cp_async16(W.pos0 + i, s.positions + (p0 - o0) + i);
    for (u32 i = 4u * (u32)lane; i < m1 + o1; i += 128) cp_async16(W.pos1 + i, s.positions + (p1 - o1) + i);
    const u32* A = W.pos0 + o0;
    const u32* B = W.pos1 + o1;
    cp_async_commit_wait_all();
    __syncwarp();
    for (u32 i = lane; i < n1; i += 32) W.inv[doc1[i]] = (u8)i;
    __syncwarp();
    // ---- close pairs: lane takes positions [ia, ia1) of the first term
    const u32 ca = (m0 + 31) >> 5;
    u32 ia = (u32)lane * ca;
    const u32 ia1 = min(m0, ia + ca);
    if (ia < ia1) {
        const int lo = (int)A[ia] - slop;
        u32 jb = 0, hi = m1;                                                 // first b' >= lo, b' = b - 1 (phrase offset of t1)
        while (jb < hi) {
            const u32 mid = (jb + hi) >> 1;
            if ((int)B[mid] - 1 < lo) jb = mid + 1; else hi = mid;
        }
        while (ia < ia1 && jb < m1) {
            const int d = (int)B[jb] - 1 - (int)A[ia];
            if (d > slop) ia++;                                              // b' beyond this a: next a
            else if (d < -slop) jb++;                                        // b' behind: next b'
            else {                                                           // within the slop: flag a's posting
                const u32 target = p0 + ia;
                u32 l = 0, h2 = n0;                                          // last posting with pp <= target
                while (h2 - l > 1) {
                    const u32 mid = (l + h2) >> 1;
                    if (pp0[mid] <= target) l = mid; else h2 = mid;
                }
                W.hit[l] = 1;
                ia++;
            }
        }
    }
    __syncwarp();
    // ---- the literal sweep for the flagged docs
    for (u32 i = lane; i < n0; i += 32) {
        if (!W.hit[i]) continue;
        W.hit[i] = 0;
        const u32 d = doc0[i];
        const u32 j = W.inv[d];                                              // verified: may be a stale entry
        if (j < n1 && doc1[j] == d) {
            const u32 q0 = pp0[i], q1 = pp1[j];
            const float f = sloppy_freq_distinct(A + (q0 - p0), (int)(pp0[i + 1] - q0),
                                                 B + (q1 - p1), (int)(pp1[j + 1] - q1), slop);
            if (f != 0.0f) on_match(d, f);
        }
    }
    __syncwarp();
}
__host__ __device__ inline size_t posw_warp_bytes(int dcap, int pcap) {
    return (size_t)dcap * 8 + phrase_stage_bytes(dcap, pcap) + 32 * sizeof(ClauseStageW) + (size_t)dcap * 2;
}
__global__ void __launch_bounds__(POSW_WARPS * 32, POSW_MINB) score_positional_warp_kernel(
        const SegDev* __restrict__ segs, int n_segs, const i32* __restrict__ read_list, i32 n_list,
        const i64* __restrict__ tok_off, const i32* __restrict__ n_tok, const u64* __restrict__ tok_key,
        const u32* __restrict__ tok_pos, const float* __restrict__ clause_val, const i32* __restrict__ n_clauses,
        const i32* __restrict__ msm, const float* __restrict__ doc_w, QueryParams qp, SimParams sp, u32 global_doc_base,
        i32* __restrict__ part_n, u32* __restrict__ part_doc, float* __restrict__ part_score, int stage_ok,
        int dcap /* >= docs of any segment, multiple of 32 */, int pcap) {
    extern __shared__ __align__(16) u8 pw_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u8* wbase = pw_raw + (size_t)warp * posw_warp_bytes(dcap, pcap);
    double* acc = reinterpret_cast<double*>(wbase);
    const PhraseStage W = phrase_stage_at(wbase + (size_t)dcap * 8, dcap, pcap);
    ClauseStageW* stage = reinterpret_cast<ClauseStageW*>(wbase + (size_t)dcap * 8 + phrase_stage_bytes(dcap, pcap));
    u16* cnt = reinterpret_cast<u16*>(stage + 32);
    const i64 pair = (i64)blockIdx.x * POSW_WARPS + warp;
    if (pair >= (i64)n_list * n_segs) return;
    const int sidx = (int)(pair / n_list);              // segment-major, see score_positional_kernel
    const i32 r = read_list[pair % n_list];
    const SegDev s = segs[sidx];
    for (u32 d = lane; d < (u32)dcap; d += 32) { acc[d] = 0.0; cnt[d] = 0; W.inv[d] = 0; W.hit[d] = 0; }
    const i64 base = tok_off[r];
    const i32 n = n_tok[r], nc = n_clauses[r];
    const float* dw = doc_w + s.doc_base;
    for (i32 c0 = 0; c0 < nc; c0 += 32) {
        __syncwarp();
        const i32 c = c0 + lane;
        if (c < nc) {
            i32 j0, j1;
            clause_tokens(qp.alg, n, c, j0, j1);
            ClauseStageW st;
            st.val = clause_val[base + c];
            const u64 k0 = tok_key[base + j0];
            const u32 t0 = seg_find_term(s, k0);
            st.a0 = st.b0 = st.a1 = st.b1 = 0; st.slop = 0;
            if (t0 != NO_TERM) { st.a0 = s.term_off[t0]; st.b0 = s.term_off[t0 + 1]; }
            if (j1 >= 0) {
                const u64 k1 = tok_key[base + j1];
                const int slop = (int)(tok_pos[base + j1] - tok_pos[base + j0]) + 1;     // Classifier.java:146,182
                if (k1 == k0) { st.slop = -slop; st.a1 = st.a0; st.b1 = st.b0; }
                else {
                    st.slop = slop;
                    const u32 t1 = seg_find_term(s, k1);
                    if (t1 != NO_TERM) { st.a1 = s.term_off[t1]; st.b1 = s.term_off[t1 + 1]; }
                    else st.b0 = st.a0;                                                  // conjunction is empty
                }
            }
            // position ranges of both slices, so that the clause loop can issue every copy of a clause in one batch
            st.p0 = st.m0 = st.p1 = st.m1 = 0;
            if (st.slop > 0 && st.b0 > st.a0 && st.b1 > st.a1) {
                st.p0 = s.post_pos[st.a0]; st.m0 = s.post_pos[st.b0] - st.p0;
                st.p1 = s.post_pos[st.a1]; st.m1 = s.post_pos[st.b1] - st.p1;
            }
            stage[lane] = st;
        }
        __syncwarp();
        const i32 lim = min(32, nc - c0);
        for (i32 ci = 0; ci < lim; ci++) {
            const ClauseStageW st = stage[ci];
            const u32 n0 = st.b0 - st.a0, n1 = st.b1 - st.a1;
            bool staged = false;
            if (stage_ok && st.slop > 0 && n0 && n1 && st.m0 + 3 <= (u32)pcap && st.m1 + 3 <= (u32)pcap) {   // (n0, n1 <= docs <= dcap)
                phrase_matches_staged(s, W, st.a0, n0, st.p0, st.m0, st.a1, n1, st.p1, st.m1, st.slop, lane, [&](u32 d, float f) {
                    const float sc = clause_doc_score(sp.sim, sp.sim == SIM_TFIDF ? tf_of(f) : f, st.val, dw[d], sp.k1p1);
                    acc[d] += (double)sc;                                            // BooleanScorer: double sum
                    cnt[d] += 1;
                });
                staged = true;
            }
            if (!staged) {
                for (u32 i = st.a0 + lane; i < st.b0; i += 32) {
                    const u32 d = s.post_doc[i];
                    float f;
                    if (st.slop == 0) f = (float)(s.post_pos[i + 1] - s.post_pos[i]);   // TermScorer: freq as float
                    else {
                        const bool same = st.slop < 0;
                        f = phrase_freq_at(s, i, st.a1, st.b1, st.a1 + (i - st.a0), same ? -st.slop : st.slop, same);
                    }
                    if (f != 0.0f) {
                        const float sc = clause_doc_score(sp.sim, sp.sim == SIM_TFIDF ? tf_of(f) : f, st.val, dw[d], sp.k1p1);
                        acc[d] += (double)sc;                                            // BooleanScorer: double sum
                        cnt[d] += 1;
                    }
                }
            }
            __syncwarp();
        }
    }
    __syncwarp();
    // final scores, msm filter, per-lane candidates (lane owns docs lane, lane+32, ...: at most 8)
    const i32 need = max(1, msm[r]);
    u64 keys[POSW_DOCS / 32];
#pragma unroll
    for (int j = 0; j < POSW_DOCS / 32; j++) {
        const u32 d = lane + 32 * j;
        keys[j] = 0;
        if (d < s.n_docs) {
            const i32 m = cnt[d];
            if (m >= need) keys[j] = cand_key(final_score(sp.sim, acc[d], m, nc), d);
        }
    }
    const i64 pbase = (pair % n_list) * n_segs + sidx;                   // (entry of the read list, segment)
    i32 found = 0;
    for (int round = 0; round < TOPN; round++) {
        u64 lm = 0;
#pragma unroll
        for (int j = 0; j < POSW_DOCS / 32; j++) if (keys[j] > lm) lm = keys[j];
        u64 best = lm;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const u64 x = __shfl_xor_sync(0xFFFFFFFFu, best, o);
            if (x > best) best = x;
        }
        if (best == 0) break;
#pragma unroll
        for (int j = 0; j < POSW_DOCS / 32; j++) if (keys[j] == best) keys[j] = 0;
        if (lane == 0) {
            part_doc[pbase * TOPN + round] = global_doc_base + s.doc_base + (0xFFFFFFFFu - (u32)(best & 0xFFFFFFFFu));
            part_score[pbase * TOPN + round] = __uint_as_float((u32)(best >> 32));
        }
        found++;
    }
    if (lane == 0) part_n[pbase] = found;
}

// ------------------------------------------------------------------------ K8
// Thread per read: merge the per-part top-10 lists (each sorted score desc, doc asc)
// into the final list; n_hits = entries equal to the max (Classifier.java:358-366);
// label by score tier (Classifier.java:263-339 without taxonomy).
__global__ void __launch_bounds__(128) merge_kernel(i32 n_reads, const i32* __restrict__ route, const i32* __restrict__ status_in,
                                                     int parts_dense, int parts_pos, int parts_filter, int parts_stride,
                                                     const i32* __restrict__ part_n_d, const u32* __restrict__ part_doc_d,
                                                     const float* __restrict__ part_score_d,
                                                     const i32* __restrict__ pos_slot, const i32* __restrict__ part_n_p,
                                                     const u32* __restrict__ part_doc_p, const float* __restrict__ part_score_p,
                                                     i32* __restrict__ status, i32* __restrict__ label, i32* __restrict__ n_hits,
                                                     i32* __restrict__ n_top, i32* __restrict__ top_doc,
                                                     float* __restrict__ top_score, int full_topn) {
    const i32 r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    float bs[TOPN];
    u32 bd[TOPN];
    int nt = 0;
    const int rt = route ? route[r] : ROUTE_POSITIONAL;
    const int parts = rt == ROUTE_DENSE ? parts_dense : (rt == ROUTE_POSITIONAL ? parts_pos : (rt == ROUTE_FILTER ? parts_filter : 0));
    // the positional scorer keeps its partial lists per (entry of its read list, segment), the table scorers per (read, tile)
    const bool pos = rt == ROUTE_POSITIONAL;
    const i32* part_n = pos ? part_n_p : part_n_d;
    const u32* part_doc = pos ? part_doc_p : part_doc_d;
    const float* part_score = pos ? part_score_p : part_score_d;
    const i64 pb0 = pos ? (i64)(pos_slot ? pos_slot[r] : r) * parts_pos : (i64)r * parts_stride;
    for (int p = 0; p < parts; p++) {
        const i64 pb = pb0 + p;
        const int np = part_n[pb];
        for (int j = 0; j < np; j++) {
            const float s = part_score[pb * TOPN + j];
            const u32 d = part_doc[pb * TOPN + j];
            int pos = nt;
            while (pos > 0 && (bs[pos - 1] < s || (bs[pos - 1] == s && bd[pos - 1] > d))) pos--;
            if (pos >= TOPN) break;                       // the rest of this part is not better
            const int last = nt < TOPN ? nt : TOPN - 1;
            for (int q = last; q > pos; q--) { bs[q] = bs[q - 1]; bd[q] = bd[q - 1]; }
            bs[pos] = s; bd[pos] = d;
            if (nt < TOPN) nt++;
        }
    }
    int nh = 0;
    for (int j = 0; j < nt; j++) if ((double)bs[0] - (double)bs[j] == 0.0) nh++;
    if (!full_topn) nt = nh;        // result-set mode: only the hits equal to the maximum are reported
    if (status) status[r] = status_in[r];
    if (label) label[r] = nh == 0 ? 0 : (nh == 1 ? 2 : 1);
    if (n_hits) n_hits[r] = nh;
    if (n_top) n_top[r] = nt;
    for (int j = 0; j < TOPN; j++) {
        if (top_doc) top_doc[(i64)r * TOPN + j] = j < nt ? (i32)bd[j] : -1;
        if (top_score) top_score[(i64)r * TOPN + j] = j < nt ? bs[j] : 0.0f;
    }
}

// Doc-partitioned mode: one record per read and rank, so that the hit lists cross NVLink in ONE all_gather and the merge
// reads them where the collective left them.  Record = 21 words: n_top | doc[10] (global ids) | score[10] (float bits).
#define LIST_REC_WORDS (1 + 2 * TOPN)
__global__ void __launch_bounds__(256) pack_lists_kernel(i32 n_reads, const i32* __restrict__ n_top, const i32* __restrict__ top_doc,
                                                          const float* __restrict__ top_score, i32* __restrict__ rec) {
    const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (i64)n_reads * LIST_REC_WORDS) return;
    const i32 r = (i32)(t / LIST_REC_WORDS), w = (i32)(t % LIST_REC_WORDS);
    rec[t] = w == 0 ? n_top[r] : (w <= TOPN ? top_doc[(i64)r * TOPN + (w - 1)] : __float_as_int(top_score[(i64)r * TOPN + (w - 1 - TOPN)]));
}
// Thread per read over the gathered records, rank-major as all_gather_into_tensor lays them out: rec[(p * n_reads + r) * 21].
// Same collector order, equal-to-max rule and label as merge_kernel.
__global__ void __launch_bounds__(128) merge_records_kernel(i32 n_reads, int parts, const i32* __restrict__ rec,
                                                             const i32* __restrict__ status_in, i32* __restrict__ status,
                                                             i32* __restrict__ label, i32* __restrict__ n_hits, i32* __restrict__ n_top,
                                                             i32* __restrict__ top_doc, float* __restrict__ top_score, int full_topn) {
    const i32 r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    float bs[TOPN];
    u32 bd[TOPN];
    int nt = 0;
    for (int p = 0; p < parts; p++) {
        const i32* q = rec + ((i64)p * n_reads + r) * LIST_REC_WORDS;
        const int np = min(max(q[0], 0), TOPN);
        for (int j = 0; j < np; j++) {
            const float s = __int_as_float(q[1 + TOPN + j]);
            const u32 d = (u32)q[1 + j];
            int pos = nt;
            while (pos > 0 && (bs[pos - 1] < s || (bs[pos - 1] == s && bd[pos - 1] > d))) pos--;
            if (pos >= TOPN) break;
            const int last = nt < TOPN ? nt : TOPN - 1;
            for (int x = last; x > pos; x--) { bs[x] = bs[x - 1]; bd[x] = bd[x - 1]; }
            bs[pos] = s; bd[pos] = d;
            if (nt < TOPN) nt++;
        }
    }
    int nh = 0;
    for (int j = 0; j < nt; j++) if ((double)bs[0] - (double)bs[j] == 0.0) nh++;
    if (!full_topn) nt = nh;
    if (status) status[r] = status_in[r];
    if (label) label[r] = nh == 0 ? 0 : (nh == 1 ? 2 : 1);
    if (n_hits) n_hits[r] = nh;
    if (n_top) n_top[r] = nt;
    for (int j = 0; j < TOPN; j++) {
        if (top_doc) top_doc[(i64)r * TOPN + j] = j < nt ? (i32)bd[j] : -1;
        if (top_score) top_score[(i64)r * TOPN + j] = j < nt ? bs[j] : 0.0f;
    }
}

// Taxonomy-aware label of a read from its hit list (Classifier.java:263-339, after the hits equal to the maximum are
// known).  lin_off/lin_tax: per document the taxids of TaxonTreeDescription#getClassifiableTaxonomyTree (the entries
// with a rank, leaf -> root); flags bit0 = the document has a taxonomy string, bit1 = it is malformed (the reference
// throws while parsing it and the caller skips the read: label -1).  tax_index = position in hit 0's ranked lineage of
// the taxon the read is classified at, -1 = ("unknown", "").  Thread per read.
__global__ void __launch_bounds__(128) taxonomy_label_kernel(i32 n_reads, const i32* __restrict__ n_hits, const i32* __restrict__ top_doc,
                                                              const i64* __restrict__ lin_off, const i32* __restrict__ lin_tax,
                                                              const u8* __restrict__ flags, i64 n_docs, i32* __restrict__ tax_label,
                                                              i32* __restrict__ tax_index) {
    const i32 r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    const i32 nh = min(max(n_hits[r], 0), TOPN);
    const i32* hd = top_doc + (i64)r * TOPN;
    i32 tl = 0, ti = -1;                                      // no hit: UNKNOWN
    bool bad = false;
    for (int j = 0; j < nh; j++) if (hd[j] < 0 || hd[j] >= n_docs) bad = true;
    if (bad) tl = -1;
    else if (nh == 1) {                                       // one hit: CLASSIFIED at its lowest ranked taxon (:266-279)
        tl = 2;
        const i32 d = hd[0];
        const u8 f = flags[d];
        if (f & 1) { if (f & 2) tl = -1; else if (lin_off[d + 1] > lin_off[d]) ti = 0; }
    } else if (nh >= 2) {                                     // several: first taxon of hit 0 common to all (:280-338)
        bool all = true;
        for (int j = 0; j < nh; j++) {                        // entries are parsed in order; a missing string fails quickly
            const u8 f = flags[hd[j]];
            if (!(f & 1)) { all = false; break; }
            if (f & 2) { bad = true; break; }
        }
        tl = 1;
        if (bad) tl = -1;
        else if (all) {
            const i64 a0 = lin_off[hd[0]], b0 = lin_off[hd[0] + 1];
            for (i64 x = a0; x < b0 && tl == 1; x++) {
                const i32 t = lin_tax[x];
                bool common = true;
                for (int j = 1; j < nh && common; j++) {
                    bool found = false;
                    for (i64 y = lin_off[hd[j]]; y < lin_off[hd[j] + 1]; y++) if (lin_tax[y] == t) { found = true; break; }
                    common = found;
                }
                if (common) { tl = 2; ti = (i32)(x - a0); }
            }
        }
    }
    tax_label[r] = tl;
    tax_index[r] = ti;
}

// ------------------------------------------------------------- cost model (8d)
// Warp per read.  out[0] reads, [1] tokens, [2] unique terms (= probes), [3] sum df,
// [4] sum ttf, [5] A_read (CSR formula of SURVEY.md 8d), [6] dense-layout bytes, [7] clauses.
__global__ void __launch_bounds__(256) read_cost_kernel(const i64* __restrict__ seq_off, i32 n_reads, const i64* __restrict__ tok_off,
                                                         const i32* __restrict__ n_tok, const u64* __restrict__ tok_key,
                                                         const u32* __restrict__ tok_df, const u32* __restrict__ ttf_dense,
                                                         const SegDev* __restrict__ segs, int n_segs, int alg, u64 row_bytes,
                                                         u64 n_docs, unsigned long long* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const i32 r = (i32)((blockIdx.x * (u64)blockDim.x + threadIdx.x) >> 5);
    if (r >= n_reads) return;
    const i64 base = tok_off[r];
    const i32 n = n_tok[r];
    const i32 nc = clause_count(alg, n);
    unsigned long long uniq = 0, sdf = 0, sttf = 0;
    if (nc > 0 && nc <= MAX_CLAUSES) {
        for (i32 j = lane; j < n; j += 32) {
            u64 key = tok_key[base + j];
            bool first = true;
            for (i32 q = 0; q < j; q++) if (tok_key[base + q] == key) { first = false; break; }
            if (!first) continue;
            uniq++;
            sdf += tok_df[base + j];
            if (ttf_dense) sttf += ttf_dense[key];
            else for (int s = 0; s < n_segs; s++) {
                u32 t = seg_find_term(segs[s], key);
                if (t != NO_TERM) sttf += segs[s].post_pos[segs[s].term_off[t + 1]] - segs[s].post_pos[segs[s].term_off[t]];
            }
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        uniq += __shfl_xor_sync(0xFFFFFFFFu, uniq, d);
        sdf += __shfl_xor_sync(0xFFFFFFFFu, sdf, d);
        sttf += __shfl_xor_sync(0xFFFFFFFFu, sttf, d);
    }
    if (lane == 0 && nc > 0 && nc <= MAX_CLAUSES) {
        unsigned long long len = (unsigned long long)(seq_off[r + 1] - seq_off[r]);
        unsigned long long P = alg == ALG_NAIVE ? 0ull : 1ull;
        atomicAdd(&out[0], 1ull);
        atomicAdd(&out[1], (unsigned long long)n);
        atomicAdd(&out[2], uniq);
        atomicAdd(&out[3], sdf);
        atomicAdd(&out[4], sttf);
        atomicAdd(&out[5], len + 16ull * uniq + 8ull * sdf + P * 4ull * sttf + 80ull);
        atomicAdd(&out[6], len + (unsigned long long)nc * (unsigned long long)row_bytes + 4ull * n_docs + 80ull);
        atomicAdd(&out[7], (unsigned long long)nc);
    }
}
