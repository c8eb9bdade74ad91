// dense4.cuh -- 4-bit dense frequency tables and the two scorers that stream them.
//
// Layout in HBM (per handle): one table of rows x row_bytes, row_bytes = d_pad/2 (d_pad =
// docs rounded up to 32, so a row is a whole number of 16-byte vectors).  Cell (row, doc) is
// the nibble at byte doc/2, bits 4*(doc&1): 8 consecutive docs per little-endian u32.
//   term rows   [0, 4^k)            cell = tf(term, doc)                       if 1 <= tf <= 14
//   pair rows   [4^k, 4^k+4^(k+1))  cell = sloppyFreq(t0,t1,slop 2) in the doc  if it is an integer 1..7
//                                   (row u = the raw (k+1)-mer; t0 = canon(u[0:k]), t1 = canon(u[1:k+1]))
//   everything else (no match, or a frequency the nibble cannot hold) is 0; the latter are
//   kept exactly in the exception list: per row, sorted (doc, float freq) pairs.
//
// Scorers (CTA per read):
//   filter_score_kernel   tf-idf fast path.  Pass 1 streams the read's rows once with packed
//                         integer SIMD: a per-clause 8-entry byte LUT applied with PRMT turns 8
//                         nibbles into quantised weights (q = round(S*tf(f)*value_c)), summed in
//                         packed lanes (x += the four weight bytes as one integer, y += the odd
//                         bytes; u16 lanes are recovered after the loop), match counts in u8 lanes:
//                         ~2 issue slots per (clause, doc) cell, so the kernel is bound by the HBM
//                         stream and the issue rate, not by fp64.  The phases around the loop are
//                         kept short (warp-wide candidate pass, batched cell loads): they are serial
//                         chains that the other warps of the CTA wait for.  The sums give
//                         rigorous lower/upper bounds of Lucene's score; docs whose upper bound
//                         reaches the 10th best lower bound (plus docs with exception cells) are
//                         rescored exactly (float ops in Lucene's order, double accumulation) and the
//                         top-10 is taken among them -- identical to scoring every doc exactly.
//   exact_score_kernel    every doc exact (tf-idf or BM25, any clause count, doc tiles): BM25, reads
//                         with > 255 clauses, candidate overflow, and the reference path for tests.
#pragma once
#include "classify.cuh"

#define T4_TF_MAX 14u
#define T4_PAIR_MAX 7u

struct Tab4Dev {
    const u8* tab;
    u64 row_bytes;
    u64 pair_row0;          // 4^k
    const u32* exc_off;     // [n_rows + 2]
    const u32* exc_doc;     // table columns (see docmap), ascending within a row
    const float* exc_freq;
    u32 zero_row;           // = n_rows: an all-zero row without exceptions
    u32 temp_row0;          // = n_rows + 1: temporary rows of the current sub-batch's gap clauses
    u32 tmp_base;           // their match lists (sorted by column) live at exc_doc/exc_freq[tmp_base + gap_off[g] ...]
    const u32* gap_off;     // [gap clauses + 1] first match of every gap clause (gapjoin.cuh)
    // Table columns are the handle's documents sorted by (norm byte, doc id): docs that share a length norm sit next to
    // each other, so a thread's 32 columns belong to one norm class except at the few class boundaries (the BM25 bound
    // needs a per-class LUT; tf-idf is indifferent to the order).  docmap[column] = doc id (0xFFFFFFFF for padding
    // columns); the build kernels take the inverse (colmap[doc]).  Everything inside the table kernels is in columns;
    // doc ids appear only in the order keys and the outputs.
    const u32* docmap;
};

// exception list (first entry, count) and table row of a clause row id
__device__ __forceinline__ void row_exceptions(const Tab4Dev& T, u32 row, u32& a, u32& n) {
    if (row >= T.temp_row0) { const u32 g = row - T.temp_row0, o = T.gap_off[g]; a = T.tmp_base + o; n = min(T.gap_off[g + 1] - o, (u32)GAP_LIST); }
    else { a = T.exc_off[row]; n = T.exc_off[row + 1] - a; }
}
__device__ __forceinline__ u32 row_data(const Tab4Dev& T, u32 row) { return row >= T.temp_row0 ? T.zero_row : row; }

struct ExcBuild {           // append buffer used while the tables are built
    u32* row; u32* doc; float* freq;
    unsigned long long* count;
    u64 cap;
};

__device__ __forceinline__ void exc_append(const ExcBuild& ex, u32 row, u32 doc, float f) {
    unsigned long long i = atomicAdd(ex.count, 1ull);
    if (i < ex.cap) { ex.row[i] = row; ex.doc[i] = doc; ex.freq[i] = f; }
}

// All 32 lanes call this.  Lanes whose docs share a u32 (8 docs) combine their nibbles and
// the first of them issues one atomicOr (rows are zero-initialised; a row is written by one
// warp per segment, segments run one after another on one stream).
__device__ __forceinline__ void nibble_or(u32* __restrict__ row32, u32 doc, u32 code, bool has, int lane) {
    u32 wi = has ? (doc >> 3) : (0x80000000u | (u32)lane);
    u32 grp = __match_any_sync(0xFFFFFFFFu, wi);
    u32 v = has ? (code << (4 * (doc & 7u))) : 0u;
    u32 orv = __reduce_or_sync(grp, v);
    if (has && lane == __ffs(grp) - 1) atomicOr(&row32[wi], orv);
}

// warp per term of the segment
__global__ void __launch_bounds__(256) tf_table4_kernel(SegDev s, const u32* __restrict__ colmap, u32* __restrict__ tab32, u64 row_words, ExcBuild ex) {
    const int lane = threadIdx.x & 31;
    const u32 t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (t >= s.n_terms) return;
    const u64 key = s.term_key[t];
    const u32 a = s.term_off[t], b = s.term_off[t + 1];
    u32* row = tab32 + key * row_words;
    for (u32 i0 = a; i0 < b; i0 += 32) {
        const u32 i = i0 + lane;
        bool has = i < b;
        u32 doc = 0, code = 0;
        if (has) {
            doc = colmap[s.doc_base + s.post_doc[i]];
            u32 tf = s.post_pos[i + 1] - s.post_pos[i];
            if (tf <= T4_TF_MAX) code = tf;
            else { exc_append(ex, (u32)key, doc, (float)tf); has = false; }
        }
        nibble_or(row, doc, code, has, lane);
    }
}

// warp per (k+1)-mer row u
__global__ void __launch_bounds__(256) pair_table4_kernel(SegDev s, const u32* __restrict__ colmap, int k, int min_strand, u32* __restrict__ tab32, u64 row_words,
                                                           u64 pair_row0, ExcBuild ex, u64 n_rows) {
    const int lane = threadIdx.x & 31;
    const u64 u = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (u >= n_rows) return;
    const u64 kmask = (1ull << (2 * k)) - 1ull;
    const u64 t0 = canon_kmer(u >> 2, k, min_strand), t1 = canon_kmer(u & kmask, k, min_strand);
    const u32 lt0 = s.direct[t0];
    if (lt0 == NO_TERM) return;
    const u32 lt1 = s.direct[t1];
    if (lt1 == NO_TERM) return;
    const bool same = t0 == t1;
    const u32 a0 = s.term_off[lt0], b0 = s.term_off[lt0 + 1];
    const u32 a1 = s.term_off[lt1], b1 = s.term_off[lt1 + 1];
    u32* row = tab32 + (pair_row0 + u) * row_words;
    for (u32 i0 = a0; i0 < b0; i0 += 32) {
        const u32 i = i0 + lane;
        bool has = false;
        u32 doc = 0, code = 0;
        if (i < b0) {
            const float f = phrase_freq_at(s, i, a1, b1, a1 + (i - a0), 2, same);
            if (f != 0.0f) {
                doc = colmap[s.doc_base + s.post_doc[i]];
                const u32 c = (u32)f;
                if ((float)c == f && c >= 1u && c <= T4_PAIR_MAX) { code = c; has = true; }
                else exc_append(ex, (u32)(pair_row0 + u), doc, f);
            }
        }
        nibble_or(row, doc, code, has, lane);
    }
}

// Same rows, staged: warp per first k-mer t0 (segments of at most 256 docs).  The four (k+1)-mers t0+A/C/G/T share
// t0's postings, so its slice of the segment (docs, position offsets, positions) is copied to shared memory once (one
// batch of asynchronous 4-byte copies); each second k-mer's slice follows, a doc -> posting byte map replaces the
// binary search and the literal sweep runs out of shared memory.  Slices beyond the position cap and same-term rows take the direct path.
#define PT_WARPS 8
__host__ __device__ inline size_t pair_stage_warp_bytes(int dcap, int pcap) {
    return 2 * ((size_t)dcap * 4 + (size_t)(dcap + 4) * 4 + (size_t)pcap * 4) + (size_t)dcap;
}
// all copies of one term's slice (docs, position offsets, positions) as asynchronous 4-byte copies; the caller waits
__device__ __forceinline__ void stage_term_async(const SegDev& s, u32 a, u32 n, u32 p, u32 m, u32* __restrict__ w_doc,
                                                 u32* __restrict__ w_pp, u32* __restrict__ w_pos, int lane) {
    for (u32 i = lane; i <= n; i += 32) { if (i < n) cp_async4(w_doc + i, s.post_doc + a + i); cp_async4(w_pp + i, s.post_pos + a + i); }
    for (u32 i = lane; i < m; i += 32) cp_async4(w_pos + i, s.positions + p + i);
}
__global__ void __launch_bounds__(PT_WARPS * 32) pair_table4_staged_kernel(SegDev s, const u32* __restrict__ colmap, int k, int min_strand, u32* __restrict__ tab32,
                                                                            u64 row_words, u64 pair_row0, ExcBuild ex, u64 n_kmers,
                                                                            int dcap, int pcap) {
    extern __shared__ __align__(16) u8 pt_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u64 t0raw = (u64)blockIdx.x * PT_WARPS + warp;
    if (t0raw >= n_kmers) return;
    u8* wbase = pt_raw + (size_t)warp * pair_stage_warp_bytes(dcap, pcap);
    u32* w_pos0 = reinterpret_cast<u32*>(wbase);
    u32* w_pos1 = w_pos0 + pcap;
    u32* w_doc0 = w_pos1 + pcap;
    u32* w_doc1 = w_doc0 + dcap;
    u32* w_pp0 = w_doc1 + dcap;
    u32* w_pp1 = w_pp0 + dcap + 4;
    u8* inv = reinterpret_cast<u8*>(w_pp1 + dcap + 4);
    const u64 kmask = (1ull << (2 * k)) - 1ull;
    const u64 t0 = canon_kmer(t0raw, k, min_strand);
    const u32 lt0 = s.direct[t0];
    if (lt0 == NO_TERM) return;
    // lanes 0..3 look up the four second k-mers (term, posting slice, position range), lane 4 the first one's range
    u32 my_a = 0, my_n = 0, my_p = 0, my_m = 0;
    u64 my_t = 0;
    if (lane < 4) {
        my_t = canon_kmer(((t0raw << 2) | (u64)lane) & kmask, k, min_strand);
        const u32 lt1 = s.direct[my_t];
        if (lt1 != NO_TERM) { my_a = s.term_off[lt1]; my_n = s.term_off[lt1 + 1] - my_a; }
    } else if (lane == 4) { my_a = s.term_off[lt0]; my_n = s.term_off[lt0 + 1] - my_a; }
    if (lane < 5 && my_n) { my_p = s.post_pos[my_a]; my_m = s.post_pos[my_a + my_n] - my_p; }
    const u32 a0 = __shfl_sync(0xFFFFFFFFu, my_a, 4), n0 = __shfl_sync(0xFFFFFFFFu, my_n, 4);
    const u32 p0 = __shfl_sync(0xFFFFFFFFu, my_p, 4), m0 = __shfl_sync(0xFFFFFFFFu, my_m, 4);
    for (u32 d = lane; d < (u32)dcap; d += 32) inv[d] = 0;
    const bool fit0 = m0 <= (u32)pcap;
    if (fit0) stage_term_async(s, a0, n0, p0, m0, w_doc0, w_pp0, w_pos0, lane);     // waited for together with the first t1
#pragma unroll 1
    for (u32 c = 0; c < 4; c++) {
        const u64 u = (t0raw << 2) | c;
        const u64 t1 = __shfl_sync(0xFFFFFFFFu, my_t, (int)c);
        const u32 a1 = __shfl_sync(0xFFFFFFFFu, my_a, (int)c), n1 = __shfl_sync(0xFFFFFFFFu, my_n, (int)c);
        const u32 p1 = __shfl_sync(0xFFFFFFFFu, my_p, (int)c), m1 = __shfl_sync(0xFFFFFFFFu, my_m, (int)c);
        if (n1 == 0) continue;
        const bool same = t0 == t1;
        u32* row = tab32 + (pair_row0 + u) * row_words;
        const bool staged = fit0 && !same && m1 <= (u32)pcap;
        if (staged) stage_term_async(s, a1, n1, p1, m1, w_doc1, w_pp1, w_pos1, lane);
        cp_async_commit_wait_all();
        __syncwarp();
        if (staged) {
            for (u32 i = lane; i < n1; i += 32) inv[w_doc1[i]] = (u8)i;
            __syncwarp();
        }
        for (u32 i0 = 0; i0 < n0; i0 += 32) {
            const u32 i = i0 + lane;
            bool has = false;
            u32 doc = 0, code = 0;
            if (i < n0) {
                float f = 0.0f;
                if (staged) {
                    const u32 d = w_doc0[i], j = inv[d];                  // verified: may be a stale entry
                    if (j < n1 && w_doc1[j] == d) {
                        const u32 q0 = w_pp0[i], q1 = w_pp1[j];
                        f = sloppy_freq_distinct(w_pos0 + (q0 - p0), (int)(w_pp0[i + 1] - q0), w_pos1 + (q1 - p1), (int)(w_pp1[j + 1] - q1), 2);
                    }
                } else {
                    f = phrase_freq_at(s, a0 + i, a1, a1 + n1, a1 + i, 2, same);
                }
                if (f != 0.0f) {
                    doc = colmap[s.doc_base + (staged ? w_doc0[i] : s.post_doc[a0 + i])];
                    const u32 cc = (u32)f;
                    if ((float)cc == f && cc >= 1u && cc <= T4_PAIR_MAX) { code = cc; has = true; }
                    else exc_append(ex, (u32)(pair_row0 + u), doc, f);
                }
            }
            nibble_or(row, doc, code, has, lane);
        }
        __syncwarp();
    }
    cp_async_commit_wait_all();            // (nothing pending unless all four rows were empty)
}

// ---- exception list finalisation: sort by (row, doc), then per-row offsets
__global__ void __launch_bounds__(256) exc_keys_kernel(const u32* __restrict__ row, const u32* __restrict__ doc, u32 n,
                                                        u64* __restrict__ keys, u32* __restrict__ vals) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    keys[i] = ((u64)row[i] << 32) | (u64)doc[i];
    vals[i] = i;
}
__global__ void __launch_bounds__(256) exc_gather_kernel(const u64* __restrict__ keys, const u32* __restrict__ vals,
                                                          const float* __restrict__ freq_in, u32 n, u32* __restrict__ doc_out,
                                                          float* __restrict__ freq_out) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    doc_out[i] = (u32)(keys[i] & 0xFFFFFFFFull);
    freq_out[i] = freq_in[vals[i]];
}
// exc_off[r] = first sorted entry whose row >= r, for r in [0, n_rows]
__global__ void __launch_bounds__(256) exc_offsets_kernel(const u64* __restrict__ keys, u32 n, u64 n_rows, u32* __restrict__ exc_off) {
    u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n_rows) return;
    u32 lo = 0, hi = n;
    while (lo < hi) {
        u32 mid = (lo + hi) >> 1;
        if ((keys[mid] >> 32) < r) lo = mid + 1; else hi = mid;
    }
    exc_off[r] = lo;
}

// ------------------------------------------------------------------ shared helpers
__device__ __forceinline__ u64 warp_max_u64(u64 v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        u64 o = __shfl_xor_sync(0xFFFFFFFFu, v, d);
        if (o > v) v = o;
    }
    return v;
}

// PRMT with a register selector: byte i of the result = byte (c >> 4i) & 7 of {a, b}.  Bit 3 of a selector
// nibble would switch to sign replication; pair rows only hold codes 0..7 (the __byte_perm intrinsic masks the
// selector with an extra LOP, which this avoids).
__device__ __forceinline__ u32 prmt(u32 a, u32 b, u32 c) {
    u32 d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// build-time variants of the filter kernel
#ifndef FK_MAXT
#define FK_MAXT 512                // largest CTA of the filter kernel
#endif
#ifndef FK_MINB
#define FK_MINB 2                  // 64 registers: 5 CTAs of 192 threads per SM (80 regs / 4 CTAs: 42.2 ms, 56 regs / 6 CTAs: 38.7 ms)
#endif

// one u32 of a pair row (8 docs) through the clause's 8-entry LUT into the u16 weight lanes and u8 count lanes
__device__ __forceinline__ void pair_word(u32 word, const uint2 lut, u32& w0, u32& w1, u32& w2, u32& w3, u32& c0, u32& c1) {
    const u32 hi = word >> 16;
    const u32 m0 = prmt(lut.x, lut.y, word);
    const u32 m1 = prmt(lut.x, lut.y, hi);
    w0 += prmt(m0, 0u, 0x4140u);
    w1 += prmt(m0, 0u, 0x4342u);
    w2 += prmt(m1, 0u, 0x4140u);
    w3 += prmt(m1, 0u, 0x4342u);
    c0 += prmt(0x01010100u, 0x01010101u, word);
    c1 += prmt(0x01010100u, 0x01010101u, hi);
}

// cp.async (LDGSTS) 16-byte copy global -> shared, bypassing L1; one commit group per table row
// (cp_async16 lives in classify.cuh)
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// shared-window forms of the ring accesses (32-bit shared addresses: no generic-to-shared conversion in the loop)
__device__ __forceinline__ void cp_async16_sa(u32 sa, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ uint4 lds128(u32 sa) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(sa) : "memory");
    return v;
}

// acc += x on the FMA pipe (IMAD x*1+acc): PRMT/SHF/IADD3 all issue on the ALU pipe, which runs at half the
// warp-issue rate, so the lane adds are moved to the otherwise idle pipe
// (`one` is a kernel argument equal to 1: with a literal 1 ptxas folds the multiply back into IADD3)
__device__ __forceinline__ void add_fma_pipe(u32& acc, u32 x, u32 one) {
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc) : "r"(x), "r"(one));
}
// Two rows at once, weights kept in a packed form that needs one PRMT per four docs instead of two: with p = the four
// LUT bytes (b3 b2 b1 b0) of docs 4j..4j+3 read as one integer,
//   x += p                      -> B0 + 2^8 B1 + 2^16 B2 + 2^24 B3  (mod 2^32),  Bi = the doc's running sum
//   y += (b3 << 16) | b1        -> (B3 << 16) | B1                  (exact while the sums stay below 2^16)
// and after the loop x - (y << 8) = (B2 << 16) | B0 (mod 2^32, exact for the same reason): unpack_xy turns the pair
// back into the u16 lanes of the single-row path.
// (cnt_lo = 0x01010100 arrives in a register computed from a kernel argument: as a literal, ptxas re-materialises it
// with a move in front of every count PRMT)
__device__ __forceinline__ void pair_word2(u32 wa, u32 wb, const uint2 la, const uint2 lb, u32& x0, u32& y0, u32& x1, u32& y1,
                                           u32& c0, u32& c1, u32 one, u32 cnt_lo) {
    const u32 ha = wa >> 16, hb = wb >> 16;
    const u32 a0 = prmt(la.x, la.y, wa), a1 = prmt(la.x, la.y, ha);
    const u32 b0 = prmt(lb.x, lb.y, wb), b1 = prmt(lb.x, lb.y, hb);
    add_fma_pipe(x0, a0, one); add_fma_pipe(x0, b0, one);
    add_fma_pipe(y0, prmt(a0, 0u, 0x4341u), one); add_fma_pipe(y0, prmt(b0, 0u, 0x4341u), one);
    add_fma_pipe(x1, a1, one); add_fma_pipe(x1, b1, one);
    add_fma_pipe(y1, prmt(a1, 0u, 0x4341u), one); add_fma_pipe(y1, prmt(b1, 0u, 0x4341u), one);
    add_fma_pipe(c0, prmt(cnt_lo, 0x01010101u, wa), one); add_fma_pipe(c0, prmt(cnt_lo, 0x01010101u, wb), one);
    add_fma_pipe(c1, prmt(cnt_lo, 0x01010101u, ha), one); add_fma_pipe(c1, prmt(cnt_lo, 0x01010101u, hb), one);
}
// packed (x, y) of docs 4j..4j+3 -> u16 lanes (doc 4j, 4j+1) and (doc 4j+2, 4j+3)
__device__ __forceinline__ void unpack_xy(u32 x, u32 y, u32& w0, u32& w1) {
    const u32 e = x - (y << 8);                        // (B2 << 16) | B0
    w0 = prmt(e, y, 0x5410u);                          // (B1 << 16) | B0
    w1 = prmt(e, y, 0x7632u);                          // (B3 << 16) | B2
}

// Term rows hold codes 0..14: a 16-entry LUT out of two 8-entry PRMT lookups.  PRMT reads bit 3 of a selector nibble as
// "replicate the sign (bit 7) of the selected byte".  Weights are kept <= 127, so a lookup with the raw code yields
// LUT_lo[code] for codes < 8 and 0x00 for codes >= 8, and with bit 3 of the code flipped LUT_hi[code - 8] for codes >= 8
// and 0x00 below: the two results are disjoint and simply add (on the FMA pipe).
__device__ __forceinline__ u32 lut16(const uint4 t, u32 sel /* 4 codes in the low 16 bits */, u32 one) {
    u32 r = prmt(t.x, t.y, sel);
    add_fma_pipe(r, prmt(t.z, t.w, sel ^ 0x8888u), one);
    return r;
}
// 1 per non-zero code.  Looked up in the bytes {0x80, 0x81 x 7}: code 0 -> 0x80, 1..7 -> 0x81, 8 -> sign of 0x80 = 0xFF,
// 9..15 -> sign of 0x81 = 0xFF: bit 0 is the flag.
__device__ __forceinline__ u32 cnt16(u32 sel) { return prmt(0x81818180u, 0x81818181u, sel) & 0x01010101u; }
// one u32 (8 docs) of each of two term rows, into the packed (x, y) lanes of pair_word2
__device__ __forceinline__ void term_word2(u32 wa, u32 wb, const uint4 ta, const uint4 tb, u32& x0, u32& y0, u32& x1, u32& y1,
                                           u32& c0, u32& c1, u32 one) {
    const u32 ha = wa >> 16, hb = wb >> 16;
    u32 s0 = lut16(ta, wa, one), s1 = lut16(ta, ha, one);
    add_fma_pipe(s0, lut16(tb, wb, one), one);        // weights <= 127: the two rows' bytes add without carry
    add_fma_pipe(s1, lut16(tb, hb, one), one);
    add_fma_pipe(x0, s0, one);
    add_fma_pipe(y0, prmt(s0, 0u, 0x4341u), one);
    add_fma_pipe(x1, s1, one);
    add_fma_pipe(y1, prmt(s1, 0u, 0x4341u), one);
    add_fma_pipe(c0, cnt16(wa), one); add_fma_pipe(c0, cnt16(wb), one);
    add_fma_pipe(c1, cnt16(ha), one); add_fma_pipe(c1, cnt16(hb), one);
}

__device__ __forceinline__ uint4 ldg_stream16(const void* p) {     // read-only path, do not keep in L1
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint2 ldg_stream8(const void* p) {
    uint2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}

__device__ __forceinline__ u32 cell_code(const Tab4Dev& T, u32 row, u32 doc) {
    const u32 byte = __ldg(T.tab + (u64)row * T.row_bytes + (doc >> 1));
    return (byte >> (4u * (doc & 1u))) & 15u;
}

// ------------------------------------------------------------------ fast path
#define FK_DOCS 32                 // docs per thread: one 16-byte vector per row
#define FK_MAX_CLAUSES 255         // u8 match counters
#define FK_CAND_CAP 128
#ifndef FK_STAGES
#define FK_STAGES 8                // rows in flight per thread (cp.async ring in shared memory); even
#endif
#define FK_EXS_ROW 4                // exception lists of up to this many cells are copied to shared memory by the set-up
#define FK_EXS_CAP 96              // ... this many cells per CTA in all (the rest is searched in global memory)
#define FK_EXA_SHARED 0xFFFFFFFFu    // s_exa of a clause whose list sits in shared memory
#define FK_QMAX 255.0f
#define FK_BM25_CLASSES 32             // norm classes of a handle the BM25 filter takes (8 classes per factor 4 of document length: a 256 x range)
#define FK_BM25_LUT_BYTES (40 * 1024)   // BM25: per-class clause LUTs behind the ring (classes x (clauses + 1) x 16 B, NAIVE_KMER 24 B)
#define FK_ROUND_SLACK 0.51f       // |q - S*tf*value| <= 0.5 (+ float rounding of the product)
#define FK_REL_SLACK 4e-6f         // float roundings of Lucene's score and of the bound arithmetic

struct FilterStats { unsigned long long candidates, forced, overflow_reads; };

// 10th largest of the per-thread values (bit patterns of non-negative floats; 0 when fewer than 10 are
// non-zero).  Equal values may be removed together, which can only lower the result (a valid, weaker bound).
__device__ __forceinline__ u32 warp_max_u32(u32 v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = max(v, __shfl_xor_sync(0xFFFFFFFFu, v, d));
    return v;
}
__device__ __forceinline__ u32 block_tenth_largest_u32(u32 v, u32* sm /* [16*TOPN + 1] */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (int)(blockDim.x >> 5);
    u32 mine = v;
#pragma unroll 1
    for (int r = 0; r < TOPN; r++) {
        const u32 mx = warp_max_u32(mine);
        if (lane == 0) sm[warp * TOPN + r] = mx;
        if (mine == mx) mine = 0;
    }
    __syncthreads();
    if (warp == 0) {
        u32 loc[5];                                   // nw <= 16 warps -> <= 160 values
#pragma unroll
        for (int j = 0; j < 5; j++) { int idx = lane + 32 * j; loc[j] = idx < nw * TOPN ? sm[idx] : 0u; }
        u32 res = 0;
#pragma unroll 1
        for (int r = 0; r < TOPN; r++) {
            u32 lm = max(max(max(loc[0], loc[1]), max(loc[2], loc[3])), loc[4]);
            res = warp_max_u32(lm);
#pragma unroll
            for (int j = 0; j < 5; j++) if (loc[j] == res) loc[j] = 0;
        }
        if (lane == 0) sm[16 * TOPN] = res;
    }
    __syncthreads();
    return sm[16 * TOPN];
}

// E(d) = Q*m*norm_d for docs with m >= need (else 0): the score estimate both bounds derive from
__device__ __forceinline__ float estimate(u32 Q, u32 m, i32 need, float nrm) {
    const u32 t = (i32)m >= need ? Q * m : 0u;        // <= 65535 * 255 < 2^24: exact in float
    return (float)t * nrm;
}

// exception entry of (row c, doc d): binary search in the row's sorted doc list; 0 if the cell is a true zero
__device__ __forceinline__ float exc_lookup(const Tab4Dev& T, u32 a, u32 n, u32 d) {
    u32 lo = 0, hi = n;
    while (lo < hi) {
        const u32 mid = (lo + hi) >> 1;
        if (__ldg(T.exc_doc + a + mid) < d) lo = mid + 1; else hi = mid;
    }
    return (lo < n && __ldg(T.exc_doc + a + lo) == d) ? __ldg(T.exc_freq + a + lo) : 0.0f;
}

// BM25 (Configuration.java:169-179): score(d) = sum_c w_c (k1+1) f / (f + K_d), K_d = cache[norm byte of d]; no coord.  The
// cell weight depends on the document through K_d, so the byte LUT of a clause exists once per norm class; table columns
// are sorted by norm byte (Tab4Dev::docmap), a thread's 32 columns share a class except at the class boundaries.  A thread
// takes the LUT of the smallest K among its columns: exact for the columns of that class, an over-estimate for the others
// (larger K, smaller true weights) -- those may become candidates (rescoring is exact) but never set the threshold.
struct Bm25Classes {
    const float* class_k;       // [n_classes] K of each norm class present in the handle, in column order
    const u8* group_class;      // [d_pad / 32] bits 0-6: class with the smallest K among the group's columns; bit 7: the group
                                //              holds columns of more than one class (or padding columns)
    int n_classes;
    int nc_cap;                 // clause capacity of one class's LUT array (>= clauses of every read of the launch + 1)
    float k_min, k_max, k1p1;
};

// TILED: several doc tiles per read (more than 16,384 docs, or a capped CTA size); ALLTERMS: NAIVE_KMER (term rows only)
template <bool TILED, bool ALLTERMS, int SIM>
__global__ void __launch_bounds__(FK_MAXT, FK_MINB) filter_score_kernel(
        Tab4Dev T, u32 n_docs, const float* __restrict__ tf16, Bm25Classes bm, const i32* __restrict__ read_list,
        const i64* __restrict__ tok_off, const i32* __restrict__ n_tok, const float* __restrict__ clause_val,
        const u32* __restrict__ clause_row, const i32* __restrict__ n_clauses, const i32* __restrict__ msm,
        const float2* __restrict__ read_vrange, const i32* __restrict__ read_ngap, const float* __restrict__ doc_w, int alg, u32 global_doc_base, int parts_stride,
        i32* __restrict__ part_n, u32* __restrict__ part_doc, float* __restrict__ part_score, i32* __restrict__ route,
        i32* __restrict__ overflow_list, i32* __restrict__ overflow_count, FilterStats* __restrict__ stats, int full_topn,
        u32 one /* = 1, see add_fma_pipe */, int n_tiles /* doc tiles of blockDim.x * 32 docs; CTA per (tile, read) */) {
    // per clause: q for codes 0..3 | 4..7, and in .z the table offset of the row FK_STAGES clauses further on (the row the
    // streaming loop prefetches while it works on this one): one 16-byte shared load per row
    __shared__ __align__(16) uint4 s_lo[FK_MAX_CLAUSES + 1];
    __shared__ uint2 s_lut_hi[ALLTERMS ? FK_MAX_CLAUSES + 1 : 1];       // term rows (NAIVE_KMER): q for codes 8..11 | 12..15
    __shared__ u32 s_row[FK_MAX_CLAUSES + 1];
    __shared__ u32 s_off[FK_MAX_CLAUSES + 1];           // row offset in the table, in 16-byte units (table < 64 GiB)
    __shared__ float s_val[FK_MAX_CLAUSES + 1];
    __shared__ u32 s_exa[FK_MAX_CLAUSES + 1], s_exn[FK_MAX_CLAUSES + 1];     // exception list of each clause's row
    __shared__ u8 s_exrows[FK_MAX_CLAUSES + 1];          // the clauses whose row has a long exception list (searched per thread)
    __shared__ u32 s_sm_doc[FK_EXS_CAP];                 // cells of the short lists: doc, frequency, clause
    __shared__ float s_sm_f[FK_EXS_CAP];
    __shared__ u8 s_sm_c[FK_EXS_CAP];
    __shared__ u32 s_nsmall;
    __shared__ float s_tf[16];
    __shared__ u32 s_sel[16 * TOPN + 1];
    __shared__ u32 s_wmax[16];
    __shared__ u32 s_cand[FK_CAND_CAP];
    __shared__ u64 s_ckey[FK_CAND_CAP];
    __shared__ int s_ncand, s_nexrows, s_over, s_ndense, s_ngap;
    __shared__ u32 s_qmin, s_nexc, s_cntlo;

    __shared__ float s_ratio[SIM == SIM_BM25 ? FK_BM25_CLASSES * 16 : 1];      // BM25: f / (f + K_class) per (class, code)
    extern __shared__ __align__(16) uint4 ring[];         // FK_STAGES x blockDim.x x 16 B: row ring of pass 1, then the staged cell codes
    // BM25: behind the ring, per class the clause LUTs (uint4: q codes 0..3 | 4..7 | prefetch offset), then for NAIVE_KMER the
    // high halves (uint2: codes 8..11 | 12..15)
    uint4* const lut_cls = ring + (size_t)FK_STAGES * blockDim.x;
    uint2* const lut_cls_hi = reinterpret_cast<uint2*>(lut_cls + (size_t)bm.n_classes * bm.nc_cap);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (int)(blockDim.x >> 5);
    const int tile = TILED ? blockIdx.x % n_tiles : 0;
    // (read_list == nullptr: every read of the batch takes this kernel, in order -- one dependent load less at the start)
    const i32 rslot = (i32)(TILED ? blockIdx.x / n_tiles : blockIdx.x);
    const i32 r = read_list ? read_list[rslot] : rslot;
    const i64 base = tok_off[r];
    const i32 n = n_tok[r], nc = n_clauses[r];
    const i32 n_pair = alg == ALG_NAIVE ? 0 : (alg == ALG_CHAIN ? nc : n / 2);
    const i32 need = max(1, msm[r]);
    const i32 n_gap = read_ngap[r];                    // gap clauses of the read (all of them pair clauses)
    // scale S: S * value_c * tf(7) <= 255 for every clause; q_min = smallest non-zero cell weight (weights_kernel
    // recorded the largest and smallest clause value of the read)
    const float2 vr = read_vrange[r];
    // (NAIVE_KMER has term rows only: all 15 codes go through the byte LUT, weights <= 127)
    const bool all_terms = ALLTERMS;
    // BM25: the largest cell weight is w_max (k1+1) f_max / (f_max + K_min), the smallest w_min (k1+1) / (1 + K_max)
    const float cmax = all_terms ? (float)T4_TF_MAX : (float)T4_PAIR_MAX;
    const float S = SIM == SIM_BM25 ? (all_terms ? 127.0f : FK_QMAX) / (vr.x * bm.k1p1 * (cmax / (cmax + bm.k_min)))
                                    : (all_terms ? 127.0f / (vr.x * tf16[T4_TF_MAX]) : FK_QMAX / (vr.x * tf16[T4_PAIR_MAX]));
    if (threadIdx.x < 16) s_tf[threadIdx.x] = tf16[threadIdx.x];
    if (SIM == SIM_BM25) {
        for (int i = threadIdx.x; i < bm.n_classes * 16; i += blockDim.x) {
            const float f = (float)(i & 15);
            s_ratio[i] = f / (f + bm.class_k[i >> 4]);
        }
    }
    if (threadIdx.x == 0) {
        s_ncand = 0; s_nexrows = 0; s_over = 0; s_nexc = 0; s_ndense = 0; s_ngap = 0; s_nsmall = 0;
        s_qmin = SIM == SIM_BM25 ? (u32)(S * vr.y * bm.k1p1 * (1.0f / (1.0f + bm.k_max)) * 0.9999f) : (u32)__float2int_rn(S * vr.y);
        s_cntlo = 0x01010100u * one;
    }
    __syncthreads();
    // Clause order in shared memory: table-backed pair clauses [0, n_dense), term clauses, then the gap pair clauses
    // (virtual all-zero rows, matches in their exception lists) at the end: rows [0, n_stream) are streamed.  The double
    // sums are exact, so the order is free.
    const i32 n_dense = n_pair - n_gap, n_stream = nc - n_gap;
    for (i32 c0 = threadIdx.x; c0 < nc; c0 += blockDim.x) {
        const float v = clause_val[base + c0];
        const u32 row = clause_row[base + c0];
        i32 c = c0;
        if (n_gap) {
            if (c0 >= n_pair) c = c0 - n_gap;
            else c = row >= T.temp_row0 ? nc - 1 - atomicAdd(&s_ngap, 1) : atomicAdd(&s_ndense, 1);
        }
        s_val[c] = v;
        s_row[c] = row;
        const u32 off16 = row < T.temp_row0 ? row * (u32)(T.row_bytes >> 4) : 0u;
        s_off[c] = off16;
        if (c >= FK_STAGES) {
            s_lo[c - FK_STAGES].z = off16;
            if (SIM == SIM_BM25) for (int cl = 0; cl < bm.n_classes; cl++) lut_cls[(size_t)cl * bm.nc_cap + c - FK_STAGES].z = off16;
        }
        const float sv = SIM == SIM_BM25 ? S * v * bm.k1p1 : S * v;
        if (SIM == SIM_BM25) {
            if (c0 < n_pair || all_terms) {
                for (int cl = 0; cl < bm.n_classes; cl++) {
                    const float* rt = s_ratio + cl * 16;
                    u32 q[16];
                    q[0] = 0; q[15] = 0;
                    const u32 cap = all_terms ? 127u : 255u;
#pragma unroll
                    for (int j = 1; j < 15; j++) q[j] = min((u32)__float2int_rn(sv * rt[j]), cap);
                    uint4* e = lut_cls + (size_t)cl * bm.nc_cap + c;
                    e->x = q[0] | (q[1] << 8) | (q[2] << 16) | (q[3] << 24);
                    e->y = q[4] | (q[5] << 8) | (q[6] << 16) | (q[7] << 24);
                    if (all_terms)
                        lut_cls_hi[(size_t)cl * bm.nc_cap + c] = make_uint2(q[8] | (q[9] << 8) | (q[10] << 16) | (q[11] << 24), q[12] | (q[13] << 8) | (q[14] << 16) | (q[15] << 24));
                }
            }
        } else if (c0 < n_pair) {
            u32 q[8];
            q[0] = 0;
#pragma unroll
            for (int j = 1; j < 8; j++) q[j] = min((u32)__float2int_rn(sv * tf16[j]), 255u);
            s_lo[c].x = q[0] | (q[1] << 8) | (q[2] << 16) | (q[3] << 24);
            s_lo[c].y = q[4] | (q[5] << 8) | (q[6] << 16) | (q[7] << 24);
        } else if (all_terms) {
            u32 q[16];
            q[0] = 0; q[15] = 0;
#pragma unroll
            for (int j = 1; j < 15; j++) q[j] = min((u32)__float2int_rn(sv * tf16[j]), 127u);
            s_lo[c].x = q[0] | (q[1] << 8) | (q[2] << 16) | (q[3] << 24);
            s_lo[c].y = q[4] | (q[5] << 8) | (q[6] << 16) | (q[7] << 24);
            s_lut_hi[c] = make_uint2(q[8] | (q[9] << 8) | (q[10] << 16) | (q[11] << 24), q[12] | (q[13] << 8) | (q[14] << 16) | (q[15] << 24));
        }
    }
    __syncthreads();

    // ---- pass 1: stream the rows.  Thread owns docs [32t, 32t+32): wacc[i] = u16 lanes (doc 2i, 2i+1),
    //      cacc[j] = u8 lanes (docs 4j..4j+3)
    const u32 d0 = ((u32)tile * blockDim.x + threadIdx.x) * FK_DOCS;
    const bool active = (u64)d0 < T.row_bytes * 2;
    u32 wacc[16], cacc[8];
#pragma unroll
    for (int i = 0; i < 16; i++) wacc[i] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) cacc[i] = 0;
    const u8* tbase = T.tab + (u64)d0 / 2;
    asm volatile("" : "+l"(tbase));          // keep the thread's table base in a register pair (one IMAD.WIDE per row address)
    // clause LUTs this thread reads: the read's (tf-idf), or those of the thread's norm class (BM25)
    const uint4* lo_tab = s_lo;
    const uint2* hi_tab = s_lut_hi;
    int cls = 0;
    float k_thr = 0.0f;
    const u32 gcls = active ? (u32)bm.group_class[d0 / FK_DOCS] : 0x80u;
    const bool uniform = !(gcls & 0x80u);       // all 32 columns share one norm class: one per-doc weight for the thread
    if (SIM == SIM_BM25) {
        cls = (int)(gcls & 0x7Fu);
        k_thr = bm.class_k[cls];
        lo_tab = lut_cls + (size_t)cls * bm.nc_cap;
        hi_tab = lut_cls_hi + (size_t)cls * bm.nc_cap;
    }
    // Each thread prefetches its own 16-byte slice of the next FK_STAGES rows into its private slots of the
    // ring (no cross-thread hand-off, so no barriers): FK_STAGES x 16 B in flight per thread.  Slots are
    // addressed in the shared window directly; one commit group per pair of rows.
    uint4* slot = ring + threadIdx.x;
    const u32 stride = blockDim.x;
    const u32 slot_sa = (u32)__cvta_generic_to_shared(slot), stride_b = stride * 16u, ring_end = slot_sa + FK_STAGES * stride_b;
    if (active) {
#pragma unroll
        for (int st = 0; st < FK_STAGES; st++) {
            if (st < n_stream) cp_async16_sa(slot_sa + st * stride_b, tbase + (u64)s_off[st] * 16);
            if (st & 1) cp_async_commit();
        }
    }
    // ---- exception lists of the clauses' rows, looked up while the first rows are on their way (two more dependent
    //      loads that would otherwise sit in front of them); published by the barrier after pass 1
    for (i32 c = threadIdx.x; c < nc; c += blockDim.x) {
        const u32 row = s_row[c];
        u32 a, ne;
        row_exceptions(T, row, a, ne);
        s_exa[c] = a;
        s_exn[c] = ne;
        if (ne) {
            // Nearly every list is one or two cells long: those are fetched here, next to the other loads of the set-up,
            // so that the fold after pass 1 reads them from shared memory instead of every thread chasing them through
            // L2.  Slots a list reserved but could not use (capacity) are marked empty.
            const u32 at = ne <= FK_EXS_ROW ? atomicAdd(&s_nsmall, ne) : (u32)FK_EXS_CAP;
            if (at + ne <= FK_EXS_CAP) {
                s_exa[c] = FK_EXA_SHARED;
                u32 xd[FK_EXS_ROW]; float xf[FK_EXS_ROW];
#pragma unroll
                for (u32 j = 0; j < FK_EXS_ROW; j++) {
                    xd[j] = j < ne ? __ldg(T.exc_doc + a + j) : 0u;
                    xf[j] = j < ne ? __ldg(T.exc_freq + a + j) : 0.0f;
                }
#pragma unroll
                for (u32 j = 0; j < FK_EXS_ROW; j++)
                    if (j < ne) { BSP_ASSERT(at + j < (u32)FK_EXS_CAP); s_sm_doc[at + j] = xd[j]; s_sm_f[at + j] = xf[j]; s_sm_c[at + j] = (u8)c; }
            } else {
                for (u32 j = at; j < FK_EXS_CAP; j++) s_sm_doc[j] = 0xFFFFFFFFu;
                const int xr = atomicAdd(&s_nexrows, 1);
                BSP_ASSERT(xr >= 0 && xr <= FK_MAX_CLAUSES && c <= FK_MAX_CLAUSES);
                s_exrows[xr] = (u8)c;
            }
        }
    }
    if (active) {
        i32 i = 0;
        const i32 n_two = all_terms ? nc : n_dense;   // rows streamed two at a time through a byte LUT
        // count LUT low word, read from shared memory so that it lives in a vector register: as a literal (or a
        // uniform value) ptxas re-materialises it with a move in front of every count PRMT
        const u32 cnt_lo = s_cntlo;
        // Two rows per step: wait for them, pull them into registers, the LUT arithmetic into the packed (x, y) lanes
        // (pair_word2), then refill their slots with the rows FK_STAGES ahead.  (Four or eight rows per iteration
        // measured slower at C2: the larger bodies lose more to instruction fetch than they save in address arithmetic.)
        u32 xacc[8], yacc[8];
#pragma unroll
        for (int j = 0; j < 8; j++) { xacc[j] = 0; yacc[j] = 0; }
        u32 cur = slot_sa;                            // slot of row i; rows i and i + 1 sit one stage apart
#pragma unroll 1
        for (; i + 2 <= n_two; i += 2) {
            cp_async_wait<FK_STAGES / 2 - 1>();
            const u32 sa0 = cur, sa1 = cur + stride_b;
            BSP_ASSERT(sa0 >= slot_sa && sa1 < ring_end && i + 1 < nc && nc <= FK_MAX_CLAUSES);       // both slots inside this thread's ring
            const uint4 q0 = lds128(sa0), q1 = lds128(sa1);
            const uint4 l0 = lo_tab[i], l1 = lo_tab[i + 1];
            if (all_terms) {
                const uint2 h0 = hi_tab[i], h1 = hi_tab[i + 1];
                const uint4 t0 = make_uint4(l0.x, l0.y, h0.x, h0.y), t1 = make_uint4(l1.x, l1.y, h1.x, h1.y);
                term_word2(q0.x, q1.x, t0, t1, xacc[0], yacc[0], xacc[1], yacc[1], cacc[0], cacc[1], one);
                term_word2(q0.y, q1.y, t0, t1, xacc[2], yacc[2], xacc[3], yacc[3], cacc[2], cacc[3], one);
                term_word2(q0.z, q1.z, t0, t1, xacc[4], yacc[4], xacc[5], yacc[5], cacc[4], cacc[5], one);
                term_word2(q0.w, q1.w, t0, t1, xacc[6], yacc[6], xacc[7], yacc[7], cacc[6], cacc[7], one);
            } else {
                const uint2 p0 = make_uint2(l0.x, l0.y), p1 = make_uint2(l1.x, l1.y);
                pair_word2(q0.x, q1.x, p0, p1, xacc[0], yacc[0], xacc[1], yacc[1], cacc[0], cacc[1], one, cnt_lo);
                pair_word2(q0.y, q1.y, p0, p1, xacc[2], yacc[2], xacc[3], yacc[3], cacc[2], cacc[3], one, cnt_lo);
                pair_word2(q0.z, q1.z, p0, p1, xacc[4], yacc[4], xacc[5], yacc[5], cacc[4], cacc[5], one, cnt_lo);
                pair_word2(q0.w, q1.w, p0, p1, xacc[6], yacc[6], xacc[7], yacc[7], cacc[6], cacc[7], one, cnt_lo);
            }
            if (i + FK_STAGES < n_stream) cp_async16_sa(sa0, tbase + (u64)l0.z * 16);
            if (i + 1 + FK_STAGES < n_stream) cp_async16_sa(sa1, tbase + (u64)l1.z * 16);
            cp_async_commit();
            cur += 2u * stride_b;
            if (cur == ring_end) cur = slot_sa;
        }
#pragma unroll
        for (int j = 0; j < 8; j++) unpack_xy(xacc[j], yacc[j], wacc[2 * j], wacc[2 * j + 1]);
        // single rows (an odd pair row, Term clauses of a mixed read).  Their waits are stricter than needed (the
        // groups before them hold two rows each), never weaker.
#pragma unroll 1
        for (; i < n_stream; i++) {
            cp_async_wait<FK_STAGES / 2 - 1>();
            uint4* my = slot + (i % FK_STAGES) * stride;
            const uint4 q = *my;
            const i32 c = i;
            if (c < n_dense) {
                const uint2 lut = make_uint2(lo_tab[c].x, lo_tab[c].y);
                pair_word(q.x, lut, wacc[0], wacc[1], wacc[2], wacc[3], cacc[0], cacc[1]);
                pair_word(q.y, lut, wacc[4], wacc[5], wacc[6], wacc[7], cacc[2], cacc[3]);
                pair_word(q.z, lut, wacc[8], wacc[9], wacc[10], wacc[11], cacc[4], cacc[5]);
                pair_word(q.w, lut, wacc[12], wacc[13], wacc[14], wacc[15], cacc[6], cacc[7]);
            } else {                                  // Term clause (tf 1..14): nibble at a time
                const float sv = SIM == SIM_BM25 ? S * s_val[c] * bm.k1p1 : S * s_val[c];
                const float* wt = SIM == SIM_BM25 ? s_ratio + cls * 16 : s_tf;
                const u32 w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int x = 0; x < 4; x++) {
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const u32 code = (w[x] >> (4 * j)) & 15u;
                        const u32 qv = (u32)__float2int_rn(sv * wt[code]);
                        wacc[4 * x + (j >> 1)] += qv << (16 * (j & 1));
                        cacc[2 * x + (j >> 2)] += (code ? 1u : 0u) << (8 * (j & 3));
                    }
                }
            }
            // the slot's data now lives in registers: refill it with the row FK_STAGES ahead
            if (i + FK_STAGES < n_stream) cp_async16(my, tbase + (u64)s_off[i + FK_STAGES] * 16);
            cp_async_commit();
        }
    }
    __syncthreads();                                  // exception lists (s_exa, s_exn, s_sm_*, s_exrows) are complete
    if (active) {
        // ---- exception cells (stored as 0 in the table) of this thread's docs join their lanes with the same
        //      quantisation: the short lists from shared memory, the long ones by binary search to the thread's doc range
        u32 exc_sum = 0, exc_cnt = 0;
        auto fold_cell = [&](u32 li, u32 c, float freq) {
            const u32 qe = SIM == SIM_BM25 ? (u32)min(__float2int_rn(S * s_val[c] * bm.k1p1 * (freq / (freq + k_thr))), 1 << 20)
                                           : (u32)min(__float2int_rn(S * s_val[c] * tf_of(freq)), 1 << 20);
            const u32 qv = qe << (16 * (li & 1u)), cnt1 = 1u << (8 * (li & 3u));
#pragma unroll
            for (int k = 0; k < 16; k++) wacc[k] += (u32)k == (li >> 1) ? qv : 0u;     // predicated, statically indexed
#pragma unroll
            for (int k = 0; k < 8; k++) cacc[k] += (u32)k == (li >> 2) ? cnt1 : 0u;
            exc_sum += qe;
            exc_cnt++;
            atomicMin(&s_qmin, qe);
        };
        const int n_small = (int)min(s_nsmall, (u32)FK_EXS_CAP);
#pragma unroll 1
        for (int j = 0; j < n_small; j++) {
            const u32 li = s_sm_doc[j] - d0;
            if (li < (u32)FK_DOCS) fold_cell(li, s_sm_c[j], s_sm_f[j]);
        }
        const int n_exrows = s_nexrows;
#pragma unroll 1
        for (int i = 0; i < n_exrows; i++) {
            const u32 c = s_exrows[i], a = s_exa[c], ne = s_exn[c];
            u32 lo = 0, hi = ne;
            while (lo < hi) {
                const u32 mid = (lo + hi) >> 1;
                if (__ldg(T.exc_doc + a + mid) < d0) lo = mid + 1; else hi = mid;
            }
#pragma unroll 1
            for (; lo < ne; lo++) {
                const u32 d = __ldg(T.exc_doc + a + lo);
                if (d >= d0 + FK_DOCS) break;
                fold_cell(d - d0, c, __ldg(T.exc_freq + a + lo));
            }
        }
        // u16 lanes: pair-row weights <= 255 and term-row weights <= 255*tf(14)/tf(7) < 362 (all-term reads: <= 127 each),
        // plus this thread's exception cells
        // (BM25 term rows: 14 / (14 + K) against 7 / (7 + K) of the pair rows the scale is set for: < 2 x 255)
        const u32 lane_max = all_terms ? 127u * (u32)nc : 255u * (u32)n_pair + (SIM == SIM_BM25 ? 510u : 362u) * (u32)(nc - n_pair);
        if (lane_max + exc_sum > 65535u) s_over = 1;
        if (exc_cnt) atomicAdd(&s_nexc, exc_cnt);
    }
    // ---- pass 2: per-thread best estimate, then the block's 10th best (full list) or best (result set only).
    //      Padded docs have no cells (m = 0).
    float e_best = 0.0f;          // best estimate among the thread's docs that may set the threshold
    float e_any = 0.0f;           // best estimate among all of them (BM25: incl. the over-estimated docs of another class)
    // Only docs with m >= need have an estimate, and a read matches few of them that often: test the 32 count bytes of
    // the thread at once.  byte >= need  <=>  bit 7 of ((byte | 0x80) - need) | byte, for need <= 128 (no borrow
    // crosses a byte: byte | 0x80 >= need).
    bool any_m = active;
    if (active && need <= 128) {
        const u32 need4 = (u32)need * 0x01010101u;
        u32 f = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) f |= ((cacc[j] | 0x80808080u) - need4) | cacc[j];
        any_m = (f & 0x80808080u) != 0u;
    }
    if (any_m && uniform) {
        // one norm class (all but the few groups at class boundaries): the best estimate is the largest integer Q*m (tf-idf)
        // or Q (BM25) times the class weight -- no per-doc weights to load, no per-doc float arithmetic
        u32 tmax = 0;
#pragma unroll
        for (int i = 0; i < FK_DOCS; i++) {
            const u32 Q = prmt(wacc[i >> 1], 0u, (i & 1) ? 0x4432u : 0x4410u);
            const u32 m = prmt(cacc[i >> 2], 0u, 0x4440u | (u32)(i & 3));
            const u32 t = (i32)m >= need ? (SIM == SIM_BM25 ? Q : Q * m) : 0u;
            tmax = max(tmax, t);
        }
        e_best = SIM == SIM_BM25 ? (float)tmax : (float)tmax * bm.class_k[gcls & 0x7Fu];
        e_any = e_best;
    } else if (any_m) {
#pragma unroll
        for (int i4 = 0; i4 < FK_DOCS; i4 += 4) {
            const float4 nv = *reinterpret_cast<const float4*>(doc_w + d0 + i4);
            const float nr[4] = {nv.x, nv.y, nv.z, nv.w};
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int i = i4 + k;
                const u32 Q = prmt(wacc[i >> 1], 0u, (i & 1) ? 0x4432u : 0x4410u);
                const u32 m = prmt(cacc[i >> 2], 0u, 0x4440u | (u32)(i & 3));
                if (SIM == SIM_BM25) {
                    const float e = (i32)m >= need ? (float)Q : 0.0f;          // no coord, K_d is in the weights
                    e_any = fmaxf(e_any, e);
                    if (nr[k] == k_thr) e_best = fmaxf(e_best, e);
                } else e_best = fmaxf(e_best, estimate(Q, m, need, nr[k]));
            }
        }
    }
    if (SIM != SIM_BM25) e_any = e_best;
    // full_topn: 10 docs are known to score >= tau_e*(1-delta); otherwise only the docs that can tie at the maximum
    // matter (Classifier.java:358-366 keeps the hits equal to the top score), so tau_e is the best estimate
    u32 sel;
    if (full_topn) sel = block_tenth_largest_u32(__float_as_uint(e_best), s_sel);
    else {
        const u32 wm = warp_max_u32(__float_as_uint(e_best));
        if (lane == 0) s_wmax[warp] = wm;
        __syncthreads();
        sel = 0;
        for (int w = 0; w < nw; w++) sel = max(sel, s_wmax[w]);
    }
    const float tau_e = __uint_as_float(sel);
    const bool lane_overflow = s_over != 0;              // published by the barrier inside the selection
    // exact score of doc d = E(d)/(S*nc) * (1 +- delta): each weight errs by <= 0.51 and is >= q_min;
    // FK_REL_SLACK covers the float roundings of Lucene's score and of the estimate
    const float qmin = (float)s_qmin;
    const float delta = qmin >= 2.0f ? FK_ROUND_SLACK / qmin + FK_REL_SLACK : 1.0f;
    // a doc can reach the list only if E(d)*(1+delta) >= tau_e*(1-delta)
    const float theta = delta < 1.0f ? tau_e * ((1.0f - delta) / (1.0f + delta)) * (1.0f - FK_REL_SLACK) : 0.0f;
    // ---- candidates: estimate >= theta.  Few threads hold one (1.5 warps per read at C2 see any), and a thread walking
    //      its own 32 docs is a long serial chain that the rest of the CTA waits for at the barrier below (unrolled over
    //      the register lanes it was also 19 KB of rarely executed SASS).  So the warp does it together: a thread with a
    //      candidate parks its lanes in its own ring slots (pass 1 is over in this warp's threads, the slots are private
    //      to them), then the 32 lanes evaluate that thread's 32 docs, one each.
    {
        const bool mine = active && !lane_overflow && e_any >= theta && e_any > 0.0f;
        u32 holders = __ballot_sync(0xFFFFFFFFu, mine);
        if (holders) {
            const u32 stride = blockDim.x;
            if (mine) {
                uint4* my = ring + threadIdx.x;
#pragma unroll
                for (int k = 0; k < 4; k++) my[k * stride] = make_uint4(wacc[4 * k], wacc[4 * k + 1], wacc[4 * k + 2], wacc[4 * k + 3]);
#pragma unroll
                for (int k = 0; k < 2; k++) my[(4 + k) * stride] = make_uint4(cacc[4 * k], cacc[4 * k + 1], cacc[4 * k + 2], cacc[4 * k + 3]);
            }
            __syncwarp();
#pragma unroll 1
            while (holders) {
                const u32 src = (threadIdx.x & ~31u) + (u32)(__ffs((int)holders) - 1);      // the thread whose docs these are
                holders &= holders - 1u;
                const u32* parked = reinterpret_cast<const u32*>(ring + src);              // slot k: + k * stride * 4 words
                const u32 ww = parked[(u32)(lane >> 3) * stride * 4u + ((u32)(lane >> 1) & 3u)];        // u16 lanes (doc 2i, 2i+1) = word i
                const u32 cw = parked[(4u + (u32)(lane >> 4)) * stride * 4u + ((u32)(lane >> 2) & 3u)]; // u8 lanes (docs 4j..4j+3) = word j
                const u32 Q = (lane & 1) ? ww >> 16 : ww & 0xFFFFu;
                const u32 m = (cw >> (8 * (lane & 3))) & 0xFFu;
                const u32 d = ((u32)tile * blockDim.x + src) * FK_DOCS + (u32)lane;
                const float e = SIM == SIM_BM25 ? ((i32)m >= need ? (float)Q : 0.0f) : estimate(Q, m, need, doc_w[d]);
                if (e >= theta && e > 0.0f) {
                    int idx = atomicAdd(&s_ncand, 1);
                    BSP_ASSERT(idx >= 0);
                    if (idx < FK_CAND_CAP) s_cand[idx] = d;
                }
            }
        }
    }
    __syncthreads();
    const int n_cand = s_ncand;
    const i64 pbase = (i64)r * parts_stride + tile;
    if (lane_overflow || n_cand > FK_CAND_CAP) {          // hand the read to exact_score_kernel (once, whichever tile asks)
        if (threadIdx.x == 0) {
            part_n[pbase] = 0;
            if (atomicExch(&route[r], ROUTE_DENSE) != ROUTE_DENSE) {
                overflow_list[atomicAdd(overflow_count, 1)] = r;
                if (stats) atomicAdd(&stats->overflow_reads, 1ull);
            }
        }
        return;
    }
    // ---- exact rescoring.  Few candidates: one warp each, lanes over clauses.  Many (full list): the cells are first
    //      fetched in one batch of independent loads (item = candidate j, clause c) into shared memory.
    const int n_small = (int)min(s_nsmall, (u32)FK_EXS_CAP);
    int pitch_log = 5;                                    // clause pitch of the staging buffer: power of two >= nc
    while ((1 << pitch_log) < nc) pitch_log++;
    const int n_items = n_cand << pitch_log;
    u8* s_code = reinterpret_cast<u8*>(ring);             // pass 1 is over in every warp (barriers above): the ring is free
    const bool staged = n_cand > nw && n_items <= FK_STAGES * (int)blockDim.x * 16;
    if (staged) {
#pragma unroll 1
        for (int it0 = threadIdx.x; it0 < n_items; it0 += 4 * blockDim.x) {
            u32 code[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int it = it0 + u * blockDim.x;
                code[u] = 0;
                const int j = it >> pitch_log, c = it & ((1 << pitch_log) - 1);
                if (it < n_items && c < nc && s_row[c] < T.temp_row0) code[u] = cell_code(T, s_row[c], s_cand[j]);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int it = it0 + u * blockDim.x;
                if (it < n_items) s_code[it] = (u8)code[u];
            }
        }
        __syncthreads();
    }
#pragma unroll 1
    for (int j = warp; j < n_cand; j += nw) {
        const u32 d = s_cand[j];
        const float nrm = doc_w[d];
        double sum = 0.0;
        i32 m = 0;
        // four clauses per lane and round: their cell loads (one DRAM access each) are in flight together
#pragma unroll 1
        for (i32 cb = lane; cb < nc; cb += 128) {
            u32 code[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const i32 c = cb + 32 * u;
                code[u] = 0u;
                if (c < nc) code[u] = staged ? (u32)s_code[(j << pitch_log) + c]
                                             : (s_row[c] >= T.temp_row0 ? 0u : cell_code(T, s_row[c], d));
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const i32 c = cb + 32 * u;
                if (c >= nc) break;
                // the exception cell of (clause, doc), if any (its table cell is stored as 0): a scan of the short lists
                // in shared memory, or a search of the row's list
                float fe = 0.0f;
                if (code[u] == 0u && s_exn[c]) {
                    if (s_exa[c] == FK_EXA_SHARED) {
                        for (int x = 0; x < n_small; x++)
                            if (s_sm_doc[x] == d && s_sm_c[x] == (u8)c) fe = s_sm_f[x];
                    } else fe = exc_lookup(T, s_exa[c], s_exn[c], d);
                }
                float tf = s_tf[code[u]];                                              // BM25: the frequency itself
                if (fe != 0.0f) tf = SIM == SIM_BM25 ? fe : tf_of(fe);
                if (tf != 0.0f) {
                    sum += (double)clause_doc_score(SIM, tf, s_val[c], nrm, bm.k1p1);     // TFIDFSimScorer#score / BM25DocScorer#score
                    m++;
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
            m += __shfl_xor_sync(0xFFFFFFFFu, m, o);
        }
        const u64 key = m >= need ? cand_key(final_score(SIM, sum, m, nc), T.docmap[d]) : 0ull;
        if (n_cand == 1) {                                // the usual case: nothing to sort, nobody to wait for
            if (lane == 0) {
                if (key) {
                    part_doc[pbase * TOPN] = global_doc_base + (0xFFFFFFFFu - (u32)(key & 0xFFFFFFFFu));
                    part_score[pbase * TOPN] = __uint_as_float((u32)(key >> 32));
                }
                part_n[pbase] = key ? 1 : 0;
                if (stats) { atomicAdd(&stats->candidates, 1ull); atomicAdd(&stats->forced, (unsigned long long)s_nexc); }
            }
            return;
        }
        if (lane == 0) s_ckey[j] = key;
    }
    if (n_cand == 1) return;                              // (the warps without the candidate)
    __syncthreads();
    // ---- top-10 among the candidates
    if (warp == 0) {
        u64 loc[FK_CAND_CAP / 32];
#pragma unroll
        for (int j = 0; j < FK_CAND_CAP / 32; j++) { int idx = lane + 32 * j; loc[j] = idx < n_cand ? s_ckey[idx] : 0ull; }
        i32 found = 0;
#pragma unroll 1
        for (int round = 0; round < TOPN; round++) {
            u64 lm = 0;
#pragma unroll
            for (int j = 0; j < FK_CAND_CAP / 32; j++) if (loc[j] > lm) lm = loc[j];
            const u64 best = warp_max_u64(lm);
            if (best == 0) break;
#pragma unroll
            for (int j = 0; j < FK_CAND_CAP / 32; j++) if (loc[j] == best) loc[j] = 0;
            if (lane == 0) {
                part_doc[pbase * TOPN + round] = global_doc_base + (0xFFFFFFFFu - (u32)(best & 0xFFFFFFFFu));
                part_score[pbase * TOPN + round] = __uint_as_float((u32)(best >> 32));
            }
            found++;
        }
        if (lane == 0) {
            part_n[pbase] = found;
            if (stats) { atomicAdd(&stats->candidates, (unsigned long long)n_cand); atomicAdd(&stats->forced, (unsigned long long)s_nexc); }
        }
    }
}

// ------------------------------------------------------------------ exact path
// CTA per (doc tile, read); thread owns 16 consecutive docs (one 8-byte vector per row), 16 double
// accumulators + packed u16 match counts in registers.  Exception cells are applied per clause stage.
#define EX_DPT 16
#define EX_STAGE 256
#define EX_EXC_CAP 256
template <int SIM>
__global__ void __launch_bounds__(512) exact_score_kernel(
        Tab4Dev T, u32 n_docs, const float* __restrict__ tf16, const i32* __restrict__ read_list,
        const i64* __restrict__ tok_off, const i32* __restrict__ n_tok, const float* __restrict__ clause_val,
        const u32* __restrict__ clause_row, const i32* __restrict__ n_clauses, const i32* __restrict__ msm,
        const float* __restrict__ doc_w, float k1p1, u32 global_doc_base, int n_tiles, int parts_stride,
        i32* __restrict__ part_n, u32* __restrict__ part_doc, float* __restrict__ part_score) {
    __shared__ float s_tf[16];
    __shared__ float s_val[EX_STAGE];
    __shared__ u32 s_row[EX_STAGE];
    __shared__ u64 red[33];
    __shared__ u32 s_scan[33];
    __shared__ u32 s_exc_c[EX_EXC_CAP], s_exc_d[EX_EXC_CAP];
    __shared__ float s_exc_f[EX_EXC_CAP];
    const int tile = blockIdx.x % n_tiles;
    const i32 r = read_list[blockIdx.x / n_tiles];
    if (threadIdx.x < 16) s_tf[threadIdx.x] = tf16[threadIdx.x];
    const i64 base = tok_off[r];
    const i32 nc = n_clauses[r];
    const u32 d0 = ((u32)tile * blockDim.x + threadIdx.x) * EX_DPT;
    const bool active = (u64)d0 < T.row_bytes * 2;
    double acc[EX_DPT];
    u32 cnt[EX_DPT / 2];
    float w[EX_DPT];
#pragma unroll
    for (int i = 0; i < EX_DPT; i++) { acc[i] = 0.0; w[i] = 1.0f; }
#pragma unroll
    for (int i = 0; i < EX_DPT / 2; i++) cnt[i] = 0;
    if (active) {
#pragma unroll
        for (int i = 0; i < EX_DPT; i += 4) {
            float4 v = *reinterpret_cast<const float4*>(doc_w + d0 + i);
            w[i] = v.x; w[i + 1] = v.y; w[i + 2] = v.z; w[i + 3] = v.w;
        }
    }
    const u8* tbase = T.tab + (u64)d0 / 2;
    const i32 stage = min(EX_STAGE, (i32)blockDim.x);     // thread i owns the exception list of staged clause i
    for (i32 c0 = 0; c0 < nc; c0 += stage) {
        __syncthreads();
        const i32 lim = min(stage, nc - c0);
        for (int i = threadIdx.x; i < lim; i += blockDim.x) {
            s_val[i] = clause_val[base + c0 + i];
            s_row[i] = clause_row[base + c0 + i];
        }
        __syncthreads();
        if (active) {
#pragma unroll 2
            for (i32 ci = 0; ci < lim; ci++) {
                const uint2 q = ldg_stream8(tbase + (u64)row_data(T, s_row[ci]) * T.row_bytes);
                const float val = SIM == SIM_TFIDF ? s_val[ci] : __fmul_rn(s_val[ci], k1p1);
#pragma unroll
                for (int i = 0; i < EX_DPT; i++) {
                    const u32 code = ((i < 8 ? q.x : q.y) >> (4 * (i & 7))) & 15u;
                    const float f = s_tf[code];                 // tf-idf: tf(freq); BM25: freq; [0] = 0
                    float sc;
                    if (SIM == SIM_TFIDF) sc = __fmul_rn(__fmul_rn(f, val), w[i]);
                    else sc = code ? __fdiv_rn(__fmul_rn(val, f), __fadd_rn(f, w[i])) : 0.0f;
                    acc[i] += (double)sc;
                    cnt[i >> 1] += (code ? 1u : 0u) << (16 * (i & 1));
                }
            }
        }
        // exception cells of this stage, in windows of EX_EXC_CAP
        u32 my_a = 0, my_n = 0;
        if ((int)threadIdx.x < lim) {
            row_exceptions(T, s_row[threadIdx.x], my_a, my_n);
        }
        u32 total;
        const u32 my_base = block_exclusive_scan<u32>(my_n, &total, s_scan);
        for (u32 w0 = 0; w0 < total; w0 += EX_EXC_CAP) {
            __syncthreads();
            for (u32 j = 0; j < my_n; j++) {
                const u32 g = my_base + j;
                if (g >= w0 && g < w0 + EX_EXC_CAP) {
                    s_exc_c[g - w0] = threadIdx.x;
                    s_exc_d[g - w0] = T.exc_doc[my_a + j];
                    s_exc_f[g - w0] = T.exc_freq[my_a + j];
                }
            }
            __syncthreads();
            const u32 ne = min((u32)EX_EXC_CAP, total - w0);
            for (u32 e = 0; e < ne; e++) {
                const u32 d = s_exc_d[e];
                if (active && d >= d0 && d < d0 + EX_DPT) {
                    const float fr = s_exc_f[e];
                    const float val = SIM == SIM_TFIDF ? s_val[s_exc_c[e]] : __fmul_rn(s_val[s_exc_c[e]], k1p1);
                    const float f = SIM == SIM_TFIDF ? tf_of(fr) : fr;
#pragma unroll
                    for (int i = 0; i < EX_DPT; i++) {      // predicated, statically indexed: accumulators stay in registers
                        const bool hit = d - d0 == (u32)i;
                        float sc;
                        if (SIM == SIM_TFIDF) sc = __fmul_rn(__fmul_rn(f, val), w[i]);
                        else sc = __fdiv_rn(__fmul_rn(val, f), __fadd_rn(f, w[i]));
                        acc[i] += hit ? (double)sc : 0.0;
                        cnt[i >> 1] += hit ? (1u << (16 * (i & 1))) : 0u;
                    }
                }
            }
        }
    }
    // final scores + fused top-10
    const i32 need = max(1, msm[r]);
    float sc[EX_DPT];
#pragma unroll
    for (int i = 0; i < EX_DPT; i++) {
        i32 m = (i32)((cnt[i >> 1] >> (16 * (i & 1))) & 0xFFFFu);
        sc[i] = (active && m >= need) ? final_score(SIM, acc[i], m, nc) : -1.0f;      // padding columns have no cells: m = 0
    }
    const i64 pbase = (i64)r * parts_stride + tile;
    i32 found = 0;
    u64 mine = 0;
    bool dirty = true;
    for (int round = 0; round < TOPN; round++) {
        if (dirty) {
            mine = 0;
#pragma unroll
            for (int i = 0; i < EX_DPT; i++)
                if (sc[i] >= 0.0f) { u64 kk = cand_key(sc[i], T.docmap[d0 + i]); if (kk > mine) mine = kk; }
            dirty = false;
        }
        u64 best = block_max_u64(mine, red);
        if (best == 0) break;
        if (best == mine) {
            u32 d = 0xFFFFFFFFu - (u32)(best & 0xFFFFFFFFu);
            part_doc[pbase * TOPN + round] = global_doc_base + d;
            part_score[pbase * TOPN + round] = __uint_as_float((u32)(best >> 32));
#pragma unroll
            for (int i = 0; i < EX_DPT; i++) if (sc[i] >= 0.0f && T.docmap[d0 + i] == d) sc[i] = -1.0f;
            dirty = true;
        }
        found++;
    }
    if (threadIdx.x == 0) part_n[pbase] = found;
}
