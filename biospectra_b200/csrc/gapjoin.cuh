// gapjoin.cuh -- the gap clauses of a whole sub-batch against the positional index, index-stationary.
//
// A gap clause is a sloppy phrase (t0@0, t1@1, slop) whose two k-mers are not adjacent in the read: every phrase clause
// when kmer_skips > 0 (Configuration.java:37-44 defaults to 5; Classifier.java:160-200 builds slop = offset difference + 1),
// and the clauses next to an N otherwise.  No table row exists for it; its matches are the documents in which an
// occurrence of t1 lies within the slop of an occurrence of t0 (SloppyPhraseScorer#phraseFreq: every matchLength it
// tests is |(b - 1) - a| of an actual pair a in P(t0), b in P(t1)) -- about one document per clause at C2.
//
// score_positional_warp_kernel evaluates such a read segment by segment and re-stages both terms' slices for every
// (read, segment).  Here the batch's clauses are sorted by first k-mer and cut into chunks of at most GJ_C clauses that
// share it; a warp owns a chunk and walks the segments once:
//   * the first term's positions of the segment (contiguous in the term-major layout; (start, length) comes from a dense
//     per-segment directory `pse`, looked up two segments ahead) arrive by cp.async one segment ahead, double-buffered;
//   * a bit table over them (bucket = position >> bs with 2^bs >= the window 2 slop + 1, folded onto 32,768 bits) answers
//     "may there be an a in [b - 1 - slop, b - 1 + slop]" with one shared load, a shift and a mask;
//   * every clause of the chunk then streams its second term's positions straight from global memory into registers
//     (16-byte non-allocating loads, the next round in flight while this one is probed) and probes the table;
//   * a flagged position is checked exactly (binary search in the staged block);
//   * the rare hit is verified by its own lane on the global slices: document of the occurrence, both terms' positions
//     inside it, the literal sweep (sloppy_freq_distinct / _same); the first close pair of a document reports it.
// Matches are appended as (clause, table column, frequency); the host sorts them by (clause, column) and turns them
// into the per-clause exception lists the filter and exact scorers already consume (Tab4Dev::gap_off).
// Bytes per clause and segment: the second term's slice, plus the first term's divided by the clauses sharing it.
#pragma once
#include "dense4.cuh"

#define GJ_C 8                   // clauses per chunk (they share the first k-mer)
#ifndef GJ_ACAP
#define GJ_ACAP 384              // positions of the first term staged at a time (longer slices: block by block)
#endif
#ifndef GJ_WARPS
#define GJ_WARPS 10               // 2 CTAs of 10 warps per SM: 111 KB of shared memory each (measured: 8 x 2 aligned 283 ms, 10 x 2 248 ms, 8 x 3 with 32 Kbit tables 256 ms)
#endif
#ifndef GJ_MINB
#define GJ_MINB 2
#endif
#ifndef GJ_VEC
#define GJ_VEC 3                 // 16-byte vectors per lane and round of the second term (vector q = positions [128 q, 128 q + 128))
#endif
#define GJ_ROUND (GJ_VEC * 128)  // 384 positions per round
#ifndef GJ_BM_WORDS
#define GJ_BM_WORDS 2048         // words of a warp's bucket bit table (power of two, multiple of 128): 65,536 bits
#endif
#define GJ_BM_BYTES (GJ_BM_WORDS * 4)
// dynamic shared memory of a CTA: the bit tables (+ one table of slack: they are aligned to their size) and the staging buffers
#ifndef GJ_ALIGNED
#define GJ_ALIGNED 0             // tables aligned to their size (address = base | offset) at the cost of one table of slack per CTA
#endif
#define GJ_SMEM_BYTES (GJ_WARPS * GJ_BM_BYTES + GJ_ALIGNED * GJ_BM_BYTES + GJ_WARPS * 2 * (GJ_ACAP + 8) * 4)
#define GJ_LUT_SHIFT 16          // doc_lut[ordinal >> 16] = document of the bucket's first ordinal

struct __align__(16) GjSeg { const u32* positions; const u32* doc_tok; const u16* doc_lut; u32 n_docs, doc_base; };
struct GjOut {
    u32* g; u32* col; float* freq;         // appended matches
    u32 cap;
    u32* n_match;                          // appended (may exceed cap: overflow)
    unsigned long long* pos_bytes;         // position bytes the launch fetched (roofline numerator)
};

// (start, length) of every term's positions in one segment
__global__ void __launch_bounds__(256) pse_build_kernel(SegDev s, uint2* __restrict__ pse_seg) {
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= s.n_terms) return;
    const u32 a = s.term_off[t], b = s.term_off[t + 1];
    const u32 p = s.post_pos[a];
    pse_seg[s.term_key[t]] = make_uint2(p, s.post_pos[b] - p);
}

// doc_lut[x] = the document that holds ordinal x << GJ_LUT_SHIFT (segments hold at most 4,096 documents)
__global__ void __launch_bounds__(256) gj_doc_lut_kernel(const u32* __restrict__ doc_tok, u32 n_docs, u32 n_buckets, u16* __restrict__ lut) {
    const u32 x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n_buckets) return;
    const u32 o = x << GJ_LUT_SHIFT;
    u32 lo = 0, hi = n_docs;                                 // last doc whose first ordinal is <= o
    while (hi - lo > 1) {
        const u32 mid = (lo + hi) >> 1;
        if (doc_tok[mid] <= o) lo = mid; else hi = mid;
    }
    lut[x] = (u16)lo;
}

__global__ void __launch_bounds__(256) gj_keys_kernel(const GapClause* __restrict__ gaps, u32 n, u32* __restrict__ keys, u32* __restrict__ vals) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    keys[i] = (u32)gaps[i].k0;
    vals[i] = i;
}

// runs of equal first k-mers in the sorted key array -> chunks (first sorted index, clause count <= GJ_C), grouped by
// clause count, largest first: the warps of a CTA then carry similar loads (a CTA's shared memory is held until its
// slowest warp is done) and the long chunks start first.  sizes[0..8]: chunks per clause count, sizes[9..17]: first
// chunk of every count, sizes[18..26]: cursors.  mode 0 counts, gj_chunk_offsets_kernel lays the groups out, mode 1 writes.
__global__ void __launch_bounds__(256) gj_chunk_kernel(const u32* __restrict__ keys, u32 n, uint2* __restrict__ chunks, u32* __restrict__ sizes,
                                                        int mode) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u32 key = keys[i];
    if (i > 0 && keys[i - 1] == key) return;
    u32 lo = i + 1, hi = n;
    while (lo < hi) {
        const u32 mid = (lo + hi) >> 1;
        if (keys[mid] == key) lo = mid + 1; else hi = mid;
    }
    const u32 len = lo - i, full = len / GJ_C, rest = len % GJ_C;
    if (mode == 0) {
        if (full) atomicAdd(&sizes[GJ_C], full);
        if (rest) atomicAdd(&sizes[rest], 1u);
        return;
    }
    if (full) {
        const u32 base = sizes[9 + GJ_C] + atomicAdd(&sizes[18 + GJ_C], full);
        for (u32 j = 0; j < full; j++) chunks[base + j] = make_uint2(i + j * GJ_C, (u32)GJ_C);
    }
    if (rest) chunks[sizes[9 + rest] + atomicAdd(&sizes[18 + rest], 1u)] = make_uint2(i + full * GJ_C, rest);
}
__global__ void gj_chunk_offsets_kernel(u32* __restrict__ sizes, u32* __restrict__ n_chunks) {
    u32 at = 0;
    for (int c = GJ_C; c >= 1; c--) { sizes[9 + c] = at; at += sizes[c]; }
    *n_chunks = at;
}

__device__ __forceinline__ uint4 ldg_stream16(const u32* p) {
    uint4 v;
    asm("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ u32 lds32(u32 sa) {
    u32 v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(sa) : "memory");
    return v;
}
__device__ __forceinline__ void sts128(u32 sa, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(sa), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void cp_async_wait1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait0() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ u32 gj_lower_bound(const u32* __restrict__ P, u32 n, u32 v) {
    u32 lo = 0, hi = n;
    while (lo < hi) {
        const u32 mid = (lo + hi) >> 1;
        if (P[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// Candidates: (clause, segment, index of an occurrence b of the second k-mer) for which the join found an occurrence of the
// first k-mer inside the window [b - 1 - slop, b - 1 + slop] -- each (clause, segment, b) at most once.  The join only
// appends them; gj_verify_kernel (thread per candidate) does everything that needs the document: which document b lies in
// (bucket LUT), both k-mers' positions inside it (binary search / short scans on the global slices), whether b is the
// first of the document's occurrences with a partner INSIDE the document (that one reports the document; a partner in the
// neighbouring document is no match), the literal sweep (sloppy_freq_distinct / _same), and the append of
// (clause, table column, frequency).
struct __align__(16) GjCand { u32 g, seg, jb, ia; };     // ia: index of the partner the join found in the first k-mer's slice

__global__ void __launch_bounds__(256) gj_verify_kernel(const GjCand* __restrict__ cand, const u32* __restrict__ n_cand, u32 cand_cap,
                                                         const GjSeg* __restrict__ segs, const uint2* __restrict__ pse, u64 n_terms,
                                                         const GapClause* __restrict__ gaps, const u32* __restrict__ colmap, const GjOut out) {
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= min(*n_cand, cand_cap)) return;
    const GjCand c = cand[t];
    const GapClause gc = gaps[c.g];
    const GjSeg sg = segs[c.seg];
    const uint2 dA = __ldg(pse + (u64)c.seg * n_terms + gc.k0), dB = __ldg(pse + (u64)c.seg * n_terms + gc.k1);
    const bool same = gc.k0 == gc.k1;
    const int slop = gc.slop;
    const u32* A = sg.positions + dA.x;
    const u32* B = sg.positions + dB.x;
    const u32 mA = dA.y, mB = dB.y, jb = c.jb;
    BSP_ASSERT(jb < mB && c.ia <= mA && mA > 0u);
    const u32 b = B[jb];
    u32 d = sg.doc_lut[b >> GJ_LUT_SHIFT];
    BSP_ASSERT(d < sg.n_docs && sg.doc_tok[d] <= b);
    while (sg.doc_tok[d + 1] <= b) d++;
    const u32 ds = sg.doc_tok[d], de = sg.doc_tok[d + 1];
    BSP_ASSERT(d < sg.n_docs && ds <= b && b < de);
    // the first k-mer's positions inside the document: from the lower bound of the window's low end, outwards
    u32 ia0 = min(c.ia, mA);                                     // the partner the join found; outwards from there
    if (ia0 < mA && A[ia0] >= ds) { while (ia0 > 0 && A[ia0 - 1] >= ds) ia0--; }
    else { while (ia0 < mA && A[ia0] < ds) ia0++; }
    u32 ia1 = ia0;
    while (ia1 < mA && A[ia1] < de) ia1++;
    if (ia0 == ia1) return;
    u32 jb0 = jb, jb1 = jb + 1;
    while (jb0 > 0 && B[jb0 - 1] >= ds) jb0--;
    while (jb1 < mB && B[jb1] < de) jb1++;
    bool mine = false;
    for (u32 j = jb0; j <= jb; j++) {
        const u32 bb = B[j];
        bool partner = false;
        for (u32 x = ia0; x < ia1; x++) {
            const u32 a = A[x];
            if (same && a == bb) continue;
            const int dist = (int)bb - 1 - (int)a;
            if (dist >= -slop && dist <= slop) { partner = true; break; }
        }
        if (partner) { mine = j == jb; break; }
    }
    if (!mine) return;
    const float f = same ? sloppy_freq_same(A + ia0, (int)(ia1 - ia0), slop)
                         : sloppy_freq_distinct(A + ia0, (int)(ia1 - ia0), B + jb0, (int)(jb1 - jb0), slop);
    if (f != 0.0f) {
        const u32 at = atomicAdd(out.n_match, 1u);
        if (at < out.cap) { out.g[at] = c.g; out.col[at] = colmap[sg.doc_base + d]; out.freq[at] = f; }
    }
}

__device__ __forceinline__ u32 gj_sel(const uint4 (&v)[GJ_VEC], int bit) {      // element `bit` of the round's registers
    const uint4 x = bit < 4 ? v[0] : (bit < 8 ? v[1] : v[GJ_VEC - 1]);
    const int j = bit & 3;
    return j == 0 ? x.x : (j == 1 ? x.y : (j == 2 ? x.z : x.w));
}
__device__ __forceinline__ u32 gj_hash2(u32 word_index) { return ((word_index * 0x9E3779B1u) >> 15) & (u32)(GJ_BM_WORDS - 1); }

// warp per chunk
__global__ void __launch_bounds__(GJ_WARPS * 32, GJ_MINB) gap_join_kernel(
        const GjSeg* __restrict__ segs, int n_segs, const uint2* __restrict__ pse, u64 n_terms,
        const GapClause* __restrict__ gaps, const u32* __restrict__ order, const uint2* __restrict__ chunks,
        const u32* __restrict__ n_chunks, GjCand* __restrict__ cand, u32* __restrict__ n_cand, u32 cand_cap,
        unsigned long long* __restrict__ pos_bytes, i32* __restrict__ route, i32* __restrict__ status, int positional_ok) {
    extern __shared__ __align__(16) u8 gj_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 w = blockIdx.x * GJ_WARPS + warp;
    if (w >= *n_chunks) return;
    // shared memory: the warps' bit tables first, each aligned to its own size (a table address is base | offset), then
    // their staging buffers
    const u32 raw_sa = (u32)__cvta_generic_to_shared(gj_raw);
#if GJ_ALIGNED
    const u32 bm0_sa = (raw_sa + (u32)(GJ_BM_BYTES - 1)) & ~(u32)(GJ_BM_BYTES - 1);
#else
    const u32 bm0_sa = raw_sa;
#endif
    const u32 bm_sa = bm0_sa + (u32)warp * GJ_BM_BYTES;
    u32* const bm = reinterpret_cast<u32*>(gj_raw + (bm_sa - raw_sa));
    u32* const abuf = reinterpret_cast<u32*>(gj_raw + (bm0_sa - raw_sa) + GJ_WARPS * GJ_BM_BYTES) + (size_t)warp * 2 * (GJ_ACAP + 8);
    const uint2 ch = chunks[w];
    const int cnt = (int)ch.y;
    // lanes [0, cnt): the chunk's clauses; lane GJ_C: the shared first k-mer
    u32 my_term = 0, my_g = 0;
    int my_slop = 0;
    u32 t0 = 0;
    if (lane < cnt) {
        my_g = order[ch.x + lane];
        const GapClause gc = gaps[my_g];
        my_term = (u32)gc.k1; my_slop = gc.slop; t0 = (u32)gc.k0;
    }
    t0 = __shfl_sync(0xFFFFFFFFu, t0, 0);
    if (lane == GJ_C) my_term = t0;
    const bool has_dir = lane < cnt || lane == GJ_C;
    // bucket width of the bit table: a clause's window [b - 1 - slop, b - 1 + slop] spans at most two buckets.  Positions are
    // biased by one bucket so that the window's low end never goes below zero.
    int max_slop = my_slop;
#pragma unroll
    for (int dd = 16; dd > 0; dd >>= 1) max_slop = max(max_slop, __shfl_xor_sync(0xFFFFFFFFu, max_slop, dd));
    const int bs = 32 - __clz(2 * max_slop);                 // 2^bs >= 2 slop + 1
    const u32 bias = 1u << bs;
    if (lane < cnt && my_term == t0) my_slop = -my_slop;     // same-term clause: marked by the sign
    u32 fetched = 0;                                         // positions fetched (warp-uniform)

    auto stage_first = [&](int sidx, uint2 d, u32* buf) {       // first block of the first term's slice of segment sidx
        const u32 pA = __shfl_sync(0xFFFFFFFFu, d.x, GJ_C), mA = __shfl_sync(0xFFFFFFFFu, d.y, GJ_C);
        if (mA) {
            const u32 off = pA & 3u, nf = min(mA, (u32)GJ_ACAP);
            const u32* src = segs[sidx].positions + (pA - off);
            for (u32 i = 4u * (u32)lane; i < nf + off; i += 128) cp_async16(buf + i, src + i);
        }
    };
    // one round of a second-term slice: GJ_VEC vectors of 128 consecutive positions, lane l holding 4 of each
    auto load_round = [&](uint4 (&dst)[GJ_VEC], const u32* positions, u32 pB, u32 mB, u32 r0) {
        const u32 offB = pB & 3u;
        const u32* src = positions + (pB - offB) + r0 + 4u * (u32)lane;
#pragma unroll
        for (int q = 0; q < GJ_VEC; q++)
            dst[q] = r0 + (u32)(128 * q + 4 * lane) < mB + offB ? ldg_stream16(src + 128 * q) : make_uint4(0, 0, 0, 0);
    };
    auto clear_table = [&]() {
#pragma unroll
        for (int q = 0; q < GJ_BM_WORDS / 128; q++) sts128(bm_sa + (u32)(q * 32 + lane) * 16u, make_uint4(0, 0, 0, 0));
    };

    uint2 d_cur = make_uint2(0, 0), d_nxt = make_uint2(0, 0);
    if (has_dir) {
        d_cur = __ldg(pse + my_term);
        if (n_segs > 1) d_nxt = __ldg(pse + n_terms + my_term);
    }
    stage_first(0, d_cur, abuf);
    cp_async_commit();
    // the segment's live clauses (non-empty slices on both sides) and the first round of the first one: requested a whole
    // loop turn before it is needed
    u32 live = 0, pB = 0, mB = 0, r0 = 0;
    int c = -1;
    uint4 v[GJ_VEC], vn[GJ_VEC];
    const u32* positions = nullptr;
    auto open_segment = [&](int sidx, uint2 d) {
        live = __ballot_sync(0xFFFFFFFFu, lane < cnt && d.y != 0u);
        if (__shfl_sync(0xFFFFFFFFu, d.y, GJ_C) == 0u) live = 0u;
        positions = segs[sidx].positions;
        c = __ffs(live) - 1; r0 = 0;
        if (c >= 0) {
            pB = __shfl_sync(0xFFFFFFFFu, d.x, c); mB = __shfl_sync(0xFFFFFFFFu, d.y, c);
            load_round(v, positions, pB, mB, 0);
        }
    };
    open_segment(0, d_cur);
    for (int s = 0; s < n_segs; s++) {
        __syncwarp();                                        // the buffer the next copies land in is no longer read
        uint2 d_nn = make_uint2(0, 0);
        if (has_dir && s + 2 < n_segs) d_nn = __ldg(pse + (u64)(s + 2) * n_terms + my_term);
        u32* const buf = abuf + (s & 1) * (GJ_ACAP + 8);
        if (s + 1 < n_segs) stage_first(s + 1, d_nxt, abuf + ((s + 1) & 1) * (GJ_ACAP + 8));
        cp_async_commit();
        const u32 pA = __shfl_sync(0xFFFFFFFFu, d_cur.x, GJ_C), mA = __shfl_sync(0xFFFFFFFFu, d_cur.y, GJ_C);
        if (c >= 0) clear_table();                           // while the copies land
        cp_async_wait1();
        __syncwarp();
        if (live != 0u) {
            const u32 mine_b = (live >> lane) & 1u ? d_cur.y : 0u;
            fetched += mA + __reduce_add_sync(0xFFFFFFFFu, mine_b) * ((mA + GJ_ACAP - 1) / GJ_ACAP);
            for (u32 blk0 = 0; blk0 < mA; blk0 += GJ_ACAP) {
                const u32 nA = min((u32)GJ_ACAP, mA - blk0);
                const u32 offA = (pA + blk0) & 3u;
                if (blk0) {                                  // later blocks of a long slice: staged on the spot, every clause again
                    __syncwarp();
                    const u32* src = positions + (pA + blk0 - offA);
                    for (u32 i = 4u * (u32)lane; i < nA + offA; i += 128) cp_async16(buf + i, src + i);
                    cp_async_commit();
                    clear_table();
                    cp_async_wait0();
                    __syncwarp();
                    c = __ffs(live) - 1; r0 = 0;
                    pB = __shfl_sync(0xFFFFFFFFu, d_cur.x, c); mB = __shfl_sync(0xFFFFFFFFu, d_cur.y, c);
                    load_round(v, positions, pB, mB, 0);
                }
                u32* const A = buf + offA;
                const u32 prev_last = blk0 ? positions[pA + blk0 - 1u] : 0u;             // last position of the block before this one
                if (lane == 0) { A[nA] = 0xFFFFFFFFu; A[nA + 1] = 0xFFFFFFFFu; }
                // ---- bit table (a Bloom filter with two hashes of the word index): bucket k = (a + bias) >> bs sets bit k & 31 of
                // words h1(k >> 5) and h2(k >> 5); bit 31 of a word also stands for bucket 0 of the next one, so that the two
                // buckets of a window are always found in one word
                BSP_ASSERT(offA + nA + 1u < (u32)(GJ_ACAP + 8) && nA >= 1u);                  // block + sentinels inside the staging buffer
                for (u32 i = lane; i < nA; i += 32) {
                    BSP_ASSERT(i == 0u || A[i - 1] < A[i]);                                     // a slice ascends strictly
                    const u32 k = (A[i] + bias) >> bs;
                    const u32 wi = k >> 5, bit = 1u << (k & 31u);
                    atomicOr(&bm[wi & (GJ_BM_WORDS - 1)], bit);
                    atomicOr(&bm[gj_hash2(wi)], bit);
                    if ((k & 31u) == 0u) {
                        atomicOr(&bm[(wi - 1u) & (GJ_BM_WORDS - 1)], 0x80000000u);
                        atomicOr(&bm[gj_hash2(wi - 1u)], 0x80000000u);
                    }
                }
                __syncwarp();
                const int steps = 32 - __clz(nA);            // binary search steps: 2^steps > nA
                u32 todo = live;
                while (c >= 0) {
                    // next step of the (clause, round) sequence: its vectors are in flight while this one is probed
                    int c2 = c; u32 pB2 = pB, mB2 = mB, r2 = r0 + GJ_ROUND;
                    if (r2 >= mB + (pB & 3u)) {
                        todo &= todo - 1u;
                        c2 = __ffs(todo) - 1;
                        r2 = 0;
                        if (c2 >= 0) { pB2 = __shfl_sync(0xFFFFFFFFu, d_cur.x, c2); mB2 = __shfl_sync(0xFFFFFFFFu, d_cur.y, c2); }
                    }
                    if (c2 >= 0) load_round(vn, positions, pB2, mB2, r2);
                    // ---- probe this round: x = b - 1 - slop + bias is the biased low end of the window
                    const int sslop = __shfl_sync(0xFFFFFFFFu, my_slop, c);
                    const bool same = sslop < 0;
                    const u32 slop = (u32)(same ? -sslop : sslop);
                    const u32 xoff = bias - 1u - slop;
                    const u32 offB = pB & 3u, endB = offB + mB;
                    // (elements outside the slice -- zeros behind its end in the last vector, the previous term's last positions in
                    //  front of an unaligned start -- are probed like the others and dropped where a flag is looked at)
                    u32 hits = 0;                                        // two bits per element
#pragma unroll
                    for (int q = 0; q < GJ_VEC; q++) {
                        if (r0 + (u32)(128 * q) < endB) {                              // (warp-uniform)
                            const u32 bv[4] = {v[q].x, v[q].y, v[q].z, v[q].w};
#pragma unroll
                            for (int j = 0; j < 4; j++) {
                                const u32 x = bv[j] + xoff;
#if GJ_ALIGNED
                                const u32 word = lds32(bm_sa | ((x >> (bs + 3)) & (u32)((GJ_BM_WORDS - 1) << 2)));
#else
                                const u32 word = lds32(bm_sa + ((x >> (bs + 3)) & (u32)((GJ_BM_WORDS - 1) << 2)));
#endif
                                const u32 t = __funnelshift_r(word, 0u, x >> bs) & 3u;          // (shift count taken mod 32)
                                hits = t * (1u << (2 * (4 * q + j))) + hits;                      // (the multiply-add pipe is idle)
                            }
                        }
                    }
                    if (__any_sync(0xFFFFFFFFu, hits != 0u)) {
                        // flagged elements: the second hash, then the exact check on the staged block (binary search), then the document
                        const u32 g = __shfl_sync(0xFFFFFFFFu, my_g, c);
                        while (hits) {
                            const int bit = (__ffs(hits) - 1) >> 1;
                            hits &= ~(3u << (2 * bit));
                            const u32 e = r0 + (u32)(128 * (bit >> 2) + 4 * lane + (bit & 3));
                            if (e < offB || e >= endB) continue;                       // not an element of the slice
                            const u32 b = gj_sel(v, bit);
                            const u32 k0 = (b + xoff) >> bs;
                            if (((bm[gj_hash2(k0 >> 5)] >> (k0 & 31u)) & 3u) == 0u) continue;
                            const u32 lo = b > slop + 1u ? b - 1u - slop : 0u;
                            u32 i = 0;                                   // lower bound of lo in A
                            for (int st = steps - 1; st >= 0; st--) {
                                const u32 t = i + (1u << st);
                                if (t <= nA && A[t - 1] < lo) i = t;
                            }
                            // a slice staged block by block: the block that holds the FIRST position >= lo reports the candidate
                            const bool edge = blk0 != 0u && i == 0u && prev_last >= lo;
                            if (same && A[i] == b) i++;
                            BSP_ASSERT(i <= nA + 1u && e - offB < mB);
                            if (A[i] - lo > b - 1u + slop - lo) continue;        // the bits were other buckets'
                            if (edge) {
                                if (same) {          // (the skipped a == b may sit at a block edge: leave this read to the positional scorer)
                                    const i32 r = gaps[g].read;
                                    route[r] = positional_ok ? ROUTE_POSITIONAL : ROUTE_NONE;
                                    if (!positional_ok) status[r] = ST_ERROR;
                                }
                                continue;
                            }
                            const u32 at = atomicAdd(n_cand, 1u);
                            if (at < cand_cap) { GjCand cd; cd.g = g; cd.seg = (u32)s; cd.jb = e - offB; cd.ia = blk0 + i; cand[at] = cd; }
                        }
                        __syncwarp();
                    }
#pragma unroll
                    for (int q = 0; q < GJ_VEC; q++) v[q] = vn[q];
                    c = c2; pB = pB2; mB = mB2; r0 = r2;
                }
            }
        }
        d_cur = d_nxt; d_nxt = d_nn;
        if (s + 1 < n_segs) open_segment(s + 1, d_cur);
    }
    cp_async_wait0();
    if (lane == 0) atomicAdd(pos_bytes, 4ull * (unsigned long long)fetched);
}

// after the matches were sorted by (clause, column): a clause with more than GAP_LIST matches sends its read to the
// positional scorer (the lists the scorers read are capped there); `all` != 0: the match buffer overflowed, every read
// with a gap clause goes
__global__ void __launch_bounds__(256) gj_overflow_kernel(const GapClause* __restrict__ gaps, u32 n_g, const u32* __restrict__ gap_off,
                                                           int all, i32* __restrict__ route, i32* __restrict__ status, int positional_ok) {
    const u32 g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_g) return;
    if (all || gap_off[g + 1] - gap_off[g] > (u32)GAP_LIST) {
        const i32 r = gaps[g].read;
        route[r] = positional_ok ? ROUTE_POSITIONAL : ROUTE_NONE;
        if (!positional_ok) status[r] = ST_ERROR;
    }
}
