// index_build.cuh -- GPU index construction kernels (K1-K4 of SURVEY.md 2.1).
//
//   K1 pack_kernel        ASCII FASTA sequence -> 2-bit packed words + invalid-base mask
//   K2 count/extract      sliding k-windows with the reference analyzer's exact
//                         N-handling, reverse-complement document and min-strand rule
//                         -> (term key, token ordinal) tuples in (doc, position) order
//   K3 radix sort + CSR   stable LSD radix sort by term, run-length -> term-major
//                         postings (doc, tf via position offsets) + positions
//   K4 term dictionary    direct table (4^k <= 2^24) or open-addressing hash
//
// Reference semantics: lucene/KmerSequenceTokenizer.java:115-154 (skips=0 on the index
// side, lucene/KmerIndexAnalyzer.java:48), lucene/SequenceCompressFilter.java:51-91,
// utils/SequenceHelper.java:26-35,85-91,225-271, index/Indexer.java:178-232, and
// lucene!index/DefaultIndexingChain$PerField#invert (positions = token ordinals).
#pragma once
#include "common.cuh"
#include "scan_sort.cuh"

#define CHUNK_THREADS 256
#define CHUNK_WPT 8                                // windows per thread
#define CHUNK_WINDOWS (CHUNK_THREADS * CHUNK_WPT)  // windows per chunk (= per CTA)

// ------------------------------------------------------------------------ K1
// One thread packs 32 bases: 2 packed words + 1 mask word.  flags[0] |= 1 when a char
// lies outside 'A'..'Z' (SequenceHelper.getComplement would throw -> reverse doc lost).
__global__ void __launch_bounds__(256) pack_kernel(const u8* __restrict__ seq, i64 len, u32* __restrict__ packed,
                                                    u32* __restrict__ mask, i64 n_groups, u32* __restrict__ flags) {
    i64 g = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    i64 b0 = g * 32;
    u32 w0 = 0, w1 = 0, m = 0;
    bool out_of_range = false;
    __align__(16) u8 c[32];
    if (b0 + 32 <= len && ((((uintptr_t)seq) + (uintptr_t)b0) & 15) == 0) {
        const uint4* v = reinterpret_cast<const uint4*>(seq + b0);
        uint4 a = __ldg(v), b = __ldg(v + 1);
        *reinterpret_cast<uint4*>(c) = a;
        *reinterpret_cast<uint4*>(c + 16) = b;
    } else {
#pragma unroll
        for (int i = 0; i < 32; i++) c[i] = (b0 + i < len) ? __ldg(seq + b0 + i) : (u8)'A';
    }
#pragma unroll
    for (int i = 0; i < 32; i++) {
        u8 ch = c[i];
        u32 code = 0;
        bool ok = true;
        switch (ch) {
            case 'A': code = 0; break;
            case 'C': code = 1; break;
            case 'G': code = 2; break;
            case 'T': code = 3; break;
            default: ok = false; break;
        }
        if (b0 + i < len) {
            if (!ok) m |= (1u << i);
            if (ch < 'A' || ch > 'Z') out_of_range = true;
        }
        if (i < 16) w0 |= code << (30 - 2 * i);
        else w1 |= code << (30 - 2 * (i - 16));
    }
    packed[2 * g] = w0;
    packed[2 * g + 1] = w1;
    mask[g] = m;
    if (out_of_range) atomicOr(flags, 1u);
}

// ------------------------------------------------------------------------ K2
// Device view of the staged reference entries.
struct EntryDev {
    u64 base_off;      // first base of the entry in the packed/mask arrays (multiple of 64)
    u32 len;           // bases
    u32 first_chunk;   // global index of the entry's first chunk
    u32 tok_fwd;       // segment-relative token start of the forward (or min_strand) doc
    u32 tok_rev;       // ... of the reverse doc, 0xFFFFFFFF if there is none
};

// number of valid windows per chunk
__global__ void __launch_bounds__(CHUNK_THREADS) chunk_count_kernel(const u32* __restrict__ mask,
                                                                     const EntryDev* __restrict__ entries,
                                                                     const u32* __restrict__ chunk_entry, int k,
                                                                     u32* __restrict__ chunk_cnt) {
    __shared__ u32 sm[33];
    const u32 c = blockIdx.x;
    const EntryDev e = entries[chunk_entry[c]];
    const i64 nwin = (i64)e.len - k + 1;
    const i64 w0 = (i64)(c - e.first_chunk) * CHUNK_WINDOWS + (i64)threadIdx.x * CHUNK_WPT;
    u32 cnt = 0;
#pragma unroll
    for (int i = 0; i < CHUNK_WPT; i++) {
        i64 w = w0 + i;
        if (w < nwin && !window_invalid(mask, e.base_off + (u64)w, k)) cnt++;
    }
    u32 tot;
    block_exclusive_scan<u32>(cnt, &tot, sm);
    if (threadIdx.x == 0) chunk_cnt[c] = tot;
}

// Emit (key, ordinal) tuples for chunks [c0, c0+gridDim.x).  chunk_base[c] is the global
// forward-token ordinal of the chunk's first valid window; ordinal within the entry
// r = chunk_base[c] - chunk_base[first_chunk] + (rank inside the chunk).
// Forward doc: token r at tok_fwd + r.  Reverse doc (= revcomp(sequence)): the window is
// the reverse complement and its ordinal is ntok-1-r (valid windows mirror exactly).
template <typename K>
__global__ void __launch_bounds__(CHUNK_THREADS) extract_kernel(const u32* __restrict__ packed, const u32* __restrict__ mask,
                                                                 const EntryDev* __restrict__ entries,
                                                                 const u32* __restrict__ chunk_entry,
                                                                 const u64* __restrict__ chunk_base, u32 c0, int k,
                                                                 int min_strand, K* __restrict__ keys, u32* __restrict__ vals) {
    __shared__ u32 sm[33];
    const u32 c = c0 + blockIdx.x;
    const u32 ei = chunk_entry[c];
    const EntryDev e = entries[ei];
    const EntryDev en = entries[ei + 1];       // sentinel entry follows the last real one
    const i64 nwin = (i64)e.len - k + 1;
    const i64 w0 = (i64)(c - e.first_chunk) * CHUNK_WINDOWS + (i64)threadIdx.x * CHUNK_WPT;
    u32 valid = 0, cnt = 0;
#pragma unroll
    for (int i = 0; i < CHUNK_WPT; i++) {
        i64 w = w0 + i;
        if (w < nwin && !window_invalid(mask, e.base_off + (u64)w, k)) { valid |= 1u << i; cnt++; }
    }
    u32 ex = block_exclusive_scan<u32>(cnt, (u32*)nullptr, sm);
    const u64 ebase = chunk_base[e.first_chunk];
    const u32 ntok = (u32)(chunk_base[en.first_chunk] - ebase);
    u32 r = (u32)(chunk_base[c] - ebase) + ex;
#pragma unroll
    for (int i = 0; i < CHUNK_WPT; i++) {
        if (valid & (1u << i)) {
            u64 km = load_kmer(packed, e.base_off + (u64)(w0 + i), k);
            u32 o = e.tok_fwd + r;
            keys[o] = (K)canon_kmer(km, k, min_strand);
            vals[o] = o;
            if (e.tok_rev != 0xFFFFFFFFu) {
                u32 o2 = e.tok_rev + (ntok - 1u - r);
                keys[o2] = (K)rc_packed(km, k);
                vals[o2] = o2;
            }
            r++;
        }
    }
}

// ------------------------------------------------------------------------ K3
// After the stable sort by key: vals[i] = segment-relative token ordinal.
// doc_tok[d] = first ordinal of local doc d (n_docs+1 entries).
__device__ __forceinline__ u32 find_doc(const u32* __restrict__ doc_tok, u32 n_docs, u32 ord) {
    u32 lo = 0, hi = n_docs;                  // invariant: doc_tok[lo] <= ord < doc_tok[hi]
    while (hi - lo > 1) {
        u32 mid = (lo + hi) >> 1;
        if (doc_tok[mid] <= ord) lo = mid; else hi = mid;
    }
    return lo;
}

// ---- sorted (term, segment ordinal) tuples -> CSR, two passes over the tuples and no per-token scratch.
// The sorted ordinals ARE the positions array: positions[i] = token ordinal within the SEGMENT (= doc_tok[doc] +
// Lucene's position).  Every consumer of positions is a sloppy sweep over two lists of the same doc, which only looks at
// differences, so the per-doc offset is never subtracted on the device (bsp_term_postings does it for the tests);
// keeping the segment ordinal makes a term's position slice ascending across docs, which is what lets the warp scorer
// look for close (t0, t1) pairs with one balanced scan per clause.
// Tuple i starts a posting when its (term, doc) differs from tuple i-1's, and a term when its term does.
#define CSR_ROUNDS 8
#define CSR_TILE (256 * CSR_ROUNDS)
template <typename K>
__device__ __forceinline__ void csr_heads(const K* __restrict__ keys, const u32* __restrict__ vals, u32 i,
                                          const u32* __restrict__ doc_tok, u32 n_docs, u32& doc, bool& head_post, bool& head_term) {
    doc = find_doc(doc_tok, n_docs, vals[i]);
    if (i == 0) { head_post = head_term = true; return; }
    head_term = keys[i] != keys[i - 1];
    head_post = head_term || find_doc(doc_tok, n_docs, vals[i - 1]) != doc;
}
// pass 1: postings | terms << 32 that start in every tile
template <typename K>
__global__ void __launch_bounds__(256) csr_count_kernel(const K* __restrict__ keys, const u32* __restrict__ vals, u32 n,
                                                         const u32* __restrict__ doc_tok, u32 n_docs, u64* __restrict__ tile_counts) {
    __shared__ u32 s_cnt[2];
    if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const u32 base = blockIdx.x * CSR_TILE;
    u32 np = 0, nt = 0;
#pragma unroll
    for (int j = 0; j < CSR_ROUNDS; j++) {
        const u32 i = base + j * 256 + threadIdx.x;
        if (i < n) {
            u32 d; bool hp, ht;
            csr_heads<K>(keys, vals, i, doc_tok, n_docs, d, hp, ht);
            np += hp ? 1u : 0u; nt += ht ? 1u : 0u;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { np += __shfl_xor_sync(0xFFFFFFFFu, np, o); nt += __shfl_xor_sync(0xFFFFFFFFu, nt, o); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&s_cnt[0], np); atomicAdd(&s_cnt[1], nt); }
    __syncthreads();
    if (threadIdx.x == 0) tile_counts[blockIdx.x] = (u64)s_cnt[0] | ((u64)s_cnt[1] << 32);
}
// pass 2: with the exclusive scan of the tile counts, every head writes its posting (doc, first position) / term
template <typename K>
__global__ void __launch_bounds__(256) csr_write_kernel(const K* __restrict__ keys, const u32* __restrict__ vals, u32 n,
                                                         const u32* __restrict__ doc_tok, u32 n_docs, const u64* __restrict__ tile_offs,
                                                         u64* __restrict__ term_key, u32* __restrict__ term_off,
                                                         u32* __restrict__ post_doc, u32* __restrict__ post_pos) {
    __shared__ u32 sm[33];
    const u64 off = tile_offs[blockIdx.x];
    u32 pbase = (u32)(off & 0xFFFFFFFFu), tbase = (u32)(off >> 32);
    const u32 base = blockIdx.x * CSR_TILE;
#pragma unroll 1
    for (int j = 0; j < CSR_ROUNDS; j++) {
        const u32 i = base + j * 256 + threadIdx.x;
        u32 d = 0; bool hp = false, ht = false;
        if (i < n) csr_heads<K>(keys, vals, i, doc_tok, n_docs, d, hp, ht);
        const u32 v = (hp ? 1u : 0u) | (ht ? 0x10000u : 0u);                  // at most 256 of each per round
        u32 tot;
        const u32 ex = block_exclusive_scan<u32>(v, &tot, sm);
        if (hp) {
            const u32 pi = pbase + (ex & 0xFFFFu);
            post_doc[pi] = d;
            post_pos[pi] = i;
            if (ht) {
                const u32 ti = tbase + (ex >> 16);
                term_key[ti] = (u64)keys[i];
                term_off[ti] = pi;
            }
        }
        pbase += tot & 0xFFFFu;
        tbase += tot >> 16;
    }
}

__global__ void set_u32_kernel(u32* p, u32 v) { *p = v; }

// ------------------------------------------------------------------------ K4
__global__ void __launch_bounds__(256) fill_u32_kernel(u32* __restrict__ p, u64 n, u32 v) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 stride = (u64)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = v;
}
__global__ void __launch_bounds__(256) fill_u64_kernel(u64* __restrict__ p, u64 n, u64 v) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 stride = (u64)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = v;
}

__global__ void __launch_bounds__(256) direct_table_kernel(const u64* __restrict__ term_key, u32 n_terms,
                                                            u32* __restrict__ direct) {
    u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n_terms) direct[term_key[t]] = t;
}

// open addressing, linear probing; capacity is a power of two >= 2*n_terms
__global__ void __launch_bounds__(256) hash_insert_kernel(const u64* __restrict__ term_key, u32 n_terms,
                                                           u64* __restrict__ hkeys, u32* __restrict__ hvals, u64 cap_mask) {
    u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_terms) return;
    u64 key = term_key[t];
    u64 slot = mix64(key) & cap_mask;
    while (true) {
        unsigned long long prev = atomicCAS((unsigned long long*)&hkeys[slot], (unsigned long long)HASH_EMPTY,
                                            (unsigned long long)key);
        if (prev == HASH_EMPTY || prev == key) { hvals[slot] = t; return; }
        slot = (slot + 1) & cap_mask;
    }
}

// Per-segment term dictionary + postings (device pointers; lifetime owned by SegmentStore)
struct SegDev {
    const u32* direct;     // [4^k] term key -> local term id, NO_TERM if absent (direct mode)
    const u64* hkeys;      // hash mode
    const u32* hvals;
    u64 hmask;
    const u64* term_key;   // [T]
    const u32* term_off;   // [T+1] -> postings
    const u32* post_doc;   // [P]   local doc ids, ascending within a term
    const u32* post_pos;   // [P+1] -> positions (tf = post_pos[i+1]-post_pos[i])
    const u32* positions;  // [n_tok] token ordinals within the doc
    u32 n_terms, n_post, n_tok, n_docs;
    u32 doc_base;          // global id of local doc 0 (within this handle)
};

__device__ __forceinline__ u32 seg_find_term(const SegDev& s, u64 key) {
    if (s.direct) return s.direct[key];
    u64 slot = mix64(key) & s.hmask;
    while (true) {
        u64 kk = s.hkeys[slot];
        if (kk == key) return s.hvals[slot];
        if (kk == HASH_EMPTY) return NO_TERM;
        slot = (slot + 1) & s.hmask;
    }
}

