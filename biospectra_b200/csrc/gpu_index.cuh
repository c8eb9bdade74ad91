// gpu_index.cuh -- host-side orchestration of the GPU index build.
//
// Data layout in HBM (per handle / per GPU):
//   RefStore   2-bit packed reference (u32 words, 16 bases each) + invalid-base bitmask;
//              this is also what the on-disk index holds -- the inverted index is rebuilt
//              on the GPU at open time, which is faster than reading it from disk.
//   Segment    (optional, "positional") term-major postings of a contiguous doc range
//              (<= 4096 docs, <= seg_tokens tokens): term dictionary (direct table or
//              open-addressing hash), post_doc[], post_pos[], positions[].
//   Tab4       (optional, small k) 4-bit frequency table: 4^k term rows + 4^(k+1) pair rows of
//              d_pad/2 bytes, plus the sorted exception list (see dense4.cuh).
#pragma once
#include "dense4.cuh"
#include "gapjoin.cuh"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

struct EntryMeta {
    std::string filename, header, taxonomy;
    i64 len = 0;
    i64 base_off = 0;       // in bases, multiple of 64
    int rev_ok = 1;         // all chars within 'A'..'Z' (reverse doc exists)
    i64 ntok = 0;           // valid windows (tokens of each strand doc)
    i32 doc_fwd = -1, doc_rev = -1;
};

struct DocMeta { i32 entry; i32 strand; };   // strand 0 forward, 1 reverse, 2 min_strand

static inline float h_byte315_to_float(int b) {       // lucene!util/SmallFloat#byte315ToFloat
    if (b == 0) return 0.0f;
    u32 bits = ((u32)(b & 0xFF) << 21) + ((63u - 15u) << 24);
    float f; memcpy(&f, &bits, 4); return f;
}
static inline int h_float_to_byte315(float f) {       // lucene!util/SmallFloat#floatToByte315
    i32 bits; memcpy(&bits, &f, 4);
    i32 s = bits >> 21;
    if (s <= ((63 - 15) << 3)) return bits <= 0 ? 0 : 1;
    if (s >= ((63 - 15) << 3) + 0x100) return 0xFF;
    return (s - ((63 - 15) << 3)) & 0xFF;
}
static inline int h_norm_byte(i64 ntok) {             // DefaultSimilarity#lengthNorm + #encodeNormValue
    volatile float ln = 1.0f * (float)(1.0 / std::sqrt((double)ntok));
    return h_float_to_byte315(ln);
}

// ------------------------------------------------------------------ RefStore
struct RefStore {
    int k = 10, min_strand = 0;
    std::vector<EntryMeta> entries;
    DevBuf<u32> packed, mask;      // capacity in words
    i64 used_bases = 0;            // multiple of 64
    DevBuf<u32> flag;              // 1 word: out-of-range char seen
    DevBuf<u8> stage;              // H2D staging for host sequences
    u8* pinned = nullptr; size_t pinned_bytes = 0;

    ~RefStore() { if (pinned) cudaFreeHost(pinned); }

    static i64 padded_len(i64 len) { return round_up(len, 64) + 128; }

    void reserve_bases(i64 total, cudaStream_t st) {
        size_t need_p = (size_t)(total / 16) + 8, need_m = (size_t)(total / 32) + 8;
        if (need_p > packed.n) {
            size_t np = std::max(need_p, packed.n + packed.n / 2);
            size_t nm = std::max(need_m, mask.n + mask.n / 2);
            DevBuf<u32> p2, m2;
            p2.alloc(np); m2.alloc(nm);
            CUDA_CHECK(cudaMemsetAsync(p2.p, 0, np * 4, st));
            CUDA_CHECK(cudaMemsetAsync(m2.p, 0, nm * 4, st));
            if (used_bases) {
                CUDA_CHECK(cudaMemcpyAsync(p2.p, packed.p, (size_t)(used_bases / 16) * 4, cudaMemcpyDeviceToDevice, st));
                CUDA_CHECK(cudaMemcpyAsync(m2.p, mask.p, (size_t)(used_bases / 32) * 4, cudaMemcpyDeviceToDevice, st));
            }
            CUDA_CHECK(cudaStreamSynchronize(st));
            packed = std::move(p2); mask = std::move(m2);
        }
    }

    // d_seq: device pointer to len ASCII bytes.  Returns rev_ok.
    int append_device(const u8* d_seq, i64 len, cudaStream_t st) {
        i64 pl = padded_len(len);
        reserve_bases(used_bases + pl, st);
        if (!flag.p) flag.alloc(1);
        CUDA_CHECK(cudaMemsetAsync(flag.p, 0, 4, st));
        i64 groups = pl / 32;
        pack_kernel<<<(unsigned)ceil_div(groups, 256), 256, 0, st>>>(d_seq, len, packed.p + used_bases / 16,
                                                                     mask.p + used_bases / 32, groups, flag.p);
        KERNEL_CHECK();
        u32 f = 0;
        CUDA_CHECK(cudaMemcpyAsync(&f, flag.p, 4, cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaStreamSynchronize(st));
        return f ? 0 : 1;
    }

    int append_host(const char* seq, i64 len, cudaStream_t st) {
        const size_t CH = (size_t)64 << 20;
        if (!pinned) { CUDA_CHECK(cudaMallocHost((void**)&pinned, CH)); pinned_bytes = CH; }
        stage.ensure((size_t)len + 64);
        for (i64 o = 0; o < len; o += (i64)CH) {
            size_t n = (size_t)std::min<i64>((i64)CH, len - o);
            memcpy(pinned, seq + o, n);
            CUDA_CHECK(cudaMemcpyAsync(stage.p + o, pinned, n, cudaMemcpyHostToDevice, st));
            CUDA_CHECK(cudaStreamSynchronize(st));
        }
        return append_device(stage.p, len, st);
    }

    void add(EntryMeta&& m, int rev_ok) {
        m.base_off = used_bases;
        m.rev_ok = rev_ok;
        used_bases += padded_len(m.len);
        entries.push_back(std::move(m));
    }
};

// ------------------------------------------------------------------ Segment
struct Segment {
    i32 first_entry = 0, n_entries = 0;
    u32 doc_base = 0, n_docs = 0, n_tok = 0, n_terms = 0, n_post = 0;
    DevBuf<u32> direct; DevBuf<u64> hkeys; DevBuf<u32> hvals; u64 hmask = 0;
    DevBuf<u64> term_key; DevBuf<u32> term_off, post_doc, post_pos, positions, doc_tok;
    SegDev view() const {
        SegDev v;
        v.direct = direct.p; v.hkeys = hkeys.p; v.hvals = hvals.p; v.hmask = hmask;
        v.term_key = term_key.p; v.term_off = term_off.p; v.post_doc = post_doc.p; v.post_pos = post_pos.p;
        v.positions = positions.p; v.n_terms = n_terms; v.n_post = n_post; v.n_tok = n_tok; v.n_docs = n_docs;
        v.doc_base = doc_base;
        return v;
    }
    size_t bytes() const {
        return direct.bytes() + hkeys.bytes() + hvals.bytes() + term_key.bytes() + term_off.bytes() +
               post_doc.bytes() + post_pos.bytes() + positions.bytes() + doc_tok.bytes();
    }
};

#define SEG_MAX_DOCS 4096

struct BuildOptions {
    int want_tables = 0;          // 0 auto 1 on 2 off
    int want_positional = 0;      // 0 auto 1 on 2 off
    int need_pair_table = 1;      // query algorithm uses phrases
    i64 seg_tokens = (i64)1 << 28;
    i64 exc_cap = (i64)1 << 26;   // capacity of the exception append buffer (entries)
    i64 gap_cap = GAP_CAP_DEFAULT; // gap clauses per sub-batch
};

// Slab allocator of the positional index: the segments' arrays are carved out of a few large allocations (first fit over
// the slabs' tails) instead of two cudaMallocs per segment.
#define BSP_SLOTS 3              // sub-batches in flight in bsp_classify_packed (api.cu)

struct DevArena {
    struct Slab { DevBuf<u8> buf; size_t used = 0; };
    std::vector<Slab> slabs;
    size_t slab_bytes = (size_t)8 << 30;
    u8* take(size_t bytes) {
        bytes = (bytes + 255) & ~(size_t)255;
        for (Slab& sl : slabs)
            if (sl.buf.n - sl.used >= bytes) { u8* p = sl.buf.p + sl.used; sl.used += bytes; return p; }
        slabs.emplace_back();
        Slab& sl = slabs.back();
        try { sl.buf.alloc(std::max(slab_bytes, bytes)); }
        catch (const BspError&) {                      // not enough room for a whole slab: exactly what is asked for
            cudaGetLastError();
            try { sl.buf.alloc(bytes); } catch (...) { slabs.pop_back(); throw; }
        }
        sl.used = bytes;
        return sl.buf.p;
    }
    void release_all() { slabs.clear(); }
    void reset() { for (Slab& sl : slabs) sl.used = 0; }          // reuse the slabs (segment arrays not kept)
};

struct GpuIndex {
    int k = 10, min_strand = 0;
    bool direct = true;
    RefStore* ref = nullptr;
    std::vector<DocMeta> docs;
    std::vector<u32> doc_ntok; std::vector<u8> doc_norm;
    std::vector<Segment> segs;
    DevArena arena;                    // backing store of the segments' arrays
    DevBuf<u32> doc_tok_all;           // every segment's doc_tok array
    DevBuf<SegDev> d_segs;
    bool positional = false, tables = false, has_pair = false;
    i64 n_terms = 0, n_postings = 0, n_positions = 0, sum_ttf = 0;
    // direct mode term statistics (this handle's docs)
    DevBuf<u32> df_dense; DevBuf<u32> ttf_dense;
    // dense 4-bit table
    DevBuf<u8> tab4; u64 d_pad = 0, row_bytes = 0, n_rows = 0, pair_row0 = 0;
    DevBuf<u32> exc_off, exc_doc; DevBuf<float> exc_freq; i64 n_exc = 0;
    // the tail of the exception arrays holds the gap-clause match lists of the pipeline slots: gap_tail entries each
    u32 gap_cap = GAP_CAP_DEFAULT; size_t gap_tail = 0;
    // gap join (gapjoin.cuh): per segment the (start, length) of every term's positions, built when the first gap clause shows up
    DevBuf<uint2> pse; DevBuf<GjSeg> d_gjsegs; DevBuf<u16> gj_lut; bool gap_join_off = false;
    // table column order: docs sorted by (norm byte, doc id) -- see Tab4Dev::docmap
    DevBuf<u32> colmap, docmap; std::vector<u32> h_docmap;
    // slot: pipeline slot whose gap-clause lists (tail of the exception arrays) the view refers to
    Tab4Dev tab4_view(int slot, const u32* gap_off) const {
        Tab4Dev t;
        t.tab = tab4.p; t.row_bytes = row_bytes; t.pair_row0 = pair_row0;
        t.exc_off = exc_off.p; t.exc_doc = exc_doc.p; t.exc_freq = exc_freq.p;
        t.zero_row = (u32)n_rows; t.temp_row0 = (u32)n_rows + 1;
        t.tmp_base = (u32)n_exc + (u32)((size_t)slot * gap_tail);
        t.gap_off = gap_off;
        t.docmap = docmap.p;
        return t;
    }
    u32 max_seg_docs = 0;
    size_t positional_bytes = 0, table_bytes = 0;
    std::string notes;
};

__global__ void __launch_bounds__(256) accumulate_df_kernel(SegDev s, u32* __restrict__ df, u32* __restrict__ ttf) {
    u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= s.n_terms) return;
    u64 key = s.term_key[t];
    u32 a = s.term_off[t], b = s.term_off[t + 1];
    df[key] += b - a;                        // one segment at a time on one stream: no race
    ttf[key] += s.post_pos[b] - s.post_pos[a];
}

// terms of segment `s` that occur in none of segs[0..n_prev): distinct-term counting (hash mode)
__global__ void __launch_bounds__(256) count_new_terms_kernel(SegDev s, const SegDev* __restrict__ prev, int n_prev,
                                                               unsigned long long* __restrict__ counter) {
    u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    bool is_new = false;
    if (t < s.n_terms) {
        is_new = true;
        u64 key = s.term_key[t];
        for (int i = 0; i < n_prev; i++)
            if (seg_find_term(prev[i], key) != NO_TERM) { is_new = false; break; }
    }
    u32 bal = __ballot_sync(0xFFFFFFFFu, is_new);
    if ((threadIdx.x & 31) == 0 && bal) atomicAdd(counter, (unsigned long long)__popc(bal));
}

__global__ void __launch_bounds__(256) count_nonzero_kernel(const u32* __restrict__ a, u64 n, unsigned long long* __restrict__ counter) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    bool nz = i < n && a[i] != 0;
    u32 bal = __ballot_sync(0xFFFFFFFFu, nz);
    if ((threadIdx.x & 31) == 0 && bal) atomicAdd(counter, (unsigned long long)__popc(bal));
}

// BSP_BUILD_TIMING: wall time per build phase (adds a stream synchronize at every phase boundary)
struct BuildPhases {
    bool on = false;
    double ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    std::chrono::steady_clock::time_point t;
    void start(cudaStream_t st) { if (on) { cudaStreamSynchronize(st); t = std::chrono::steady_clock::now(); } }
    void tick(int i, cudaStream_t st) {
        if (!on) return;
        cudaStreamSynchronize(st);
        const auto n = std::chrono::steady_clock::now();
        ms[i] += std::chrono::duration<double, std::milli>(n - t).count();
        t = n;
    }
};
static thread_local BuildPhases g_phases;

template <typename K>
static void build_segment(GpuIndex& ix, Segment& seg, const EntryDev* d_entries, const u32* d_chunk_entry,
                          const u64* d_chunk_base, u32 c0, u32 c1, DevBuf<u8>& kbuf_a, DevBuf<u8>& kbuf_b,
                          DevBuf<u32>& vbuf_a, DevBuf<u64>& flags, RadixSortTemp& rst,
                          DevBuf<u8>& scan_tmp, DevBuf<u64>& d_total, cudaStream_t st) {
    const u32 n = seg.n_tok;
    RefStore& ref = *ix.ref;
    kbuf_a.ensure((size_t)n * sizeof(K)); kbuf_b.ensure((size_t)n * sizeof(K));
    vbuf_a.ensure(n);
    K* keys = (K*)kbuf_a.p; K* keys_alt = (K*)kbuf_b.p;
    // The sorted ordinals are the segment's positions array: the sort ping-pongs between a scratch buffer and the array
    // itself, arranged so that the last pass lands in the array.
    seg.positions.adopt((u32*)ix.arena.take((size_t)std::max<u32>(n, 1) * 4), n);
    const int sort_passes = (2 * ix.k + 7) / 8;
    u32* vals = (sort_passes & 1) ? vbuf_a.p : seg.positions.p;
    u32* vals_alt = (sort_passes & 1) ? seg.positions.p : vbuf_a.p;
    g_phases.start(st);
    if (c1 > c0) {
        extract_kernel<K><<<c1 - c0, CHUNK_THREADS, 0, st>>>(ref.packed.p, ref.mask.p, d_entries, d_chunk_entry, d_chunk_base,
                                                             c0, ix.k, ix.min_strand, keys, vals);
        KERNEL_CHECK();
    }
    g_phases.tick(0, st);
    radix_sort_pairs<K>(&keys, &vals, &keys_alt, &vals_alt, n, 2 * ix.k, rst, st);
    g_phases.tick(1, st);
    // CSR, pass 1: how many postings / terms start in every tile of the sorted tuples
    if (n && vals != seg.positions.p) throw BspError(-3, "radix sort ended in the scratch buffer");
    const u32 csr_tiles = (u32)ceil_div(n, CSR_TILE);
    flags.ensure((size_t)2 * csr_tiles + 2);              // tile counts | their exclusive scan
    u64* tile_counts = flags.p;
    u64* tile_offs = flags.p + csr_tiles + 1;
    if (n) {
        csr_count_kernel<K><<<csr_tiles, 256, 0, st>>>(keys, vals, n, seg.doc_tok.p, seg.n_docs, tile_counts);
        KERNEL_CHECK();
    }
    exclusive_scan<u64, u64>(tile_counts, tile_offs, csr_tiles, d_total.p, scan_tmp, st);
    u64 tot = 0;
    CUDA_CHECK(cudaMemcpyAsync(&tot, d_total.p, 8, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    seg.n_post = (u32)(tot & 0xFFFFFFFFu);
    seg.n_terms = (u32)(tot >> 32);
    // one allocation for the segment's dictionary and postings (a cudaMalloc costs ~2 ms with 100+ GB mapped)
    {
        auto pad = [](size_t b) { return (b + 255) & ~(size_t)255; };
        const size_t b_key = pad(((size_t)seg.n_terms + 1) * 8), b_off = pad(((size_t)seg.n_terms + 1) * 4);
        const size_t b_post = pad(((size_t)seg.n_post + 1) * 4);
        const size_t b_dir = ix.direct ? pad((size_t)4 << (2 * ix.k)) : 0;
        u8* q = ix.arena.take(b_key + b_off + 2 * b_post + b_dir);
        seg.term_key.adopt((u64*)q, (size_t)seg.n_terms + 1); q += b_key;
        seg.term_off.adopt((u32*)q, (size_t)seg.n_terms + 1); q += b_off;
        seg.post_doc.adopt((u32*)q, (size_t)seg.n_post + 1); q += b_post;
        seg.post_pos.adopt((u32*)q, (size_t)seg.n_post + 1); q += b_post;
        if (ix.direct) seg.direct.adopt((u32*)q, (size_t)1 << (2 * ix.k));
    }
    if (n) {
        csr_write_kernel<K><<<csr_tiles, 256, 0, st>>>(keys, vals, n, seg.doc_tok.p, seg.n_docs, tile_offs, seg.term_key.p,
                                                       seg.term_off.p, seg.post_doc.p, seg.post_pos.p);
        KERNEL_CHECK();
    }
    set_u32_kernel<<<1, 1, 0, st>>>(seg.term_off.p + seg.n_terms, seg.n_post);
    set_u32_kernel<<<1, 1, 0, st>>>(seg.post_pos.p + seg.n_post, n);
    g_phases.tick(2, st);
    // dictionary
    if (ix.direct) {
        u64 sz = 1ull << (2 * ix.k);
        fill_u32_kernel<<<(unsigned)std::min<u64>(ceil_div(sz, 256), 4096), 256, 0, st>>>(seg.direct.p, sz, NO_TERM);
        if (seg.n_terms)
            direct_table_kernel<<<(unsigned)ceil_div(seg.n_terms, 256), 256, 0, st>>>(seg.term_key.p, seg.n_terms, seg.direct.p);
    } else {
        u64 cap = 64;
        while (cap < 2ull * seg.n_terms) cap <<= 1;
        seg.hkeys.alloc(cap); seg.hvals.alloc(cap); seg.hmask = cap - 1;
        fill_u64_kernel<<<(unsigned)std::min<u64>(ceil_div(cap, 256), 4096), 256, 0, st>>>(seg.hkeys.p, cap, HASH_EMPTY);
        if (seg.n_terms)
            hash_insert_kernel<<<(unsigned)ceil_div(seg.n_terms, 256), 256, 0, st>>>(seg.term_key.p, seg.n_terms, seg.hkeys.p,
                                                                                      seg.hvals.p, seg.hmask);
    }
    KERNEL_CHECK();
    CUDA_CHECK(cudaStreamSynchronize(st));
    g_phases.tick(3, st);
}

// Build everything for the entries of `ref` (which the index keeps a pointer to).
static void build_gpu_index(GpuIndex& ix, RefStore& ref, const BuildOptions& opt, cudaStream_t st) {
    const auto build_t0 = std::chrono::steady_clock::now();
    const AllocStats alloc0 = g_alloc_stats;
    g_phases = BuildPhases();
    g_phases.on = getenv("BSP_BUILD_TIMING") != nullptr;
    ix.ref = &ref;
    ix.k = ref.k; ix.min_strand = ref.min_strand;
    const int k = ix.k;
    ix.direct = (2 * k <= 24);
    const i64 nE = (i64)ref.entries.size();
    // ---- chunk table
    std::vector<EntryDev> h_entries(nE + 1);
    std::vector<u32> h_chunk_entry;
    u64 n_chunks = 0;
    for (i64 e = 0; e < nE; e++) {
        const EntryMeta& m = ref.entries[e];
        if (m.len >= ((i64)1 << 32) - 64) throw BspError(-5, "entry longer than 2^32 bases is not supported");
        i64 nwin = m.len - k + 1;
        u64 nc = nwin > 0 ? (u64)ceil_div(nwin, CHUNK_WINDOWS) : 0;
        h_entries[e].base_off = (u64)m.base_off;
        h_entries[e].len = (u32)m.len;
        h_entries[e].first_chunk = (u32)n_chunks;
        h_entries[e].tok_fwd = 0; h_entries[e].tok_rev = 0xFFFFFFFFu;
        for (u64 c = 0; c < nc; c++) h_chunk_entry.push_back((u32)e);
        n_chunks += nc;
        if (n_chunks >= ((u64)1 << 32)) throw BspError(-5, "reference too large for one handle (chunk count)");
    }
    h_entries[nE].base_off = 0; h_entries[nE].len = 0; h_entries[nE].first_chunk = (u32)n_chunks;
    h_entries[nE].tok_fwd = 0; h_entries[nE].tok_rev = 0xFFFFFFFFu;
    DevBuf<EntryDev> d_entries; d_entries.alloc(nE + 1);
    DevBuf<u32> d_chunk_entry; d_chunk_entry.alloc(n_chunks + 1);
    DevBuf<u32> d_chunk_cnt; d_chunk_cnt.alloc(n_chunks + 1);
    DevBuf<u64> d_chunk_base; d_chunk_base.alloc(n_chunks + 1);
    DevBuf<u8> scan_tmp;
    DevBuf<u64> d_total; d_total.alloc(1);
    CUDA_CHECK(cudaMemcpyAsync(d_entries.p, h_entries.data(), (nE + 1) * sizeof(EntryDev), cudaMemcpyHostToDevice, st));
    if (n_chunks) {
        CUDA_CHECK(cudaMemcpyAsync(d_chunk_entry.p, h_chunk_entry.data(), n_chunks * 4, cudaMemcpyHostToDevice, st));
        chunk_count_kernel<<<(unsigned)n_chunks, CHUNK_THREADS, 0, st>>>(ref.mask.p, d_entries.p, d_chunk_entry.p, k, d_chunk_cnt.p);
        KERNEL_CHECK();
    }
    exclusive_scan<u32, u64>(d_chunk_cnt.p, d_chunk_base.p, n_chunks, d_total.p, scan_tmp, st);
    CUDA_CHECK(cudaMemcpyAsync(d_chunk_base.p + n_chunks, d_total.p, 8, cudaMemcpyDeviceToDevice, st));
    std::vector<u64> h_chunk_base(n_chunks + 1);
    CUDA_CHECK(cudaMemcpyAsync(h_chunk_base.data(), d_chunk_base.p, (n_chunks + 1) * 8, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    // ---- docs
    ix.docs.clear(); ix.doc_ntok.clear(); ix.doc_norm.clear();
    i64 total_tok = 0;
    for (i64 e = 0; e < nE; e++) {
        EntryMeta& m = ref.entries[e];
        m.ntok = (i64)(h_chunk_base[h_entries[e + 1].first_chunk] - h_chunk_base[h_entries[e].first_chunk]);
        int nd = (!ix.min_strand && m.rev_ok) ? 2 : 1;
        m.doc_fwd = (i32)ix.docs.size();
        ix.docs.push_back({(i32)e, ix.min_strand ? 2 : 0});
        m.doc_rev = -1;
        if (nd == 2) { m.doc_rev = (i32)ix.docs.size(); ix.docs.push_back({(i32)e, 1}); }
        for (int i = 0; i < nd; i++) {
            ix.doc_ntok.push_back((u32)m.ntok);
            ix.doc_norm.push_back((u8)h_norm_byte(m.ntok));
        }
        total_tok += m.ntok * nd;
    }
    const i64 D = (i64)ix.docs.size();
    ix.n_positions = total_tok;
    ix.sum_ttf = total_tok;
    // ---- residency decisions
    size_t free_b = 0, total_b = 0;
    CUDA_CHECK(cudaMemGetInfo(&free_b, &total_b));
    ix.d_pad = (u64)round_up(std::max<i64>(D, 1), 32);
    ix.row_bytes = ix.d_pad / 2;
    {
        // column order of the dense tables: stable counting sort of the docs by norm byte
        std::vector<u32> h_colmap((size_t)std::max<i64>(D, 1));
        ix.h_docmap.assign((size_t)ix.d_pad + 64, 0xFFFFFFFFu);          // (+ slack: vector loads of the last thread)
        u32 start[257] = {0};
        for (i64 d = 0; d < D; d++) start[ix.doc_norm[d] + 1]++;
        for (int b = 0; b < 256; b++) start[b + 1] += start[b];
        for (i64 d = 0; d < D; d++) { const u32 c = start[ix.doc_norm[d]]++; h_colmap[d] = c; ix.h_docmap[c] = (u32)d; }
        ix.colmap.alloc(h_colmap.size()); ix.docmap.alloc(ix.h_docmap.size());
        CUDA_CHECK(cudaMemcpyAsync(ix.colmap.p, h_colmap.data(), h_colmap.size() * 4, cudaMemcpyHostToDevice, st));
        CUDA_CHECK(cudaMemcpyAsync(ix.docmap.p, ix.h_docmap.data(), ix.h_docmap.size() * 4, cudaMemcpyHostToDevice, st));
        CUDA_CHECK(cudaStreamSynchronize(st));                           // h_colmap leaves scope
    }
    const bool can_tables = ix.direct;          // 2k <= 24: at most 5 * 4^12 rows
    ix.pair_row0 = 1ull << (2 * k);
    ix.n_rows = can_tables ? (ix.pair_row0 + (opt.need_pair_table ? (1ull << (2 * (k + 1))) : 0ull)) : 0ull;
    u64 tf_bytes = can_tables ? ix.pair_row0 * ix.row_bytes : 0;
    u64 pair_bytes = (can_tables && opt.need_pair_table) ? ((1ull << (2 * (k + 1))) * ix.row_bytes) : 0;
    ix.gap_cap = (u32)std::max<i64>(opt.gap_cap, 1);
    if (ix.n_rows + ((u64)1 << 28) >= ((u64)1 << 32)) throw BspError(-5, "dense table row count exceeds 2^32");   // (+ 2^28 gap clause rows)
    bool tables = false;
    if (can_tables && opt.want_tables != 2) {
        if (opt.want_tables == 1) tables = true;
        else tables = (double)(tf_bytes + pair_bytes) <= 0.40 * (double)free_b;
    }
    i64 seg_tokens = std::min<i64>(opt.seg_tokens, ((i64)1 << 31) - 8192);
    u64 key_bytes = (2 * k <= 32) ? 4 : 8;
    u64 temp_bytes = (u64)std::min<i64>(std::max<i64>(total_tok, 1), seg_tokens) * (2 * key_bytes + 8 + 8 + 8);
    // resident positional index: 4 B per token (positions) + 8 B per posting (doc, position offset; a posting needs a
    // (term, doc) pair: at most docs * 4^k of them) + per segment the dictionary and the term arrays
    const u64 n_seg_est = (u64)std::max<i64>(1, std::max<i64>(ceil_div(std::max<i64>(total_tok, 1), seg_tokens), ceil_div(D, SEG_MAX_DOCS)));
    const u64 terms_per_seg = ix.direct ? (1ull << (2 * k)) : (u64)std::min<i64>(std::max<i64>(total_tok, 1), seg_tokens);
    const u64 postings_est = ix.direct ? std::min<u64>((u64)total_tok, (u64)D * (1ull << (2 * k))) : (u64)total_tok;
    u64 pos_est = (u64)total_tok * 4 + postings_est * 8 + n_seg_est * terms_per_seg * (ix.direct ? 16 : 36);
    bool positional = true;
    if (opt.want_positional == 2) positional = false;
    else if (opt.want_positional == 0) {
        u64 avail = free_b > (tables ? tf_bytes + pair_bytes : 0) ? free_b - (tables ? tf_bytes + pair_bytes : 0) : 0;
        // (if the estimate turns out too low the build falls back to a non-resident positional index, see below)
        positional = (double)(pos_est + temp_bytes) <= 0.92 * (double)avail;
    }
    if (!tables && !positional) {
        if (opt.want_positional == 2 && opt.want_tables == 2) throw BspError(-1, "both dense_tables and positional are forced off");
        positional = true;    // something must be resident; let allocation fail loudly if it cannot fit
    }
    ix.tables = tables; ix.positional = positional; ix.has_pair = tables && opt.need_pair_table;
    DevBuf<u32> ex_row, ex_doc; DevBuf<float> ex_freq; DevBuf<unsigned long long> ex_count;
    ExcBuild exb{nullptr, nullptr, nullptr, nullptr, 0};
    if (tables) {
        ix.tab4.alloc(tf_bytes + pair_bytes + ix.row_bytes);          // + one all-zero row (index n_rows)
        CUDA_CHECK(cudaMemsetAsync(ix.tab4.p, 0, tf_bytes + pair_bytes + ix.row_bytes, st));
        ix.table_bytes = tf_bytes + pair_bytes + ix.row_bytes;
        // a cell can only be an exception where a term posting exists: <= tokens (term rows) + 4 * tokens (pair rows)
        exb.cap = (u64)std::min<i64>(opt.exc_cap, 5 * total_tok + 1024);
        ex_row.alloc((size_t)exb.cap); ex_doc.alloc((size_t)exb.cap); ex_freq.alloc((size_t)exb.cap);
        ex_count.alloc(1);
        CUDA_CHECK(cudaMemsetAsync(ex_count.p, 0, 8, st));
        exb.row = ex_row.p; exb.doc = ex_doc.p; exb.freq = ex_freq.p; exb.count = ex_count.p;
    }
    if (ix.direct) {
        u64 sz = 1ull << (2 * k);
        ix.df_dense.alloc(sz); ix.ttf_dense.alloc(sz);
        CUDA_CHECK(cudaMemsetAsync(ix.df_dense.p, 0, sz * 4, st));
        CUDA_CHECK(cudaMemsetAsync(ix.ttf_dense.p, 0, sz * 4, st));
    }
    // ---- segmentation (entries are never split; a segment holds >= 1 entry)
    ix.segs.clear();
    {
        i64 e = 0;
        u32 doc_base = 0;
        while (e < nE) {
            Segment seg;
            seg.first_entry = (i32)e; seg.doc_base = doc_base;
            i64 tok = 0; u32 nd = 0;
            while (e < nE) {
                const EntryMeta& m = ref.entries[e];
                int d = m.doc_rev >= 0 ? 2 : 1;
                if (seg.n_entries > 0 && (tok + m.ntok * d > seg_tokens || nd + d > SEG_MAX_DOCS)) break;
                if (m.ntok * d > ((i64)1 << 31) - 8192) throw BspError(-5, "a single entry has more than 2^31 tokens");
                tok += m.ntok * d; nd += d; seg.n_entries++; e++;
            }
            seg.n_docs = nd; seg.n_tok = (u32)tok;
            doc_base += nd;
            ix.segs.push_back(std::move(seg));
        }
    }
    // per-entry token starts (segment-relative) + per-segment doc_tok arrays
    size_t dt_total = 0;
    for (Segment& seg : ix.segs) dt_total += (size_t)seg.n_docs + 1;
    ix.doc_tok_all.alloc(dt_total);                    // one allocation + one copy for all segments
    std::vector<u32> dt_all(dt_total);
    size_t dt_at = 0;
    for (Segment& seg : ix.segs) {
        u32* dt = dt_all.data() + dt_at;
        u32 run = 0, dl = 0;
        for (i32 e = seg.first_entry; e < seg.first_entry + seg.n_entries; e++) {
            const EntryMeta& m = ref.entries[e];
            h_entries[e].tok_fwd = run; dt[dl++] = run; run += (u32)m.ntok;
            if (m.doc_rev >= 0) { h_entries[e].tok_rev = run; dt[dl++] = run; run += (u32)m.ntok; }
        }
        dt[dl] = run;
        seg.doc_tok.adopt(ix.doc_tok_all.p + dt_at, (size_t)seg.n_docs + 1);
        dt_at += (size_t)seg.n_docs + 1;
        ix.max_seg_docs = std::max(ix.max_seg_docs, seg.n_docs);
    }
    CUDA_CHECK(cudaMemcpyAsync(ix.doc_tok_all.p, dt_all.data(), dt_total * 4, cudaMemcpyHostToDevice, st));
    CUDA_CHECK(cudaMemcpyAsync(d_entries.p, h_entries.data(), (nE + 1) * sizeof(EntryDev), cudaMemcpyHostToDevice, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    // ---- per segment: extract -> sort -> CSR -> dictionary -> (tables) -> (drop)
    DevBuf<u8> kbuf_a, kbuf_b; DevBuf<u32> vbuf_a; DevBuf<u64> flags; RadixSortTemp rst;
    DevBuf<unsigned long long> d_counter; d_counter.alloc(1);
    CUDA_CHECK(cudaMemsetAsync(d_counter.p, 0, 8, st));
    i64 postings = 0;
    std::vector<SegDev> views;
    DevBuf<SegDev> d_prev;
    const bool may_drop_positional = positional && opt.want_positional == 0 && tables && ix.direct;
    for (size_t si = 0; si < ix.segs.size(); si++) {
        Segment& seg = ix.segs[si];
        u32 c0 = h_entries[seg.first_entry].first_chunk, c1 = h_entries[seg.first_entry + seg.n_entries].first_chunk;
        try {
            if (2 * k <= 32)
                build_segment<u32>(ix, seg, d_entries.p, d_chunk_entry.p, d_chunk_base.p, c0, c1, kbuf_a, kbuf_b, vbuf_a,
                                   flags, rst, scan_tmp, d_total, st);
            else
                build_segment<u64>(ix, seg, d_entries.p, d_chunk_entry.p, d_chunk_base.p, c0, c1, kbuf_a, kbuf_b, vbuf_a,
                                   flags, rst, scan_tmp, d_total, st);
        } catch (const BspError& e) {
            // out of device memory with the positional index resident: give it up (the dense tables stay complete for
            // the segments already processed) and redo this segment
            if (e.code != -6 || !positional || !may_drop_positional) throw;
            cudaGetLastError();
            for (size_t j = 0; j < si; j++) {
                Segment& o = ix.segs[j];
                o.direct.release(); o.term_key.release(); o.term_off.release(); o.post_doc.release();
                o.post_pos.release(); o.positions.release(); o.doc_tok.release();
            }
            ix.arena.release_all();
            positional = false; ix.positional = false; ix.positional_bytes = 0; views.clear();
            ix.notes += "positional index dropped: out of device memory; ";
            fprintf(stderr, "WARNING biospectra_gpu: out of device memory, the positional index was dropped; reads whose clauses "
                            "need a slop the dense tables do not hold (reads with N, kmer_skips > 0) will be reported with status "
                            "error instead of being classified (set positional=1 in bsp_config to fail open() instead)\n");
            si--;
            continue;
        }
        postings += seg.n_post;
        SegDev v = seg.view();
        g_phases.start(st);
        if (ix.direct) {
            if (seg.n_terms) accumulate_df_kernel<<<(unsigned)ceil_div(seg.n_terms, 256), 256, 0, st>>>(v, ix.df_dense.p, ix.ttf_dense.p);
        } else {
            if (!views.empty()) {
                d_prev.ensure(views.size());
                CUDA_CHECK(cudaMemcpyAsync(d_prev.p, views.data(), views.size() * sizeof(SegDev), cudaMemcpyHostToDevice, st));
            }
            if (seg.n_terms)
                count_new_terms_kernel<<<(unsigned)ceil_div(seg.n_terms, 256), 256, 0, st>>>(v, d_prev.p, (int)views.size(), d_counter.p);
        }
        if (tables && seg.n_terms) {
            tf_table4_kernel<<<(unsigned)ceil_div((u64)seg.n_terms * 32, 256), 256, 0, st>>>(v, ix.colmap.p, (u32*)ix.tab4.p, ix.row_bytes / 4, exb);
            if (ix.has_pair) {
                u64 rows = 1ull << (2 * (k + 1));
                if (seg.n_docs <= POSW_DOCS && !getenv("BSP_PAIR_TABLE_DIRECT")) {
                    // few docs per segment: warp per first k-mer, slices staged in shared memory
                    const int dcap = (int)round_up(std::max<i64>(seg.n_docs, 1), 32);
                    const size_t smem = PT_WARPS * pair_stage_warp_bytes(dcap, PW_POS);
                    CUDA_CHECK(cudaFuncSetAttribute(pair_table4_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                    (int)(PT_WARPS * pair_stage_warp_bytes(POSW_DOCS, PW_POS))));
                    pair_table4_staged_kernel<<<(unsigned)ceil_div(rows / 4, PT_WARPS), PT_WARPS * 32, smem, st>>>(
                        v, ix.colmap.p, k, ix.min_strand, (u32*)ix.tab4.p, ix.row_bytes / 4, ix.pair_row0, exb, rows / 4, dcap, PW_POS);
                } else {
                    pair_table4_kernel<<<(unsigned)ceil_div(rows * 32, 256), 256, 0, st>>>(v, ix.colmap.p, k, ix.min_strand, (u32*)ix.tab4.p,
                                                                                          ix.row_bytes / 4, ix.pair_row0, exb, rows);
                }
            }
        }
        KERNEL_CHECK();
        CUDA_CHECK(cudaStreamSynchronize(st));
        g_phases.tick(4, st);
        if (positional) { views.push_back(v); ix.positional_bytes += seg.bytes(); }
        else {
            if (!ix.direct) throw BspError(-5, "hash-mode index (k > 12) requires the positional index");
            seg.direct.release(); seg.term_key.release(); seg.term_off.release(); seg.post_doc.release();
            seg.post_pos.release(); seg.positions.release(); seg.doc_tok.release();
            ix.arena.reset();
        }
    }
    if (!positional) ix.arena.release_all();
    ix.n_postings = postings;
    if (ix.direct) {
        u64 sz = 1ull << (2 * k);
        count_nonzero_kernel<<<(unsigned)ceil_div(sz, 256), 256, 0, st>>>(ix.df_dense.p, sz, d_counter.p);
    }
    unsigned long long cnt = 0;
    CUDA_CHECK(cudaMemcpyAsync(&cnt, d_counter.p, 8, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    ix.n_terms = (i64)cnt;
    if (positional) {
        ix.d_segs.alloc(views.size() + 1);
        if (!views.empty())
            CUDA_CHECK(cudaMemcpyAsync(ix.d_segs.p, views.data(), views.size() * sizeof(SegDev), cudaMemcpyHostToDevice, st));
    }
    if (tables) {
        // exception list: sort by (row, doc), gather, per-row offsets
        unsigned long long ne = 0;
        CUDA_CHECK(cudaMemcpyAsync(&ne, ex_count.p, 8, cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaStreamSynchronize(st));
        if (ne > (unsigned long long)exb.cap)
            throw BspError(-6, "dense table exception list overflow (" + std::to_string(ne) + " entries); raise BSP_EXC_CAP");
        ix.n_exc = (i64)ne;
        const u32 n = (u32)ne;
        ix.exc_off.alloc(ix.n_rows + 2);
        // the tail of the exception arrays holds the gap-clause match lists of the two pipeline slots
        ix.gap_tail = (size_t)ix.gap_cap * GAP_LIST;
        const size_t tail = (size_t)BSP_SLOTS * ix.gap_tail;                 // one gap-list region per pipeline slot
        if ((u64)n + tail >= ((u64)1 << 32)) throw BspError(-5, "exception + gap lists exceed 2^32 entries");
        ix.exc_doc.alloc((size_t)n + tail); ix.exc_freq.alloc((size_t)n + tail);
        DevBuf<u64> ka, kb; DevBuf<u32> va, vb;
        ka.alloc(std::max<size_t>(n, 1)); kb.alloc(std::max<size_t>(n, 1)); va.alloc(std::max<size_t>(n, 1)); vb.alloc(std::max<size_t>(n, 1));
        u64* keys = ka.p; u64* keys_alt = kb.p; u32* vals = va.p; u32* vals_alt = vb.p;
        if (n) {
            exc_keys_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(ex_row.p, ex_doc.p, n, keys, vals);
            int row_bits = 1;
            while ((1ull << row_bits) < ix.n_rows) row_bits++;
            radix_sort_pairs<u64>(&keys, &vals, &keys_alt, &vals_alt, n, 32 + row_bits, rst, st);
            exc_gather_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(keys, vals, ex_freq.p, n, ix.exc_doc.p, ix.exc_freq.p);
        }
        exc_offsets_kernel<<<(unsigned)ceil_div(ix.n_rows + 1, 256), 256, 0, st>>>(keys, n, ix.n_rows, ix.exc_off.p);
        KERNEL_CHECK();
        CUDA_CHECK(cudaStreamSynchronize(st));
        ix.table_bytes += ix.exc_off.bytes() + ix.exc_doc.bytes() + ix.exc_freq.bytes();
    }
    CUDA_CHECK(cudaStreamSynchronize(st));
    if (g_phases.on)
        fprintf(stderr, "[bsp build] %.1f ms total; cudaMalloc %.1f ms in %ld calls, cudaFree %.1f ms in %ld calls; %zu segments; "
                "phases: extract %.1f, sort %.1f, csr %.1f, dictionary %.1f, tables %.1f ms\n",
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - build_t0).count(),
                g_alloc_stats.alloc_ms - alloc0.alloc_ms, g_alloc_stats.allocs - alloc0.allocs,
                g_alloc_stats.free_ms - alloc0.free_ms, g_alloc_stats.frees - alloc0.frees, ix.segs.size(),
                g_phases.ms[0], g_phases.ms[1], g_phases.ms[2], g_phases.ms[3], g_phases.ms[4]);
}
