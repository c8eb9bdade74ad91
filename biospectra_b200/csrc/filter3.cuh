// filter3.cuh -- the bound-and-rescore filter as a persistent, barrier-free stream (tf-idf; docs within one CTA's reach).
//
// filter_score_kernel (dense4.cuh) is one CTA per read: set-up, row loop, estimates, selection, candidate pass,
// rescoring and top-10 follow each other inside the CTA behind __syncthreads(), and a warp spends two thirds of its life
// outside the row loop, mostly waiting (profiles/README.md, round 2 source page).  Here the same arithmetic is cut into
// four kernels so that the row stream never drains and no warp waits for another warp's serial phase:
//
//   filter3_prep_kernel     warp per read: the read's clause records (byte LUT of quantised weights, table row offset), its
//                           exception cells already quantised, the bound parameters -- everything the stream needs.
//   filter3_link_kernel     warp per read: for every row the table offset of the row FK3_STAGES further on in the
//                           *worker's* row sequence (reads are dealt round-robin to the persistent CTAs): the ring runs on
//                           across read boundaries without anybody looking anything up.
//   filter3_stream_kernel   persistent CTAs (grid = resident CTAs of the device), warp = 1,024 docs, lane = 32 docs.  Whole
//                           a warp's 512-byte slice of each table row arrives by cp.async.bulk (one elected lane, 2 rows
//                           per mbarrier) in the warp's own ring, refilled the moment the warp has consumed a slot -- 8 rows
//                           in flight per warp at all times, no warp ever waits for another one's rows (a CTA-wide ring of
//                           whole rows was measured first: the slowest warp gates every refill, 38 ms); the records of the
//                           read after next arrive by bulk copy in a CTA-wide double buffer, fetched by the last warp to
//                           leave the buffer's previous read (a shared-memory counter tells it so).  A warp waits on
//                           mbarriers only: for its rows and for its read's records.  Per read it streams the rows through the packed byte-LUT
//                           lanes (pair_word2 / term_word2), folds the exception cells, takes the estimates of its 1,024 docs
//                           and appends the docs that can still reach the result -- against the slab's own best estimate, a
//                           valid (weaker) threshold -- to the read's candidate list; the best estimate goes to tau[read].
//   filter3_rescore_kernel  warp per read: prunes the candidates against the read-wide tau, rescores the survivors with
//                           Lucene's exact float operations (double accumulation) and writes the top-10 part.
//
// Exactness is unchanged: a doc can equal the maximum only if E(d)(1+delta) >= tau(1-delta); every
// slab applies that test with tau_slab <= tau, so the union of the slab lists is a superset of the docs the read-wide test
// keeps, and the rescoring kernel applies the read-wide test before it spends loads on a candidate.
#pragma once
#include "dense4.cuh"

#define FK3_STAGES 8               // rows in flight per CTA (4 mbarrier-guarded slots of 2 rows)
#define FK3_PAIRS (FK3_STAGES / 2)
#define FK3_CELLS 32               // exception cells of short lists kept per read (more: the lists are searched per lane)
#define FK3_LONG 8                 // rows with long exception lists per read (more: exact_score_kernel takes the read)
#define FK3_MAXW 16                // warps (1,024-doc slabs) per CTA: 16,384 docs
#define FK3_MAXROWS 256            // rows of a read (255 clauses + the dummy row that pads an odd count)
#define FK3_SKIP 1u                // header flag: read handed to exact_score_kernel by the prep kernel
// how a warp's row slices travel: 1 = cp.async.bulk by one elected lane + an mbarrier per slot (UBLKCP / SYNCS),
// 0 = 16-byte cp.async per lane + commit groups (LDGSTS), the mover of filter_score_kernel.  Measured at C2: see profiles/README.md
#ifndef FK3_ROWS_BULK
#define FK3_ROWS_BULK 0
#endif

struct __align__(16) F3Hdr {       // 48 bytes per read
    i32 n_two;                     // rows [0, n_two) go through the two-row byte-LUT loop
    i32 n_rows;                    // rows the stream fetches for this read (even; a dummy row pads an odd count)
    i32 n_real;                    // rows that carry data: [n_two, n_real) are single rows (odd pair row, Term rows)
    i32 n_dense;                   // single rows below this index are pair rows, the others Term rows
    i32 nc_need;                   // nc | need << 16
    float theta_mul;               // theta = tau * theta_mul ; 0 when the bound is void (q_min < 2)
    u32 misc;                      // n_cells | n_long << 8 | flags << 16
    u32 budget;                    // what a lane's u16 sums can still take from cells of long lists (65535 - bound so far)
    // written by the link kernel: the read two places further on in the worker's sequence (-1: none), what to copy for it
    i32 nx_read;
    u32 nx_sizes;                  // n_rows | n_cells << 16
    i64 nx_rec;                    // its first record slot
};
struct __align__(16) F3Cell { u32 doc, qe, clause; float freq; };      // doc = table column
struct __align__(16) F3Long { u32 a, n, clause; float sv; };

// first record slot of read r: one slot per token + two spare per read (the dummy row; an even base keeps the 8-byte
// high-LUT array 16-byte aligned for the bulk copy)
__host__ __device__ __forceinline__ i64 f3_rec_base(i64 tok_off_r, i64 r) { return (tok_off_r + 2 * r + 1) & ~(i64)1; }

__device__ __forceinline__ u32 lanemask_lt() { u32 m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }

// ---- mbarrier / bulk-copy primitives (sm_90+: SYNCS / UBLKCP in SASS)
__device__ __forceinline__ void mbar_init(u32 bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(u32 bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u32 bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(u32 bar, u32 parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
// arrival counter in shared memory: returns the value before the add (acq_rel at CTA scope: the reads of the warps that
// arrived earlier happen before whatever the last arriver does next)
__device__ __forceinline__ u32 smem_arrive(u32 addr) {
    u32 old;
    asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(addr) : "memory");
    return old;
}
__device__ __forceinline__ void sts32(u32 addr, u32 v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
// generic-proxy accesses of shared memory (the warps' reads of a slot) before async-proxy writes (the bulk copy into it)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_g2s(u32 dst, const void* src, u32 bytes, u32 bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// ------------------------------------------------------------------ prep: warp per read
template <bool ALLTERMS>
__global__ void __launch_bounds__(256) filter3_prep_kernel(
        Tab4Dev T, const float* __restrict__ tf16, const i32* __restrict__ read_list, i32 n_list,
        const i64* __restrict__ tok_off, const i32* __restrict__ n_tok, const float* __restrict__ clause_val,
        const u32* __restrict__ clause_row, const i32* __restrict__ n_clauses, const i32* __restrict__ msm,
        const float2* __restrict__ read_vrange, const i32* __restrict__ read_ngap, const GapClause* __restrict__ gaps, int alg,
        F3Hdr* __restrict__ hdr, uint4* __restrict__ rec, uint2* __restrict__ rec_hi, F3Cell* __restrict__ cells,
        F3Long* __restrict__ longs, i32* __restrict__ cand_n, u32* __restrict__ tau, i32* __restrict__ long_overflow,
        i32* __restrict__ route, i32* __restrict__ overflow_list, i32* __restrict__ overflow_count,
        FilterStats* __restrict__ stats) {
    const int lane = threadIdx.x & 31;
    const i32 j = (i32)((blockIdx.x * (u64)blockDim.x + threadIdx.x) >> 5);
    if (j >= n_list) return;
    const i32 r = read_list ? read_list[j] : j;
    const i64 base = tok_off[r];
    const i64 rb = f3_rec_base(base, r);
    const i32 n = n_tok[r], nc = n_clauses[r];
    const i32 n_pair = alg == ALG_NAIVE ? 0 : (alg == ALG_CHAIN ? nc : n / 2);
    const i32 need = max(1, msm[r]);
    const i32 n_gap = read_ngap[r];
    const i32 n_dense = n_pair - n_gap, n_stream = nc - n_gap;
    const float2 vr = read_vrange[r];
    const float S = ALLTERMS ? 127.0f / (vr.x * tf16[T4_TF_MAX]) : FK_QMAX / (vr.x * tf16[T4_PAIR_MAX]);
    u32 qmin = (u32)__float2int_rn(S * vr.y);
    u32 exc_sum = 0, n_cells = 0, n_long = 0;
    bool overflow = false;
    i32 gap_seen = 0, dense_seen = 0;
    const u32 row16 = (u32)(T.row_bytes >> 4);
    for (i32 cb = 0; cb < nc; cb += 32) {
        const i32 c0 = cb + lane;
        const bool valid = c0 < nc;
        const float v = valid ? clause_val[base + c0] : 0.0f;
        const u32 row = valid ? clause_row[base + c0] : 0u;
        const bool is_pair = valid && c0 < n_pair;
        const bool is_gap = is_pair && row >= T.temp_row0;
        const u32 bg = __ballot_sync(0xFFFFFFFFu, is_gap), bd = __ballot_sync(0xFFFFFFFFu, is_pair && !is_gap);
        // streamed order: table-backed pair rows, Term rows, then (not streamed) the gap clauses
        i32 c = c0 - n_gap;
        if (is_pair) c = is_gap ? nc - 1 - (gap_seen + __popc(bg & lanemask_lt())) : dense_seen + __popc(bd & lanemask_lt());
        gap_seen += __popc(bg); dense_seen += __popc(bd);
        const float sv = S * v;
        if (valid && c < n_stream) {
            uint4 rc;
            rc.z = 0u;
            rc.w = row * row16;
            if (is_pair) {
                u32 q[8];
                q[0] = 0;
#pragma unroll
                for (int t = 1; t < 8; t++) q[t] = min((u32)__float2int_rn(sv * tf16[t]), 255u);
                rc.x = q[0] | (q[1] << 8) | (q[2] << 16) | (q[3] << 24);
                rc.y = q[4] | (q[5] << 8) | (q[6] << 16) | (q[7] << 24);
            } else if (ALLTERMS) {
                u32 q[16];
                q[0] = 0; q[15] = 0;
#pragma unroll
                for (int t = 1; t < 15; t++) q[t] = min((u32)__float2int_rn(sv * tf16[t]), 127u);
                rc.x = q[0] | (q[1] << 8) | (q[2] << 16) | (q[3] << 24);
                rc.y = q[4] | (q[5] << 8) | (q[6] << 16) | (q[7] << 24);
                rec_hi[rb + c] = make_uint2(q[8] | (q[9] << 8) | (q[10] << 16) | (q[11] << 24), q[12] | (q[13] << 8) | (q[14] << 16) | (q[15] << 24));
            } else {                                    // Term row of a mixed read: scaled value for the nibble path
                rc.x = __float_as_uint(sv);
                rc.y = 0u;
            }
            rec[rb + c] = rc;
        }
        // exception cells of the clause's row (cells the nibble cannot hold; every match of a gap clause)
        u32 a = 0, ne = 0;
        if (valid) row_exceptions(T, row, a, ne);
        const bool is_short = ne > 0 && ne <= FK_EXS_ROW;
        u32 pre = is_short ? ne : 0u;                    // exclusive prefix over the lanes
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_up_sync(0xFFFFFFFFu, pre, o); if (lane >= o) pre += t; }
        pre -= is_short ? ne : 0u;
        bool as_long = ne > FK_EXS_ROW;
        u32 endpos = 0;
        if (is_short) {
            if (n_cells + pre + ne <= FK3_CELLS) {
                endpos = pre + ne;
                for (u32 x = 0; x < ne; x++) {
                    const float f = __ldg(T.exc_freq + a + x);
                    const u32 qe = (u32)min(__float2int_rn(sv * tf_of(f)), 1 << 20);
                    F3Cell cl; cl.doc = __ldg(T.exc_doc + a + x); cl.qe = qe; cl.clause = (u32)c0; cl.freq = f;
                    cells[(i64)r * FK3_CELLS + n_cells + pre + x] = cl;
                    qmin = min(qmin, qe);
                    exc_sum += qe;
                }
            } else as_long = true;                       // no room: searched per lane like a long list
        }
        n_cells += warp_max_u32(endpos);                 // lists are stored in lane order: the fitting ones form a prefix
        const u32 bl = __ballot_sync(0xFFFFFFFFu, as_long);
        if (as_long) {
            const u32 idx = n_long + __popc(bl & lanemask_lt());
            if (idx < FK3_LONG) { F3Long lg; lg.a = a; lg.n = ne; lg.clause = (u32)c0; lg.sv = sv; longs[(i64)r * FK3_LONG + idx] = lg; }
            // q_min must cover these cells too (delta = 0.51 / q_min holds for every doc).  Their frequencies are not
            // read here: a phrase frequency is a sum of 1/(matchLength+1), matchLength <= slop, so it is at least
            // 1/(slop+1) (slop 2 for a table row, the clause's own for a gap clause); a Term row's are >= 15.
            if (is_pair) {
                const int slop = is_gap ? gaps[row - T.temp_row0].slop : 2;
                qmin = min(qmin, (u32)(sv * tf_of(1.0f / (float)(slop + 1)) * 0.99999f));
            }
        }
        n_long += __popc(bl);
    }
    if (n_long > FK3_LONG) overflow = true;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        qmin = min(qmin, __shfl_xor_sync(0xFFFFFFFFu, qmin, o));
        exc_sum += __shfl_xor_sync(0xFFFFFFFFu, exc_sum, o);
    }
    // u16 weight lanes: pair-row weights <= 255, Term-row weights <= 255*tf(14)/tf(7) < 362 (all-term reads: <= 127), plus
    // the short-list cells (all of them, as if they met in one doc); long lists are checked per lane by the stream kernel
    const u32 lane_max = ALLTERMS ? 127u * (u32)nc : 255u * (u32)n_pair + 362u * (u32)(nc - n_pair);
    if (lane_max + exc_sum > 65535u) overflow = true;
    const i32 n_rows = (n_stream + 1) & ~1;
    if (lane == 0) {
        if (n_rows > n_stream) { uint4 rc; rc.x = 0; rc.y = 0; rc.z = 0; rc.w = T.zero_row * row16; rec[rb + n_stream] = rc; }
        F3Hdr h;
        const i32 n_two_src = ALLTERMS ? n_stream : n_dense;
        h.n_two = overflow ? 0 : (n_two_src & ~1);
        h.n_rows = overflow ? 0 : n_rows;
        h.n_real = overflow ? 0 : n_stream;
        h.n_dense = n_dense;
        h.nc_need = nc | (need << 16);
        const float qf = (float)qmin;
        const float delta = qf >= 2.0f ? FK_ROUND_SLACK / qf + FK_REL_SLACK : 1.0f;
        h.theta_mul = delta < 1.0f ? ((1.0f - delta) / (1.0f + delta)) * (1.0f - FK_REL_SLACK) : 0.0f;
        h.misc = n_cells | (min(n_long, (u32)FK3_LONG) << 8) | ((overflow ? FK3_SKIP : 0u) << 16);
        h.budget = overflow ? 0u : 65535u - (lane_max + exc_sum);
        h.nx_read = -1; h.nx_sizes = 0u; h.nx_rec = 0;
        hdr[r] = h;
        cand_n[r] = 0;
        tau[r] = 0u;
        long_overflow[r] = 0;
        if (overflow) {
            route[r] = ROUTE_DENSE;
            overflow_list[atomicAdd(overflow_count, 1)] = r;
            if (stats) atomicAdd(&stats->overflow_reads, 1ull);
        } else if (stats && n_cells) atomicAdd(&stats->forced, (unsigned long long)n_cells);
    }
}

// ------------------------------------------------------------------ link: the row FK3_STAGES ahead in the worker's sequence
// Worker w of G streams the reads at list positions w, w + G, w + 2G, ... one after the other; record c of read j learns
// the table offset of the row FK3_STAGES further on in that flattened sequence (zero_row at the very end: a harmless
// fetch), and the first FK3_STAGES rows of every worker go to prolog[].  Warp per read.
__device__ __forceinline__ u32 f3_row_ahead(const F3Hdr* __restrict__ hdr, const uint4* __restrict__ rec, const i32* __restrict__ read_list,
                                            i32 n_list, const i64* __restrict__ tok_off, i32 j, i32 c, i32 G, u32 zero_off) {
    // flattened row (j, c) with c possibly >= n_rows(j): walk on through the worker's reads
    while (j < n_list) {
        const i32 r = read_list ? read_list[j] : j;
        const i32 nr = hdr[r].n_rows;
        if (c < nr) return rec[f3_rec_base(tok_off[r], r) + c].w;
        c -= nr;
        j += G;
    }
    return zero_off;
}
__global__ void __launch_bounds__(256) filter3_link_kernel(F3Hdr* __restrict__ hdr, uint4* __restrict__ rec,
                                                            const i32* __restrict__ read_list, i32 n_list,
                                                            const i64* __restrict__ tok_off, i32 G, u32 zero_off,
                                                            u32* __restrict__ prolog) {
    const int lane = threadIdx.x & 31;
    const i32 j = (i32)((blockIdx.x * (u64)blockDim.x + threadIdx.x) >> 5);
    if (j >= n_list) return;
    const i32 r = read_list ? read_list[j] : j;
    const i32 nr = hdr[r].n_rows;
    uint4* recs = rec + f3_rec_base(tok_off[r], r);
    for (i32 c = lane; c < nr; c += 32)
        recs[c].z = c + FK3_STAGES < nr ? recs[c + FK3_STAGES].w
                                        : f3_row_ahead(hdr, rec, read_list, n_list, tok_off, j + G, c + FK3_STAGES - nr, G, zero_off);
    if (j < G && lane < FK3_STAGES) prolog[j * FK3_STAGES + lane] = f3_row_ahead(hdr, rec, read_list, n_list, tok_off, j, lane, G, zero_off);
    if (lane == 0 && (i64)j + 2 * (i64)G < n_list) {
        const i32 jn = j + 2 * G;
        const i32 rn = read_list ? read_list[jn] : jn;
        hdr[r].nx_read = rn;
        hdr[r].nx_sizes = (u32)hdr[rn].n_rows | ((hdr[rn].misc & 0xFFu) << 16);
        hdr[r].nx_rec = f3_rec_base(tok_off[rn], rn);
    }
}

// ------------------------------------------------------------------ stream
// 10th largest of the 32 lanes' values (bit patterns of non-negative floats); 0 when fewer than 10 are non-zero
__device__ __forceinline__ u32 warp_tenth_largest_u32(u32 v) {
    u32 mine = v, res = 0;
#pragma unroll 1
    for (int rnd = 0; rnd < TOPN; rnd++) {
        res = warp_max_u32(mine);
        // remove ONE holder of the maximum per round (equal estimates are distinct docs)
        const u32 holders = __ballot_sync(0xFFFFFFFFu, mine == res && res != 0u);
        if (holders && (int)(threadIdx.x & 31) == __ffs((int)holders) - 1) mine = 0;
    }
    return res;
}

// shared memory of the stream kernel (dynamic): ring | records[2] | high LUT halves[2] (NAIVE_KMER) | cells[2] | headers[2] |
// park slots | mbarriers
__host__ __device__ inline size_t f3_smem_bytes(int W, bool allterms) {
    return (size_t)FK3_STAGES * W * 512 + 2 * FK3_MAXROWS * 16 + (allterms ? 2 * FK3_MAXROWS * 8 : 0) + 2 * FK3_CELLS * 16 + 2 * sizeof(F3Hdr) +
           (size_t)W * 6 * 16 + ((size_t)W * FK3_PAIRS + 2) * 8 + 2 * 4 + 8;
}

template <bool ALLTERMS>
__global__ void __launch_bounds__(FK_MAXT, FK_MINB) filter3_stream_kernel(
        Tab4Dev T, const F3Hdr* __restrict__ hdr, const uint4* __restrict__ rec, const uint2* __restrict__ rec_hi,
        const F3Cell* __restrict__ cells, const F3Long* __restrict__ longs, const u32* __restrict__ prolog,
        const i32* __restrict__ read_list, i32 n_list, const i64* __restrict__ tok_off, const float* __restrict__ tf16,
        const float* __restrict__ col_w, const float2* __restrict__ group_w /* per 32 columns: (min, max) weight */,
        int cand_cap, i32* __restrict__ cand_n, uint2* __restrict__ cand, u32* __restrict__ tau,
        i32* __restrict__ long_overflow /* per read: a lane's u16 sums overflowed */, int full_topn, u32 one) {
    extern __shared__ __align__(128) u8 f3_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = (int)(blockDim.x >> 5);
    const i32 worker = (i32)blockIdx.x, G = (i32)gridDim.x;
    if (worker >= n_list) return;
    const u32 slot_bytes = (u32)W * 512u;                 // one row slot: 16 bytes per thread (>= row_bytes)
    u8* p = f3_smem;
    const u32 ring_sa = (u32)__cvta_generic_to_shared(p);            p += (size_t)FK3_STAGES * slot_bytes;
    uint4* s_rec = reinterpret_cast<uint4*>(p);                       p += 2 * FK3_MAXROWS * 16;
    uint2* s_rec_hi = reinterpret_cast<uint2*>(p);                    p += ALLTERMS ? 2 * FK3_MAXROWS * 8 : 0;
    F3Cell* s_cells = reinterpret_cast<F3Cell*>(p);                   p += 2 * FK3_CELLS * 16;
    F3Hdr* s_hdr = reinterpret_cast<F3Hdr*>(p);                       p += 2 * sizeof(F3Hdr);
    uint4* park = reinterpret_cast<uint4*>(p) + (size_t)warp * 6;    p += (size_t)W * 6 * 16;
    const u32 bar_sa = (u32)__cvta_generic_to_shared(p);
    // mbarriers rec_full[2], full[warp][pair slot]; arrival counters of the two record buffers
    const u32 recf0 = bar_sa, full0 = bar_sa + 16 + (u32)warp * FK3_PAIRS * 8, rcnt0 = bar_sa + 16 + (u32)W * FK3_PAIRS * 8;
    const u32 row_bytes = (u32)T.row_bytes;
    const u32 d0 = (u32)threadIdx.x * FK_DOCS;
    const bool active = (u64)d0 < T.row_bytes * 2;
    // the warp's ring: FK3_STAGES slots of 512 bytes (its slice of a row: 16 bytes per lane)
    const u32 wring_sa = ring_sa + (u32)warp * (FK3_STAGES * 512u);
    const u32 my_sa = wring_sa + (u32)lane * 16u;         // this thread's 16 bytes of slot 0
    const u32 seg_bytes = min(512u, row_bytes > (u32)warp * 512u ? row_bytes - (u32)warp * 512u : 0u);   // (a multiple of 16)
    const u8* wtab = T.tab + (u64)warp * 512u;            // the warp's slice of row 0
    (void)wtab;

    // header, records and cells of read rn into buffer b, signalled on rec_full[b]
    auto load_records = [&](i32 rn, u32 nrec, u32 ncell, i64 rbn, u32 b) {
        const u32 bytes = (u32)sizeof(F3Hdr) + nrec * 16u + (ALLTERMS ? nrec * 8u : 0u) + ncell * 16u;      // (nrec is even)
        mbar_expect_tx(recf0 + b * 8, bytes);
        bulk_g2s((u32)__cvta_generic_to_shared(s_hdr + b), hdr + rn, (u32)sizeof(F3Hdr), recf0 + b * 8);
        if (nrec) {
            bulk_g2s((u32)__cvta_generic_to_shared(s_rec + b * FK3_MAXROWS), rec + rbn, nrec * 16u, recf0 + b * 8);
            if (ALLTERMS) bulk_g2s((u32)__cvta_generic_to_shared(s_rec_hi + b * FK3_MAXROWS), rec_hi + rbn, nrec * 8u, recf0 + b * 8);
        }
        if (ncell) bulk_g2s((u32)__cvta_generic_to_shared(s_cells + b * FK3_CELLS), cells + (i64)rn * FK3_CELLS, ncell * 16u, recf0 + b * 8);
    };
    if (lane == 0) {
        for (int s = 0; s < FK3_PAIRS; s++) mbar_init(full0 + s * 8, 1);
        if (warp == 0) for (int b = 0; b < 2; b++) { mbar_init(recf0 + b * 8, 1); sts32(rcnt0 + b * 4, 0u); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();                                      // the only CTA-wide barrier of the kernel
    if (threadIdx.x == 0) {
        // prologue: the records of the worker's first two reads and the first FK3_STAGES rows of its sequence
        for (int b = 0; b < 2; b++) {
            const i64 jn = (i64)worker + (i64)b * G;
            if (jn < n_list) {
                const i32 rn = read_list ? read_list[jn] : (i32)jn;
                load_records(rn, (u32)hdr[rn].n_rows, hdr[rn].misc & 0xFFu, f3_rec_base(tok_off[rn], rn), (u32)b);
            }
        }
    }
    // the first FK3_STAGES rows of the worker's sequence into this warp's ring
    const u8* ltab = T.tab + (u64)(active ? d0 : 0u) / 2;    // this lane's 16 bytes of row 0 (idle lanes of the last warp: lane 0's)
    asm volatile("" : "+l"(ltab));
#if FK3_ROWS_BULK
    if (lane == 0 && seg_bytes) {
#pragma unroll
        for (int s = 0; s < FK3_PAIRS; s++) {
            mbar_expect_tx(full0 + s * 8, 2u * seg_bytes);
            bulk_g2s(wring_sa + (u32)(2 * s) * 512u, wtab + (u64)prolog[worker * FK3_STAGES + 2 * s] * 16, seg_bytes, full0 + s * 8);
            bulk_g2s(wring_sa + (u32)(2 * s + 1) * 512u, wtab + (u64)prolog[worker * FK3_STAGES + 2 * s + 1] * 16, seg_bytes, full0 + s * 8);
        }
    }
#else
#pragma unroll
    for (int st = 0; st < FK3_STAGES; st++) {
        cp_async16_sa(my_sa + (u32)st * 512u, ltab + (u64)prolog[worker * FK3_STAGES + st] * 16);
        if (st & 1) cp_async_commit();
    }
#endif
    const u32 cnt_lo = 0x01010100u * one;
    u32 it = 0;                                           // row pairs consumed so far (the same sequence in every warp)
    u32 jc = 0;                                           // reads started so far
#pragma unroll 1
    for (i32 j = worker; j < n_list; j += G, jc++) {
        const u32 b = jc & 1u;
        const i32 r_cur = read_list ? read_list[j] : j;
        mbar_wait(recf0 + b * 8, (jc >> 1) & 1u);         // this read's records have landed
        const F3Hdr h = s_hdr[b];
        const uint4* recs = s_rec + b * FK3_MAXROWS;
        const uint2* recs_hi = s_rec_hi + b * FK3_MAXROWS;
        const bool skip = ((h.misc >> 16) & FK3_SKIP) != 0u;   // handed to the exact kernel by the prep kernel (no rows)
        const i32 need = h.nc_need >> 16;
        u32 wacc[16], cacc[8];
#pragma unroll
        for (int i = 0; i < 8; i++) cacc[i] = 0;
        // one step = two rows: wait for the pair's slot, pull this thread's 16 bytes of each row and the two records ...
        auto step_begin = [&](i32 i, uint4& l0, uint4& l1, uint4& q0, uint4& q1) {
            const u32 s = it & (FK3_PAIRS - 1);
            l0 = recs[i]; l1 = recs[i + 1];
#if FK3_ROWS_BULK
            if (seg_bytes) mbar_wait(full0 + s * 8, (it / FK3_PAIRS) & 1u);
#else
            cp_async_wait<FK3_PAIRS - 1>();
#endif
            const u32 sa0 = my_sa + (2u * s) * 512u;
            q0 = lds128(sa0); q1 = lds128(sa0 + 512u);
        };
        // ... and refill it with the two rows FK3_STAGES further on in the sequence (their table offsets came with this
        // pair's records); the lanes' reads of the slot are complete (their data has been used)
        auto step_end = [&](const uint4& l0, const uint4& l1) {
            const u32 s = it & (FK3_PAIRS - 1);
#if FK3_ROWS_BULK
            __syncwarp();
            if (lane == 0 && seg_bytes) {
                fence_proxy_async();
                mbar_expect_tx(full0 + s * 8, 2u * seg_bytes);
                bulk_g2s(wring_sa + (2u * s) * 512u, wtab + (u64)l0.z * 16, seg_bytes, full0 + s * 8);
                bulk_g2s(wring_sa + (2u * s + 1u) * 512u, wtab + (u64)l1.z * 16, seg_bytes, full0 + s * 8);
            }
#else
            const u32 sa0 = my_sa + (2u * s) * 512u;
            cp_async16_sa(sa0, ltab + (u64)l0.z * 16);
            cp_async16_sa(sa0 + 512u, ltab + (u64)l1.z * 16);
            cp_async_commit();
#endif
            it++;
        };
        {
            u32 xacc[8], yacc[8];
#pragma unroll
            for (int q = 0; q < 8; q++) { xacc[q] = 0; yacc[q] = 0; }
            // ---- two rows per step through the byte LUTs (dense4.cuh: pair_word2 / term_word2)
#pragma unroll 1
            for (i32 i = 0; i < h.n_two; i += 2) {
                uint4 l0, l1, q0, q1;
                step_begin(i, l0, l1, q0, q1);
                if (ALLTERMS) {
                    const uint2 h0 = recs_hi[i], h1 = recs_hi[i + 1];
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
                step_end(l0, l1);
            }
#pragma unroll
            for (int q = 0; q < 8; q++) unpack_xy(xacc[q], yacc[q], wacc[2 * q], wacc[2 * q + 1]);
        }
        // ---- the (at most two) remaining rows of the read, as one step: an odd pair / term row, a Term row of a mixed
        //      read, or the dummy row that pads the count
        if (h.n_two < h.n_rows) {
            const i32 i = h.n_two;
            uint4 l0, l1, q0, q1;
            step_begin(i, l0, l1, q0, q1);
#pragma unroll 1
            for (int sidx = 0; sidx < 2; sidx++) {
                const i32 c = i + sidx;
                if (c >= h.n_real) break;
                const uint4 q = sidx ? q1 : q0;
                const uint4 l = sidx ? l1 : l0;
                const u32 w[4] = {q.x, q.y, q.z, q.w};
                if (ALLTERMS) {
                    const uint2 hh = recs_hi[c];
                    const uint4 t = make_uint4(l.x, l.y, hh.x, hh.y);
#pragma unroll
                    for (int x = 0; x < 4; x++) {
                        const u32 lo = lut16(t, w[x], one), hi = lut16(t, w[x] >> 16, one);
                        wacc[4 * x + 0] += prmt(lo, 0u, 0x4140u); wacc[4 * x + 1] += prmt(lo, 0u, 0x4342u);
                        wacc[4 * x + 2] += prmt(hi, 0u, 0x4140u); wacc[4 * x + 3] += prmt(hi, 0u, 0x4342u);
                        cacc[2 * x] += cnt16(w[x]); cacc[2 * x + 1] += cnt16(w[x] >> 16);
                    }
                } else if (c < h.n_dense) {
                    const uint2 lut = make_uint2(l.x, l.y);
                    pair_word(q.x, lut, wacc[0], wacc[1], wacc[2], wacc[3], cacc[0], cacc[1]);
                    pair_word(q.y, lut, wacc[4], wacc[5], wacc[6], wacc[7], cacc[2], cacc[3]);
                    pair_word(q.z, lut, wacc[8], wacc[9], wacc[10], wacc[11], cacc[4], cacc[5]);
                    pair_word(q.w, lut, wacc[12], wacc[13], wacc[14], wacc[15], cacc[6], cacc[7]);
                } else {                                  // Term clause of a mixed read (tf 1..14): nibble at a time
                    const float sv = __uint_as_float(l.x);
#pragma unroll
                    for (int x = 0; x < 4; x++) {
#pragma unroll
                        for (int bb = 0; bb < 8; bb++) {
                            const u32 code = (w[x] >> (4 * bb)) & 15u;
                            const u32 qv = (u32)__float2int_rn(sv * __ldg(tf16 + code));
                            wacc[4 * x + (bb >> 1)] += qv << (16 * (bb & 1));
                            cacc[2 * x + (bb >> 2)] += (code ? 1u : 0u) << (8 * (bb & 3));
                        }
                    }
                }
            }
            step_end(l0, l1);
        }
        // ---- exception cells of this lane's docs (stored as 0 in the table) join the lanes with the same quantisation
        auto fold_cell = [&](u32 li, u32 qe) {
            const u32 qv = qe << (16 * (li & 1u)), cnt1 = 1u << (8 * (li & 3u));
#pragma unroll
            for (int k = 0; k < 16; k++) wacc[k] += (u32)k == (li >> 1) ? qv : 0u;
#pragma unroll
            for (int k = 0; k < 8; k++) cacc[k] += (u32)k == (li >> 2) ? cnt1 : 0u;
        };
        const u32 n_cells = h.misc & 0xFFu;
        if (n_cells && !skip) {
            const F3Cell* cl = s_cells + b * FK3_CELLS;
            u32 cdoc = 0xFFFFFFFFu, cq = 0;
            if ((u32)lane < n_cells) { const uint2 v = *reinterpret_cast<const uint2*>(cl + lane); cdoc = v.x; cq = v.y; }
            // cells of this slab only: the rest of the warp's time is not spent on the others
            u32 here = __ballot_sync(0xFFFFFFFFu, cdoc - (u32)warp * 1024u < 1024u);
#pragma unroll 1
            while (here) {
                const int src = __ffs((int)here) - 1;
                here &= here - 1u;
                const u32 d = __shfl_sync(0xFFFFFFFFu, cdoc, src), qe = __shfl_sync(0xFFFFFFFFu, cq, src);
                const u32 li = d - d0;
                if (active && li < (u32)FK_DOCS) fold_cell(li, qe);
            }
        }
        // this warp no longer needs the read's shared-memory records: the last warp to say so fetches the records of the
        // read after next into the buffer
        __syncwarp();
        if (lane == 0 && smem_arrive(rcnt0 + b * 4) == (u32)W - 1u) {
            sts32(rcnt0 + b * 4, 0u);
            if (h.nx_read >= 0) {
                fence_proxy_async();
                load_records(h.nx_read, h.nx_sizes & 0xFFFFu, h.nx_sizes >> 16, h.nx_rec, b);
            }
        }
        if (skip) continue;
        const u32 n_long = (h.misc >> 8) & 0xFFu;
        if (n_long) {
            u32 exc_sum = 0;
            const F3Long* lg = longs + (i64)r_cur * FK3_LONG;
#pragma unroll 1
            for (u32 x = 0; x < n_long; x++) {
                const F3Long L = lg[x];
                u32 lo = 0, hi = L.n;
                while (lo < hi) {
                    const u32 mid = (lo + hi) >> 1;
                    if (__ldg(T.exc_doc + L.a + mid) < d0) lo = mid + 1; else hi = mid;
                }
#pragma unroll 1
                for (; lo < L.n; lo++) {
                    const u32 d = __ldg(T.exc_doc + L.a + lo);
                    if (d >= d0 + FK_DOCS) break;
                    const u32 qe = (u32)min(__float2int_rn(L.sv * tf_of(__ldg(T.exc_freq + L.a + lo))), 1 << 20);
                    if (active) { fold_cell(d - d0, qe); exc_sum += qe; }
                }
            }
            // a lane whose u16 sums may have wrapped voids the read's bound: exact_score_kernel takes it
            if (__any_sync(0xFFFFFFFFu, exc_sum > h.budget)) {
                if (lane == 0) atomicExch(&long_overflow[r_cur], 1);
                continue;
            }
        }
        // ---- estimates of the lane's 32 docs: E(d) = Q*m*norm_d for m >= need.  Columns are sorted by norm class, so a
        //      lane's 32 docs share one weight except at the class boundaries: the largest integer Q*m times the smallest
        //      weight of the group is a valid threshold (some doc reaches it), times the largest an upper bound of them all
        float e_tau = 0.0f, e_any = 0.0f;
        bool any_m = active;
        if (active && need <= 128) {
            const u32 need4 = (u32)need * 0x01010101u;
            u32 f = 0;
#pragma unroll
            for (int q = 0; q < 8; q++) f |= ((cacc[q] | 0x80808080u) - need4) | cacc[q];
            any_m = (f & 0x80808080u) != 0u;
        }
        if (any_m) {
            u32 tmax = 0;
#pragma unroll
            for (int i = 0; i < FK_DOCS; i++) {
                const u32 Q = prmt(wacc[i >> 1], 0u, (i & 1) ? 0x4432u : 0x4410u);
                const u32 m = prmt(cacc[i >> 2], 0u, 0x4440u | (u32)(i & 3));
                tmax = max(tmax, (i32)m >= need ? Q * m : 0u);
            }
            const float2 gw = group_w[threadIdx.x];
            e_tau = (float)tmax * gw.x;
            e_any = (float)tmax * gw.y;
        }
        // the slab's threshold: its best estimate (result set) or its 10th best per-lane estimate (full list)
        const u32 eb = __float_as_uint(e_tau);
        const u32 sel = full_topn ? warp_tenth_largest_u32(eb) : warp_max_u32(eb);
        if (warp_max_u32(__float_as_uint(e_any)) == 0u) continue;        // no doc of this slab reaches need
        if (lane == 0 && sel) atomicMax(&tau[r_cur], sel);
        const float theta = __uint_as_float(sel) * h.theta_mul;
        // ---- candidates: a lane holding one parks its lanes, then the 32 lanes evaluate its 32 docs, one each
        const bool mine = active && e_any >= theta && e_any > 0.0f;
        u32 holders = __ballot_sync(0xFFFFFFFFu, mine);
#pragma unroll 1
        while (holders) {
            const int src = __ffs((int)holders) - 1;
            holders &= holders - 1u;
            __syncwarp();
            if (lane == src) {
#pragma unroll
                for (int k = 0; k < 4; k++) park[k] = make_uint4(wacc[4 * k], wacc[4 * k + 1], wacc[4 * k + 2], wacc[4 * k + 3]);
#pragma unroll
                for (int k = 0; k < 2; k++) park[4 + k] = make_uint4(cacc[4 * k], cacc[4 * k + 1], cacc[4 * k + 2], cacc[4 * k + 3]);
            }
            __syncwarp();
            const u32* pw = reinterpret_cast<const u32*>(park);
            const u32 ww = pw[lane >> 1];                 // u16 lanes (doc 2i, 2i+1) = word i
            const u32 cw = pw[16 + (lane >> 2)];          // u8 lanes (docs 4j..4j+3) = word j
            const u32 Q = (lane & 1) ? ww >> 16 : ww & 0xFFFFu;
            const u32 m = (cw >> (8 * (lane & 3))) & 0xFFu;
            const u32 d = ((u32)warp * 32u + (u32)src) * FK_DOCS + (u32)lane;
            const float e = estimate(Q, m, need, col_w[d]);
            const bool keep = e >= theta && e > 0.0f;
            const u32 kb = __ballot_sync(0xFFFFFFFFu, keep);
            if (kb) {
                i32 at = 0;
                if (lane == 0) at = atomicAdd(&cand_n[r_cur], __popc(kb));
                at = __shfl_sync(0xFFFFFFFFu, at, 0);
                const i32 idx = at + __popc(kb & lanemask_lt());
                if (keep && idx < cand_cap) cand[(i64)r_cur * cand_cap + idx] = make_uint2(d, __float_as_uint(e));
            }
        }
    }
    // bulk copies still in flight (the rows fetched past the end of the sequence) must land before the CTA's shared memory
    // is handed to another CTA: every slot has been refilled after its last use, the warp waits for all of its own
#if FK3_ROWS_BULK
    if (lane == 0 && seg_bytes)
        for (u32 k = it; k < it + FK3_PAIRS; k++) mbar_wait(full0 + (k & (FK3_PAIRS - 1)) * 8, (k / FK3_PAIRS) & 1u);
#else
    cp_async_wait<0>();
#endif
}

// ------------------------------------------------------------------ rescore: warp per read
__global__ void __launch_bounds__(256) filter3_rescore_kernel(
        Tab4Dev T, const F3Hdr* __restrict__ hdr, const F3Cell* __restrict__ cells, const F3Long* __restrict__ longs,
        const i32* __restrict__ read_list, i32 n_list, const i64* __restrict__ tok_off, const float* __restrict__ clause_val,
        const u32* __restrict__ clause_row, const float* __restrict__ tf16, const float* __restrict__ col_w,
        int cand_cap, const i32* __restrict__ cand_n, const uint2* __restrict__ cand, const u32* __restrict__ tau,
        const i32* __restrict__ long_overflow, u32 global_doc_base, int parts_stride, i32* __restrict__ part_n,
        u32* __restrict__ part_doc, float* __restrict__ part_score, i32* __restrict__ route, i32* __restrict__ overflow_list,
        i32* __restrict__ overflow_count, FilterStats* __restrict__ stats) {
    __shared__ u64 s_key[8][FK_CAND_CAP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const i32 j = (i32)((blockIdx.x * (u64)blockDim.x + threadIdx.x) >> 5);
    if (j >= n_list) return;
    const i32 r = read_list ? read_list[j] : j;
    const F3Hdr h = hdr[r];
    if ((h.misc >> 16) & FK3_SKIP) return;                // already on the exact list
    const i64 pbase = (i64)r * parts_stride;
    const i32 nraw = cand_n[r];
    if (nraw > cand_cap || long_overflow[r]) {            // the bound could not hold the read: exact_score_kernel scores it
        if (lane == 0) {
            part_n[pbase] = 0;
            route[r] = ROUTE_DENSE;
            overflow_list[atomicAdd(overflow_count, 1)] = r;
            if (stats) atomicAdd(&stats->overflow_reads, 1ull);
        }
        return;
    }
    const i32 nc = h.nc_need & 0xFFFF, need = h.nc_need >> 16;
    const float theta = __uint_as_float(tau[r]) * h.theta_mul;
    const i64 base = tok_off[r];
    const u32 n_cells = h.misc & 0xFFu, n_long = (h.misc >> 8) & 0xFFu;
    const F3Cell* cl = cells + (i64)r * FK3_CELLS;
    const F3Long* lg = longs + (i64)r * FK3_LONG;
    // this lane's exception cell (n_cells <= 32): a candidate's cells are found by ballot instead of a scan per clause
    F3Cell mycell; mycell.doc = 0xFFFFFFFFu; mycell.qe = 0; mycell.clause = 0xFFFFFFFFu; mycell.freq = 0.0f;
    if ((u32)lane < n_cells) mycell = cl[lane];
    i32 n_keep = 0;
#pragma unroll 1
    for (i32 c0 = 0; c0 < nraw; c0 += 32) {
        const i32 ci = c0 + lane;
        uint2 cd = make_uint2(0u, 0u);
        if (ci < nraw) cd = cand[(i64)r * cand_cap + ci];
        const bool keep = ci < nraw && __uint_as_float(cd.y) >= theta;
        u32 kb = __ballot_sync(0xFFFFFFFFu, keep);
#pragma unroll 1
        while (kb) {
            const int src = __ffs((int)kb) - 1;
            kb &= kb - 1u;
            const u32 d = __shfl_sync(0xFFFFFFFFu, cd.x, src);
            const float nrm = col_w[d];
            const u32 cmask = __ballot_sync(0xFFFFFFFFu, mycell.doc == d);     // the read's short-list cells in this doc
            double sum = 0.0;
            i32 m = 0;
            // four clauses per lane and round: their cell loads (one DRAM sector each) are in flight together
#pragma unroll 1
            for (i32 cb = lane; cb < nc; cb += 128) {
                u32 code[4], row[4];
                float val[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const i32 c = cb + 32 * u;
                    code[u] = 0u; row[u] = 0u; val[u] = 0.0f;
                    if (c < nc) { row[u] = clause_row[base + c]; val[u] = clause_val[base + c]; }
                }
#pragma unroll
                for (int u = 0; u < 4; u++) if (cb + 32 * u < nc && row[u] < T.temp_row0) code[u] = cell_code(T, row[u], d);
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const i32 c = cb + 32 * u;
                    if (c >= nc) break;
                    float tf = __ldg(tf16 + code[u]);
                    if (code[u] == 0u) {                  // a true zero, or an exception cell (kept exactly in the lists)
                        u32 mm = cmask;
                        while (mm) {
                            const int x = __ffs((int)mm) - 1;
                            mm &= mm - 1u;
                            const F3Cell e = cl[x];
                            if (e.clause == (u32)c) tf = tf_of(e.freq);
                        }
                        for (u32 x = 0; x < n_long; x++) {
                            const F3Long L = lg[x];
                            if (L.clause == (u32)c) { const float fe = exc_lookup(T, L.a, L.n, d); if (fe != 0.0f) tf = tf_of(fe); }
                        }
                    }
                    if (tf != 0.0f) {
                        sum += (double)__fmul_rn(__fmul_rn(tf, val[u]), nrm);          // TFIDFSimScorer#score
                        m++;
                    }
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
                m += __shfl_xor_sync(0xFFFFFFFFu, m, o);
            }
            const u64 key = m >= need ? cand_key(final_score(SIM_TFIDF, sum, m, nc), T.docmap[d]) : 0ull;
            if (lane == 0 && n_keep < FK_CAND_CAP) s_key[warp][n_keep] = key;
            n_keep++;
        }
    }
    __syncwarp();
    if (n_keep > FK_CAND_CAP) n_keep = FK_CAND_CAP;          // (cand_cap <= FK_CAND_CAP)
    u64 loc[FK_CAND_CAP / 32];
#pragma unroll
    for (int q = 0; q < FK_CAND_CAP / 32; q++) { const int idx = lane + 32 * q; loc[q] = idx < n_keep ? s_key[warp][idx] : 0ull; }
    i32 found = 0;
#pragma unroll 1
    for (int rnd = 0; rnd < TOPN; rnd++) {
        u64 lm = 0;
#pragma unroll
        for (int q = 0; q < FK_CAND_CAP / 32; q++) if (loc[q] > lm) lm = loc[q];
        const u64 best = warp_max_u64(lm);
        if (best == 0) break;
#pragma unroll
        for (int q = 0; q < FK_CAND_CAP / 32; q++) if (loc[q] == best) loc[q] = 0;
        if (lane == 0) {
            part_doc[pbase * TOPN + rnd] = global_doc_base + (0xFFFFFFFFu - (u32)(best & 0xFFFFFFFFu));
            part_score[pbase * TOPN + rnd] = __uint_as_float((u32)(best >> 32));
        }
        found++;
    }
    if (lane == 0) {
        part_n[pbase] = found;
        if (stats) atomicAdd(&stats->candidates, (unsigned long long)n_keep);
    }
}
