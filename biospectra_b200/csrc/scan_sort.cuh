// scan_sort.cuh -- device-wide exclusive scan and stable LSD radix sort (hand-written).
//
// Both are HBM-bound streaming primitives: coalesced 128-bit-friendly tile loads,
// shared-memory staging of per-tile state, grids sized from the tile count.
#pragma once
#include "common.cuh"
#include <vector>

// ------------------------------------------------------------------ exclusive scan
#define SCAN_THREADS 256
#define SCAN_ITEMS 8
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

template <typename T>
__device__ __forceinline__ T block_exclusive_scan(T v, T* total, T* smem /* [SCAN_THREADS/32 + 1] */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) smem[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        T w = lane < (blockDim.x >> 5) ? smem[lane] : (T)0;
        T winc = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            T o = __shfl_up_sync(0xFFFFFFFFu, winc, d);
            if (lane >= d) winc += o;
        }
        if (lane < (blockDim.x >> 5)) smem[lane] = winc - w;
        if (lane == 31) smem[32] = winc;
    }
    __syncthreads();
    T res = smem[warp] + inc - v;
    if (total) *total = smem[32];
    __syncthreads();
    return res;
}

// per-tile sums
template <typename TI, typename TO>
__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(const TI* __restrict__ in, TO* __restrict__ sums, u64 n) {
    __shared__ TO sm[33];
    u64 base = (u64)blockIdx.x * SCAN_TILE + (u64)threadIdx.x * SCAN_ITEMS;
    TO s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++)
        if (base + i < n) s += (TO)in[base + i];
    TO tot;
    block_exclusive_scan<TO>(s, &tot, sm);
    if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(const TI* in, TO* out,
                                                                   const TO* __restrict__ offs, u64 n) {
    __shared__ TO sm[33];
    u64 base = (u64)blockIdx.x * SCAN_TILE + (u64)threadIdx.x * SCAN_ITEMS;
    TO v[SCAN_ITEMS];
    TO s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        v[i] = base + i < n ? (TO)in[base + i] : (TO)0;
        s += v[i];
    }
    TO ex = block_exclusive_scan<TO>(s, (TO*)nullptr, sm);
    TO run = ex + (offs ? offs[blockIdx.x] : (TO)0);
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        if (base + i < n) out[base + i] = run;
        run += v[i];
    }
}

// out[i] = sum(in[0..i)) for i in [0,n); if total_out != nullptr, *total_out (device) = sum(in[0..n)).
// in and out may alias.  Temporary memory is taken from `tmp` (grown on demand).
template <typename TI, typename TO>
static void exclusive_scan(const TI* in, TO* out, u64 n, TO* d_total, DevBuf<u8>& tmp, cudaStream_t st) {
    if (n == 0) {
        if (d_total) CUDA_CHECK(cudaMemsetAsync(d_total, 0, sizeof(TO), st));
        return;
    }
    // level sizes
    std::vector<u64> sizes;
    u64 m = n;
    while (true) {
        m = (m + SCAN_TILE - 1) / SCAN_TILE;
        sizes.push_back(m);
        if (m == 1) break;
    }
    u64 total_elems = 0;
    for (u64 s : sizes) total_elems += s;
    tmp.ensure((total_elems + 1) * sizeof(TO));
    std::vector<TO*> lvl(sizes.size());
    TO* p = (TO*)tmp.p;
    for (size_t i = 0; i < sizes.size(); i++) { lvl[i] = p; p += sizes[i]; }
    // reduce up
    scan_reduce_kernel<TI, TO><<<(unsigned)sizes[0], SCAN_THREADS, 0, st>>>(in, lvl[0], n);
    for (size_t i = 1; i < sizes.size(); i++)
        scan_reduce_kernel<TO, TO><<<(unsigned)sizes[i], SCAN_THREADS, 0, st>>>(lvl[i - 1], lvl[i], sizes[i - 1]);
    // top level holds one element = grand total
    if (d_total) CUDA_CHECK(cudaMemcpyAsync(d_total, lvl.back(), sizeof(TO), cudaMemcpyDeviceToDevice, st));
    // scan down
    for (size_t i = sizes.size() - 1; i >= 1; i--) {
        const TO* offs = (i == sizes.size() - 1) ? (const TO*)nullptr : lvl[i];
        // lvl[i-1] (size sizes[i-1]) scanned in place using offsets from the (already scanned) lvl[i]
        scan_apply_kernel<TO, TO><<<(unsigned)sizes[i], SCAN_THREADS, 0, st>>>(lvl[i - 1], lvl[i - 1], offs, sizes[i - 1]);
    }
    const TO* offs0 = sizes.size() == 1 ? (const TO*)nullptr : lvl[0];
    // when there is a single level, lvl[0] holds the total, not an offset
    scan_apply_kernel<TI, TO><<<(unsigned)sizes[0], SCAN_THREADS, 0, st>>>(in, out, sizes[0] == 1 ? (const TO*)nullptr : offs0, n);
    KERNEL_CHECK();
}

// ------------------------------------------------------------------ radix sort
// Stable LSD radix sort of (key, u32 value) pairs by key bits [0, key_bits), <= 8 bits per pass (the passes split the
// key evenly: 20 bits -> 7+7+6).  Three kernels per pass: per-tile digit histogram -> scan of the digit-major
// [256][tiles] matrix -> stable scatter.  The scatter ranks its tile with warp match-any, orders the tile by digit in
// shared memory and writes it out run by run, so that consecutive threads store to consecutive addresses.
#define RS_WARPS 8
#define RS_THREADS (RS_WARPS * 32)
template <typename K> struct RsTile { static const int ROWS = sizeof(K) == 4 ? 16 : 8; };   // 32-item rows per warp
#define RS_TILE_OF(K) (RS_WARPS * 32 * RsTile<K>::ROWS)

template <typename K>
__global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(const K* __restrict__ keys, u32 n, int shift, u32 mask,
                                                              u32* __restrict__ hist, u32 tiles) {
    __shared__ u32 h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const u32 TILE = RS_TILE_OF(K);
    u32 base = blockIdx.x * TILE;
#pragma unroll 4
    for (int i = 0; i < (int)(TILE / RS_THREADS); i++) {
        u32 idx = base + i * RS_THREADS + threadIdx.x;
        if (idx < n) atomicAdd(&h[(u32)(keys[idx] >> shift) & mask], 1u);
    }
    __syncthreads();
    hist[(u64)threadIdx.x * tiles + blockIdx.x] = h[threadIdx.x];
}

template <typename K>
__global__ void __launch_bounds__(RS_THREADS) rs_scatter_kernel(const K* __restrict__ kin, const u32* __restrict__ vin,
                                                                 K* __restrict__ kout, u32* __restrict__ vout, u32 n,
                                                                 int shift, u32 mask, const u32* __restrict__ hist_scanned, u32 tiles) {
    const int ROWS = RsTile<K>::ROWS;
    const u32 TILE = RS_TILE_OF(K);
    __shared__ u32 wcnt[RS_WARPS][256];      // per warp digit counts -> exclusive offsets inside the tile
    __shared__ u32 dstart[257];              // first tile-local slot of every digit
    __shared__ u32 gbase[256];               // first global slot of every digit for this tile
    __shared__ u32 scan_sm[33];
    __shared__ K s_key[RS_WARPS * 32 * RsTile<K>::ROWS];
    __shared__ u32 s_val[RS_WARPS * 32 * RsTile<K>::ROWS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += blockDim.x) (&wcnt[0][0])[i] = 0;
    gbase[threadIdx.x] = hist_scanned[(u64)threadIdx.x * tiles + blockIdx.x];
    __syncthreads();
    const u32 tbase = blockIdx.x * TILE;
    const u32 wbase = tbase + warp * (32 * ROWS);
    K key[ROWS];
    u32 val[ROWS], rank[ROWS];
    const u32 lt = (1u << lane) - 1u;
#pragma unroll
    for (int j = 0; j < ROWS; j++) {
        u32 idx = wbase + j * 32 + lane;
        bool ok = idx < n;
        key[j] = ok ? kin[idx] : (K)0;
        val[j] = ok ? vin[idx] : 0u;
        u32 d = ok ? ((u32)(key[j] >> shift) & mask) : 256u;      // 256 = out-of-range sentinel group
        u32 grp = __match_any_sync(0xFFFFFFFFu, d);
        int leader = __ffs(grp) - 1;
        u32 b = 0;
        if (ok && lane == leader) {
            b = wcnt[warp][d];
            wcnt[warp][d] = b + __popc(grp);
        }
        b = __shfl_sync(0xFFFFFFFFu, b, leader);
        rank[j] = b + __popc(grp & lt);
        __syncwarp();
    }
    __syncthreads();
    {   // digit = threadIdx.x: exclusive prefix over warps, then over digits (tile-local layout: digit, warp, rank)
        u32 tot = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) {
            u32 c = wcnt[w][threadIdx.x];
            wcnt[w][threadIdx.x] = tot;
            tot += c;
        }
        u32 ex = block_exclusive_scan<u32>(tot, (u32*)nullptr, scan_sm);
        dstart[threadIdx.x] = ex;
        if (threadIdx.x == 255) dstart[256] = ex + tot;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < ROWS; j++) {
        u32 idx = wbase + j * 32 + lane;
        if (idx < n) {
            u32 d = (u32)(key[j] >> shift) & mask;
            u32 slot = dstart[d] + wcnt[warp][d] + rank[j];
            s_key[slot] = key[j];
            s_val[slot] = val[j];
        }
    }
    __syncthreads();
    const u32 n_tile = dstart[256];
    for (u32 i = threadIdx.x; i < n_tile; i += blockDim.x) {
        const K k = s_key[i];
        const u32 d = (u32)(k >> shift) & mask;
        const u32 dst = gbase[d] + (i - dstart[d]);
        kout[dst] = k;
        vout[dst] = s_val[i];
    }
}

struct RadixSortTemp {
    DevBuf<u32> hist;
    DevBuf<u8> scan_tmp;
};

// Sorts in place logically: on return *keys/*vals point at the buffers holding the
// sorted data (either the originals or the alternates).
template <typename K>
static void radix_sort_pairs(K** keys, u32** vals, K** keys_alt, u32** vals_alt, u32 n, int key_bits,
                             RadixSortTemp& t, cudaStream_t st) {
    if (n == 0 || key_bits <= 0) return;
    const u32 TILE = RS_TILE_OF(K);
    u32 tiles = (u32)ceil_div(n, TILE);
    t.hist.ensure((size_t)256 * tiles);
    const int passes = (key_bits + 7) / 8;
    int shift = 0;
    for (int p = 0; p < passes; p++) {
        const int bits = (key_bits - shift + (passes - p) - 1) / (passes - p);     // spread the bits evenly over the passes
        const u32 mask = (1u << bits) - 1u;
        rs_hist_kernel<K><<<tiles, RS_THREADS, 0, st>>>(*keys, n, shift, mask, t.hist.p, tiles);
        exclusive_scan<u32, u32>(t.hist.p, t.hist.p, (u64)256 * tiles, (u32*)nullptr, t.scan_tmp, st);
        rs_scatter_kernel<K><<<tiles, RS_THREADS, 0, st>>>(*keys, *vals, *keys_alt, *vals_alt, n, shift, mask, t.hist.p, tiles);
        KERNEL_CHECK();
        K* tk = *keys; *keys = *keys_alt; *keys_alt = tk;
        u32* tv = *vals; *vals = *vals_alt; *vals_alt = tv;
        shift += bits;
    }
}
