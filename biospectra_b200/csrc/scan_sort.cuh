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
// Stable LSD radix sort of (key, u32 value) pairs by key bits [0, key_bits), 8 bits
// per pass.  Three kernels per pass: per-tile digit histogram -> scan of the
// digit-major [256][tiles] matrix -> stable scatter (warp match-any ranking).
#define RS_WARPS 8
#define RS_PER_WARP 512               // items per warp = 16 rows of 32 lanes
#define RS_ROWS (RS_PER_WARP / 32)
#define RS_TILE (RS_WARPS * RS_PER_WARP)

template <typename K>
__global__ void __launch_bounds__(RS_WARPS * 32) rs_hist_kernel(const K* __restrict__ keys, u32 n, int shift,
                                                                 u32* __restrict__ hist, u32 tiles) {
    __shared__ u32 h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    u32 base = blockIdx.x * RS_TILE;
#pragma unroll 4
    for (int i = 0; i < RS_TILE / (RS_WARPS * 32); i++) {
        u32 idx = base + i * (RS_WARPS * 32) + threadIdx.x;
        if (idx < n) atomicAdd(&h[(u32)(keys[idx] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(u64)threadIdx.x * tiles + blockIdx.x] = h[threadIdx.x];
}

template <typename K>
__global__ void __launch_bounds__(RS_WARPS * 32) rs_scatter_kernel(const K* __restrict__ kin, const u32* __restrict__ vin,
                                                                    K* __restrict__ kout, u32* __restrict__ vout, u32 n,
                                                                    int shift, const u32* __restrict__ hist_scanned, u32 tiles) {
    __shared__ u32 wcnt[RS_WARPS][256];
    __shared__ u32 gbase[256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += blockDim.x) (&wcnt[0][0])[i] = 0;
    gbase[threadIdx.x] = hist_scanned[(u64)threadIdx.x * tiles + blockIdx.x];
    __syncthreads();
    const u32 wbase = blockIdx.x * RS_TILE + warp * RS_PER_WARP;
    K key[RS_ROWS];
    u32 rank[RS_ROWS];
    const u32 lt = (1u << lane) - 1u;
#pragma unroll
    for (int j = 0; j < RS_ROWS; j++) {
        u32 idx = wbase + j * 32 + lane;
        bool ok = idx < n;
        key[j] = ok ? kin[idx] : (K)0;
        u32 d = ok ? ((u32)(key[j] >> shift) & 255u) : 256u;      // 256 = out-of-range sentinel group
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
    {   // exclusive prefix over warps for digit = threadIdx.x (blockDim.x == 256)
        u32 run = gbase[threadIdx.x];
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) {
            u32 c = wcnt[w][threadIdx.x];
            wcnt[w][threadIdx.x] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < RS_ROWS; j++) {
        u32 idx = wbase + j * 32 + lane;
        if (idx < n) {
            u32 d = (u32)(key[j] >> shift) & 255u;
            u32 dst = wcnt[warp][d] + rank[j];
            kout[dst] = key[j];
            vout[dst] = vin[idx];
        }
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
    if (n == 0) return;
    u32 tiles = (u32)ceil_div(n, RS_TILE);
    t.hist.ensure((size_t)256 * tiles);
    for (int shift = 0; shift < key_bits; shift += 8) {
        rs_hist_kernel<K><<<tiles, RS_WARPS * 32, 0, st>>>(*keys, n, shift, t.hist.p, tiles);
        exclusive_scan<u32, u32>(t.hist.p, t.hist.p, (u64)256 * tiles, (u32*)nullptr, t.scan_tmp, st);
        rs_scatter_kernel<K><<<tiles, RS_WARPS * 32, 0, st>>>(*keys, *vals, *keys_alt, *vals_alt, n, shift, t.hist.p, tiles);
        KERNEL_CHECK();
        K* tk = *keys; *keys = *keys_alt; *keys_alt = tk;
        u32* tv = *vals; *vals = *vals_alt; *vals_alt = tv;
    }
}
