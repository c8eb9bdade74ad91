// common.cuh -- shared device/host helpers for libbiospectra_gpu (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <chrono>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <stdexcept>

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int32_t i32;
typedef int64_t i64;

struct BspError : public std::runtime_error {
    int code;
    BspError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define CUDA_CHECK(expr)                                                                       \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            char _b[512];                                                                      \
            snprintf(_b, sizeof _b, "CUDA error %s at %s:%d: %s", cudaGetErrorName(_e),        \
                     __FILE__, __LINE__, cudaGetErrorString(_e));                              \
            throw BspError(_e == cudaErrorMemoryAllocation ? -6 : -3, _b);                     \
        }                                                                                      \
    } while (0)

#define KERNEL_CHECK() CUDA_CHECK(cudaGetLastError())
// Debug build (make EXTRA=-DBSP_DEBUG_BOUNDS): device-side bounds assertions around the shared-memory rings, tables and
// lists of the scoring kernels -- compute-sanitizer is not available on every pool, this is what the GPU tests can be run
// under instead (a failing assertion traps the kernel: the call returns BSP_ERR_CUDA and the test fails).  Nothing in a
// release build.
#ifdef BSP_DEBUG_BOUNDS
#include <cassert>
#define BSP_ASSERT(cond) assert(cond)
#else
#define BSP_ASSERT(cond) ((void)0)
#endif

static inline i64 ceil_div(i64 a, i64 b) { return (a + b - 1) / b; }
static inline i64 round_up(i64 a, i64 b) { return ceil_div(a, b) * b; }

// ---------------------------------------------------------------- device buffer
// RAII device allocation (no torch: plain cudaMalloc).
// host time spent in cudaMalloc / cudaFree (reported by the build when BSP_BUILD_TIMING is set)
struct AllocStats { double alloc_ms = 0, free_ms = 0; long allocs = 0, frees = 0; };
static thread_local AllocStats g_alloc_stats;     // per thread: builds of different handles may run concurrently

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    bool owned = true;                  // false: a view into another allocation (adopt)
    DevBuf() {}
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n), owned(o.owned) { o.p = nullptr; o.n = 0; o.owned = true; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; owned = o.owned; o.p = nullptr; o.n = 0; o.owned = true; }
        return *this;
    }
    ~DevBuf() { release(); }
    // view of `count` elements at `ptr` inside a block somebody else owns
    void adopt(T* ptr, size_t count) { release(); p = ptr; n = count; owned = false; }
    void release() {
        if (p && !owned) { p = nullptr; n = 0; owned = true; return; }
        if (p) {
            const auto t0 = std::chrono::steady_clock::now();
            cudaFree(p);
            g_alloc_stats.free_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            g_alloc_stats.frees++;
            p = nullptr; n = 0;
        }
    }
    void alloc(size_t count) {
        release();
        if (count == 0) count = 1;
        const auto t0 = std::chrono::steady_clock::now();
        CUDA_CHECK(cudaMalloc((void**)&p, count * sizeof(T)));
        g_alloc_stats.alloc_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        g_alloc_stats.allocs++;
        n = count;
    }
    void ensure(size_t count) { if (count > n) alloc(count + count / 8); }
    size_t bytes() const { return n * sizeof(T); }
};

// --------------------------------------------------------- packed sequence access
// Packed layout: u32 words, 16 bases per word, base j of a word in bits (30-2j)
// (first base most significant, like SequenceHelper.compress: A0 C1 G2 T3).
// Invalid-base mask: u32 words, 32 bases per word, bit j = base 32i+j is not ACGT.
// Both arrays are padded with >= 4 zero words past the last base of every entry.

// k-mer (k<=32) starting at base p (relative to the word array), MSB-first in 2k bits.
__device__ __forceinline__ u64 load_kmer(const u32* __restrict__ w, u64 p, int k) {
    u64 i = p >> 4;
    int sh = (int)(p & 15) * 2;
    u64 hi = ((u64)w[i] << 32) | (u64)w[i + 1];
    u64 v = hi << sh;
    if (sh) v |= (u64)(w[i + 2] >> (32 - sh));
    return v >> (64 - 2 * k);
}

// true iff any of the k (<=32) bases starting at p is invalid
__device__ __forceinline__ bool window_invalid(const u32* __restrict__ m, u64 p, int k) {
    u64 i = p >> 5;
    int sh = (int)(p & 31);
    u32 bits = __funnelshift_r(m[i], m[i + 1], sh);
    u32 km = k >= 32 ? 0xFFFFFFFFu : ((1u << k) - 1u);
    return (bits & km) != 0;
}

// reverse complement of a packed k-mer
__host__ __device__ __forceinline__ u64 rc_packed(u64 v, int k) {
    // complement = 3 - code = bitwise NOT of each 2-bit field
    v = ~v;
    // reverse 2-bit fields of the 64-bit word
    v = ((v >> 2) & 0x3333333333333333ull) | ((v & 0x3333333333333333ull) << 2);
    v = ((v >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((v & 0x0F0F0F0F0F0F0F0Full) << 4);
    v = ((v >> 8) & 0x00FF00FF00FF00FFull) | ((v & 0x00FF00FF00FF00FFull) << 8);
    v = ((v >> 16) & 0x0000FFFF0000FFFFull) | ((v & 0x0000FFFF0000FFFFull) << 16);
    v = (v >> 32) | (v << 32);
    return v >> (64 - 2 * k);
}

// SequenceCompressFilter.java:56-68: min(kmer, revcomp) under A<C<G<T == numeric min
__host__ __device__ __forceinline__ u64 canon_kmer(u64 v, int k, int min_strand) {
    if (!min_strand) return v;
    u64 r = rc_packed(v, k);
    return v > r ? r : v;
}

__host__ __device__ __forceinline__ u64 mix64(u64 x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return x;
}

#define HASH_EMPTY 0xFFFFFFFFFFFFFFFFull
#define NO_TERM 0xFFFFFFFFu
