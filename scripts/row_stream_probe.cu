// Probe (not part of the library): what HBM bandwidth does the filter kernel's access pattern reach with no
// arithmetic at all?  CTA per "read": 192 threads each stream their 16-byte slice of 46 random 3,008-byte table
// rows through a private cp.async ring (8 in flight), XOR the words, one store per thread at the end.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o row_stream_probe row_stream_probe.cu && ./row_stream_probe
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
typedef unsigned int u32;

template <int STAGES>
__global__ void __launch_bounds__(192) probe(const uint4* __restrict__ tab, const u32* __restrict__ rows, int n_rows, u32 row16,
                                             int persistent, int n_reads, u32* __restrict__ out) {
    extern __shared__ uint4 ring[];
    const u32 sa = (u32)__cvta_generic_to_shared(ring + threadIdx.x), sb = blockDim.x * 16u;
    u32 acc = 0;
    for (int q = blockIdx.x; q < n_reads; q += gridDim.x) {
        const u32* rr = rows + (size_t)q * n_rows;
        if (threadIdx.x >= row16) break;
#pragma unroll
        for (int st = 0; st < STAGES; st++) {
            if (st < n_rows) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa + st * sb), "l"(tab + (size_t)rr[st] * row16 + threadIdx.x) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        for (int i = 0; i < n_rows; i++) {
            asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 1) : "memory");
            uint4 v;
            const u32 a = sa + (i % STAGES) * sb;
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
            acc ^= v.x ^ v.y ^ v.z ^ v.w;
            if (i + STAGES < n_rows) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(a), "l"(tab + (size_t)rr[i + STAGES] * row16 + threadIdx.x) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        if (!persistent) break;
    }
    if (acc == 0x12345678u) out[blockIdx.x] = acc;
}

int main() {
    const u32 row_bytes = 3008, row16 = row_bytes / 16;
    const size_t n_tab_rows = 5242880;                 // 4^10 + 4^11 rows = 15.8 GB
    const int n_reads = 1000000, n_rows = 46;
    uint4* tab; u32* rows; u32* out;
    CK(cudaMalloc(&tab, n_tab_rows * row_bytes));
    CK(cudaMemset(tab, 1, n_tab_rows * row_bytes));
    std::vector<u32> h((size_t)n_reads * n_rows);
    uint64_t s = 88172645463325252ull;
    for (auto& x : h) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; x = (u32)(s % n_tab_rows); }
    CK(cudaMalloc(&rows, h.size() * 4));
    CK(cudaMemcpy(rows, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&out, n_reads * 4));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double bytes = (double)n_reads * n_rows * row_bytes;
    auto run = [&](const char* name, auto kern, int stages, int pad_kb, int persistent) {
        const size_t smem = (size_t)stages * 192 * 16 + (size_t)pad_kb * 1024;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 192, smem));
        const int grid = persistent ? 148 * per_sm : n_reads;
        float best = 1e9f;
        for (int it = 0; it < 4; it++) {
            cudaEventRecord(e0);
            kern<<<grid, 192, smem>>>(tab, rows, n_rows, row16, persistent, n_reads, out);
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (it && ms < best) best = ms;
        }
        printf("%-34s stages %2d  CTAs/SM %2d  %7.2f ms  %7.1f GB/s\n", name, stages, per_sm, best, bytes / best * 1e-6);
    };
    run("cta per read", probe<8>, 8, 0, 0);
    run("cta per read, 5 CTAs/SM", probe<8>, 8, 20, 0);
    run("cta per read, 4 CTAs/SM", probe<8>, 8, 31, 0);
    run("cta per read, 16 stages", probe<16>, 16, 0, 0);
    run("persistent (ring drains per read)", probe<8>, 8, 0, 1);
    run("persistent, 5 CTAs/SM", probe<8>, 8, 20, 1);
    run("cta per read, 4 stages", probe<4>, 4, 0, 0);
    return 0;
}
