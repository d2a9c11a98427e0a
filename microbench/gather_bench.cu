// Random 64 B / 32 B gathers over a table of tens of GB (sm_100a): what does one table point cost in DRAM traffic and
// time, and do the PTX L2 prefetch-size qualifiers or cudaLimitMaxL2FetchGranularity change it?
// Round 1 of the tabulated-sum commit (mult_kernels.cuh) is exactly this access pattern: ncu shows 128 B fetched per 64 B
// point and per 32 B x-coordinate (profiles/r1i_ba_round_kernels_summary.txt).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench/_build/gather_bench microbench/gather_bench.cu
// Run:   gather_bench [table GiB] ; under ncu add --metrics dram__bytes_read.sum for the bytes per variant.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t z) {
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

template <int V>
__device__ __forceinline__ uint4 ld16(const uint4* p) {
    uint4 r;
    if (V == 0) asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    if (V == 1) asm volatile("ld.global.nc.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    if (V == 2) asm volatile("ld.global.nc.L2::128B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    if (V == 3) asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    if (V == 4) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    if (V == 5) asm volatile("ld.global.cv.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

// every thread gathers K points of BYTES bytes (32 or 64) at pseudo-random 64 B-aligned offsets
// LOCAL: the 32 lanes of a warp gather from ONE 2 MiB region per step (random region per warp and step, random 64 B slot per
// lane) -- the access pattern of a warp whose lanes are 32 rows reading the same (window, generator) column of the table.
template <int V, int BYTES, bool LOCAL = false>
__global__ void k_gather(const uint4* __restrict__ table, uint64_t npoints, int K, uint32_t seed, uint32_t* out) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t acc = 0;
    for (int k = 0; k < K; k += 4) {
        uint4 v[4][BYTES / 16];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            uint64_t idx = mix(tid * 64 + k + u + ((uint64_t)seed << 40)) % npoints;
            if (LOCAL) {
                const uint64_t region = mix((tid >> 5) * 64 + k + u + ((uint64_t)seed << 40) + 7) % (npoints >> 15);
                idx = (region << 15) + (idx & 32767);
            }
            const uint4* p = table + idx * 4;
#pragma unroll
            for (int q = 0; q < BYTES / 16; q++) v[u][q] = ld16<V>(p + q);
        }
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int q = 0; q < BYTES / 16; q++) acc ^= v[u][q].x ^ v[u][q].y ^ v[u][q].z ^ v[u][q].w;
    }
    if (acc == 0x12345678u) out[0] = acc;
}

template <int V, int BYTES, bool LOCAL = false>
static void run(const char* name, const uint4* table, uint64_t npoints, uint32_t* out) {
    const int K = 16, threads = 256, blocks = 148 * 8 * 4;
    cudaEvent_t e0, e1;
    CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
    k_gather<V, BYTES, LOCAL><<<blocks, threads>>>(table, npoints, K, 1, out);
    CHECK(cudaDeviceSynchronize());
    CHECK(cudaEventRecord(e0));
    const int reps = 5;
    for (int r = 0; r < reps; r++) k_gather<V, BYTES, LOCAL><<<blocks, threads>>>(table, npoints, K, 2 + r, out);
    CHECK(cudaEventRecord(e1));
    CHECK(cudaDeviceSynchronize());
    float ms = 0;
    CHECK(cudaEventElapsedTime(&ms, e0, e1));
    const double n = (double)reps * blocks * threads * K;
    printf("  %-44s %2d B/point: %7.1f Mpoints/ms-GPU = %6.2f Gpoints/s, useful %7.1f GB/s\n", name, BYTES, n / ms / 1e6 * 1e3 / 1e3,
           n / (ms * 1e-3) / 1e9, n * BYTES / (ms * 1e-3) / 1e9);
}

int main(int argc, char** argv) {
    const double gib = argc > 1 ? atof(argv[1]) : 32.0;
    const uint64_t bytes = (uint64_t)(gib * (1ull << 30));
    const uint64_t npoints = bytes / 64;
    uint4* table = nullptr;
    uint32_t* out = nullptr;
    CHECK(cudaMalloc(&table, bytes));
    CHECK(cudaMalloc(&out, 4));
    CHECK(cudaMemset(table, 1, bytes));
    printf("table %.0f GiB; lanes of a warp in one 2 MiB region per step:\n", gib);
    run<0, 64, true>("ld.global.nc, warp-local region", table, npoints, out);
    run<1, 64, true>("ld.global.nc.L2::64B, warp-local region", table, npoints, out);
    run<0, 32, true>("ld.global.nc, warp-local region", table, npoints, out);
    run<1, 32, true>("ld.global.nc.L2::64B, warp-local region", table, npoints, out);
    for (int gran : {0, 32}) {
        if (gran) {
            cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
            if (e != cudaSuccess) { printf("cudaLimitMaxL2FetchGranularity=%d: %s\n", gran, cudaGetErrorString(e)); cudaGetLastError(); continue; }
        }
        size_t cur = 0;
        cudaDeviceGetLimit(&cur, cudaLimitMaxL2FetchGranularity);
        printf("cudaLimitMaxL2FetchGranularity = %zu (%s), table %.0f GiB\n", cur, gran ? "set" : "default", gib);
        run<0, 64>("ld.global.nc", table, npoints, out);
        run<1, 64>("ld.global.nc.L2::64B", table, npoints, out);
        run<2, 64>("ld.global.nc.L2::128B", table, npoints, out);
        run<3, 64>("ld.global.cg", table, npoints, out);
        run<4, 64>("ld.global.nc.L1::no_allocate", table, npoints, out);
        run<5, 64>("ld.global.cv", table, npoints, out);
        run<0, 32>("ld.global.nc", table, npoints, out);
        run<1, 32>("ld.global.nc.L2::64B", table, npoints, out);
        run<3, 32>("ld.global.cg", table, npoints, out);
        run<5, 32>("ld.global.cv", table, npoints, out);
    }
    return 0;
}
