// Would a double-precision Montgomery product beat the 8 x 32-bit IMAD.WIDE one on B200?  (round-1 verdict, item 4)
//
// This is a THROUGHPUT MODEL, not a field implementation: k_dfma_product runs, per "product", exactly the instruction mix a
// 5 x 52-bit-limb Montgomery product needs with the standard split-FMA technique --
//   per limb pair:  hi = fma_rz(a, b, 2^104);  lo = fma_rz(a, b, (2^104 + 2^52) - hi);   (2 DFMA + 1 DADD)
//                   col[i + j + 1] += bits(hi);  col[i + j] += bits(lo);                 (two 64-bit integer additions)
//   25 pairs for a x b, 5 quotient digits (one low product each) and 25 pairs for q x p, then 5 limbs of carry propagation
//   and re-encoding as doubles (mask, exponent OR, one DADD each)
// -- with every product's result fed into the next one's operands, so nothing is hoisted.  The values are not reduced
// modulo anything; the instruction counts, pipes and dependencies are the real ones.  k_imad_product runs the library's
// fp_mul chain for comparison, k_mixed gives even warps the one and odd warps the other (do the two pipes overlap?).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -o microbench/_build/dfma_mix_bench microbench/dfma_mix_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../spartan_bn254_b200/csrc/fp.cuh"
using namespace sbn;

#define ITERS 2048

__device__ __forceinline__ void limb_pair(double a, double b, long long& c_hi, long long& c_lo) {
    const double C1 = 20282409603651670423947251286016.0;          // 2^104
    const double C2 = 20282409603651674927546878656512.0;          // 2^104 + 2^52
    const double hi = __fma_rz(a, b, C1);
    const double lo = __fma_rz(a, b, C2 - hi);
    c_hi += __double_as_longlong(hi);
    c_lo += __double_as_longlong(lo);
}
__device__ __forceinline__ double limb_to_double(long long v) {    // low 52 bits of v as an exact double
    const long long m = (v & 0x000fffffffffffffLL) | 0x4330000000000000LL;
    return __longlong_as_double(m) - 4503599627370496.0;
}

__device__ __forceinline__ void dfma_product(double (&a)[5], const double (&b)[5], const double (&p)[5], double inv) {
    long long col[11];
#pragma unroll
    for (int i = 0; i < 11; i++) col[i] = 0;
#pragma unroll
    for (int i = 0; i < 5; i++)
#pragma unroll
        for (int j = 0; j < 5; j++) limb_pair(a[i], b[j], col[i + j + 1], col[i + j]);
#pragma unroll
    for (int i = 0; i < 5; i++) {                                   // word-serial reduction: q_i = low(col[i] * inv), col += q_i * p << 52 i
        long long qh = 0, ql = 0;
        limb_pair(limb_to_double(col[i]), inv, qh, ql);
        const double q = limb_to_double(ql);
#pragma unroll
        for (int j = 0; j < 5; j++) limb_pair(q, p[j], col[i + j + 1], col[i + j]);
        col[i + 1] += col[i] >> 52;
    }
#pragma unroll
    for (int i = 0; i < 5; i++) {                                   // carry propagation and back to doubles
        if (i) col[5 + i] += col[5 + i - 1] >> 52;
        a[i] = limb_to_double(col[5 + i]);
    }
}

__global__ void k_dfma_product(double* out, double seed) {
    double a[5], b[5], p[5];
    for (int i = 0; i < 5; i++) { a[i] = 1000.0 + threadIdx.x + i; b[i] = 77777.0 + seed + i; p[i] = 4503599627370001.0 - i; }
    for (int it = 0; it < ITERS; it++) dfma_product(a, b, p, 4503599627365555.0);
    double r = 0;
    for (int i = 0; i < 5; i++) r += a[i];
    if (r == 1234.5) out[0] = r;
}
__global__ void k_imad_product(uint32_t* out, uint32_t seed) {
    Fq a, b;
    for (int i = 0; i < 8; i++) { a.l[i] = threadIdx.x * 2654435761u + i; b.l[i] = seed + 17 * i; }
    a.l[7] &= 0x0fffffff; b.l[7] &= 0x0fffffff;
    for (int it = 0; it < ITERS; it++) a = fp_mul(a, b);
    uint32_t r = 0;
    for (int i = 0; i < 8; i++) r ^= a.l[i];
    if (r == 0x12345678u) out[0] = r;
}
__global__ void k_mixed(double* outd, uint32_t* outu, double seed) {
    if ((threadIdx.x >> 5) & 1) {
        Fq a, b;
        for (int i = 0; i < 8; i++) { a.l[i] = threadIdx.x * 2654435761u + i; b.l[i] = (uint32_t)seed + 17 * i; }
        a.l[7] &= 0x0fffffff; b.l[7] &= 0x0fffffff;
        for (int it = 0; it < ITERS; it++) a = fp_mul(a, b);
        uint32_t r = 0;
        for (int i = 0; i < 8; i++) r ^= a.l[i];
        if (r == 0x12345678u) outu[0] = r;
    } else {
        double a[5], b[5], p[5];
        for (int i = 0; i < 5; i++) { a[i] = 1000.0 + threadIdx.x + i; b[i] = 77777.0 + seed + i; p[i] = 4503599627370001.0 - i; }
        for (int it = 0; it < ITERS; it++) dfma_product(a, b, p, 4503599627365555.0);
        double r = 0;
        for (int i = 0; i < 5; i++) r += a[i];
        if (r == 1234.5) outd[0] = r;
    }
}

template <class F>
static double time_kernel(F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 3; r++) launch();
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms / 3 * 1e-3;
}

int main() {
    double* dd; uint32_t* du;
    cudaMalloc(&dd, 8); cudaMalloc(&du, 4);
    for (int threads : {128, 256}) {
        const int blocks = 148 * 16;
        const double n = (double)blocks * threads * ITERS;
        const double td = time_kernel([&] { k_dfma_product<<<blocks, threads>>>(dd, 3.0); });
        const double ti = time_kernel([&] { k_imad_product<<<blocks, threads>>>(du, 3u); });
        const double tm = time_kernel([&] { k_mixed<<<blocks, threads>>>(dd, du, 3.0); });
        printf("threads/block %d: DFMA-mix product %.3e /s   IMAD.WIDE product (fp_mul) %.3e /s   half-and-half warps %.3e /s\n", threads,
               n / td, n / ti, n / tm);
    }
    return 0;
}
