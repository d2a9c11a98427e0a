// Integer-pipe microbenchmarks for sm_100a: what the fma pipe sustains for the instruction shapes a
// Montgomery multiplication is made of.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
#include "../spartan_bn254_b200/csrc/fp.cuh"
using namespace sbn;

#define ITERS 4096

// kind 0: independent IMAD.WIDE.U32 (64-bit accumulate, no carry)
__global__ void k_wide(uint32_t* out, uint32_t seed) {
    uint64_t w[8]; uint32_t a[8], b = seed | 1;
    for (int i = 0; i < 8; i++) { a[i] = threadIdx.x * 2654435761u + i; w[i] = a[i]; }
    for (int it = 0; it < ITERS; it++)
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(b));
    uint32_t r = 0; for (int i = 0; i < 8; i++) r ^= (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32);
    if (r == 0x12345678u) out[0] = r;
}
// kind 0b: IMAD.WIDE whose multiplicand evolves (cannot be hoisted): w = lo(w) * b + w
__global__ void k_wide_dep(uint32_t* out, uint32_t seed) {
    uint64_t w[8]; uint32_t b = seed | 1;
    for (int i = 0; i < 8; i++) w[i] = (threadIdx.x * 2654435761u + i) | 1;
    for (int it = 0; it < ITERS; it++)
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"((uint32_t)w[i]), "r"(b));
    uint32_t r = 0; for (int i = 0; i < 8; i++) r ^= (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32);
    if (r == 0x12345678u) out[0] = r;
}
// kind 0c: mul.wide (no addend) with evolving multiplicand, results xor-folded on the ALU pipe
__global__ void k_mulwide_dep(uint32_t* out, uint32_t seed) {
    uint32_t a[8]; uint32_t b = seed | 1;
    for (int i = 0; i < 8; i++) a[i] = (threadIdx.x * 2654435761u + i) | 1;
    for (int it = 0; it < ITERS; it++)
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int i = 0; i < 8; i++) { uint64_t w; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(a[i]), "r"(b)); a[i] = (uint32_t)w ^ (uint32_t)(w >> 32); }
    uint32_t r = 0; for (int i = 0; i < 8; i++) r ^= a[i];
    if (r == 0x12345678u) out[0] = r;
}
// FP64 pipe: independent DFMA chains (round-toward-zero, as in double-precision big-integer multiplication)
__global__ void k_dfma(uint32_t* out, uint32_t seed) {
    double a[8], b = 1.0 + seed * 1e-9, c = 0.5;
    for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 1.25 + i;
    for (int it = 0; it < ITERS; it++)
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = __fma_rz(a[i], b, c);
    double r = 0; for (int i = 0; i < 8; i++) r += a[i];
    if (r == 1234.5) out[0] = 1;
}
// DFMA and IMAD.WIDE carry chains interleaved in one instruction stream: do the two pipes overlap?
__global__ void k_dfma_imad(uint32_t* out, uint32_t seed) {
    double a[8], b = 1.0 + seed * 1e-9, c = 0.5;
    uint32_t acc[8], m[8], bi = seed | 1;
    for (int i = 0; i < 8; i++) { a[i] = threadIdx.x * 1.25 + i; m[i] = threadIdx.x * 2654435761u + i; acc[i] = m[i]; }
    for (int it = 0; it < ITERS; it++)
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = __fma_rz(a[i], b, c);
            acc[0] = mad_lo_cc(m[0], bi, acc[0]); acc[1] = madc_hi_cc(m[0], bi, acc[1]);
#pragma unroll
            for (int j = 2; j < 8; j += 2) { acc[j] = madc_lo_cc(m[j], bi, acc[j]); acc[j + 1] = madc_hi_cc(m[j], bi, acc[j + 1]); }
        }
    double r = 0; for (int i = 0; i < 8; i++) r += a[i] + acc[i];
    if (r == 1234.5) out[0] = 1;
}
// kind 1: carry chains of 4 wide products (8 limbs): acc += a_even * b, NCH independent accumulators
template <int NCH>
__global__ void k_chain(uint32_t* out, uint32_t seed) {
    uint32_t acc[NCH][8], a[8], b = seed | 1;
    for (int i = 0; i < 8; i++) { a[i] = threadIdx.x * 2654435761u + i; for (int c = 0; c < NCH; c++) acc[c][i] = a[i] + c; }
    for (int it = 0; it < ITERS; it++)
#pragma unroll
        for (int u = 0; u < 8 / NCH; u++)
#pragma unroll
            for (int c = 0; c < NCH; c++) {
                uint32_t* x = acc[c];
                x[0] = mad_lo_cc(a[0], b, x[0]); x[1] = madc_hi_cc(a[0], b, x[1]);
#pragma unroll
                for (int j = 2; j < 8; j += 2) { x[j] = madc_lo_cc(a[j], b, x[j]); x[j + 1] = madc_hi_cc(a[j], b, x[j + 1]); }
            }
    uint32_t r = 0; for (int c = 0; c < NCH; c++) for (int i = 0; i < 8; i++) r ^= acc[c][i];
    if (r == 0x12345678u) out[0] = r;
}
// kind 2: Montgomery multiplications, NCH independent dependency chains per thread
template <int NCH>
__global__ void k_mul(uint32_t* out, uint32_t seed) {
    Fq a[NCH], b[NCH];
    for (int c = 0; c < NCH; c++) for (int i = 0; i < 8; i++) { a[c].l[i] = (threadIdx.x + 1) * 2654435761u + i + c; b[c].l[i] = seed + 77 * i + c; }
    for (int c = 0; c < NCH; c++) { a[c].l[7] &= 0x0fffffffu; b[c].l[7] &= 0x0fffffffu; }
    for (int it = 0; it < ITERS / 16; it++)
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int c = 0; c < NCH; c++) a[c] = fp_mul(a[c], b[c]);
    uint32_t r = 0; for (int c = 0; c < NCH; c++) for (int i = 0; i < 8; i++) r ^= a[c].l[i];
    if (r == 0x12345678u) out[0] = r;
}
// kind 3: IADD3 carry chains only (alu pipe)
__global__ void k_addchain(uint32_t* out, uint32_t seed) {
    uint32_t acc[4][8], a[8];
    for (int i = 0; i < 8; i++) { a[i] = threadIdx.x * 2654435761u + i + seed; for (int c = 0; c < 4; c++) acc[c][i] = a[i] + c; }
    for (int it = 0; it < ITERS; it++)
#pragma unroll
        for (int c = 0; c < 4; c++) {
            uint32_t* x = acc[c];
            x[0] = add_cc(x[0], a[0]);
#pragma unroll
            for (int j = 1; j < 7; j++) x[j] = addc_cc(x[j], a[j]);
            x[7] = addc(x[7], a[7]);
        }
    uint32_t r = 0; for (int c = 0; c < 4; c++) for (int i = 0; i < 8; i++) r ^= acc[c][i];
    if (r == 0x12345678u) out[0] = r;
}

template <class K>
static double run(const char* name, K launch, double ops_per_thread, int threads) {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0;
    int blocks = prop.multiProcessorCount * (1024 / threads) * 2;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0); launch(blocks, threads); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double v = ops_per_thread * blocks * threads / (ms * 1e-3);
        if (rep) best = v > best ? v : best;
    }
    printf("%-34s threads/blk %4d : %.4g /s  (%.2f per clk per SM @1.965GHz)\n", name, threads, best, best / prop.multiProcessorCount / 1.965e9);
    return best;
}

int main() {
    uint32_t* d; cudaMalloc(&d, 4096);
    for (int threads : {128, 256}) {
        run("64-bit add pairs (product hoisted by ptxas)", [&](int b, int t) { k_wide<<<b, t>>>(d, 1); }, 32.0 * ITERS, threads);
        run("imad.wide acc, evolving operand", [&](int b, int t) { k_wide_dep<<<b, t>>>(d, 1); }, 32.0 * ITERS, threads);
        run("mul.wide + xor, evolving operand", [&](int b, int t) { k_mulwide_dep<<<b, t>>>(d, 1); }, 32.0 * ITERS, threads);
        run("dfma.rz (fp64 pipe)", [&](int b, int t) { k_dfma<<<b, t>>>(d, 1); }, 32.0 * ITERS, threads);
        run("dfma.rz x8 + 4 wide-carry (per dfma)", [&](int b, int t) { k_dfma_imad<<<b, t>>>(d, 1); }, 32.0 * ITERS, threads);
        run("wide carry chain x1 (products)", [&](int b, int t) { k_chain<1><<<b, t>>>(d, 1); }, 32.0 * ITERS, threads);
        run("wide carry chain x2 (products)", [&](int b, int t) { k_chain<2><<<b, t>>>(d, 1); }, 32.0 * ITERS, threads);
        run("wide carry chain x4 (products)", [&](int b, int t) { k_chain<4><<<b, t>>>(d, 1); }, 32.0 * ITERS, threads);
        run("fq_mul x1 chain (muls)", [&](int b, int t) { k_mul<1><<<b, t>>>(d, 1); }, ITERS / 4.0, threads);
        run("fq_mul x2 chains (muls)", [&](int b, int t) { k_mul<2><<<b, t>>>(d, 1); }, ITERS / 2.0, threads);
        run("fq_mul x4 chains (muls)", [&](int b, int t) { k_mul<4><<<b, t>>>(d, 1); }, ITERS * 1.0, threads);
        run("iadd3 carry chains (adds)", [&](int b, int t) { k_addchain<<<b, t>>>(d, 1); }, 32.0 * ITERS, threads);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
