// Short commitments (sm_100a): <[Scalar] as Commitments>::commit (commitments.rs:118-154) over generator sets of at most
// kSmallMaxCols points -- the 1- to 5-scalar vectors of the ZK sumchecks and Sigma-protocols (sumcheck.rs:559-649,
// nizk/mod.rs:30-400), of which a proof makes a few hundred, each on the Fiat-Shamir critical path.
//
// A row of the general commit pipeline costs six launches, two copies and ~0.3 ms of latency whatever its length.  For a
// short generator set ALL digit multiples are tabulated once,
//
//     small[(k * n_cols + j) * 128 + d - 1] = d * 2^(8k) * P_j        (k < 32 windows of 8 bits, 1 <= d <= 128, affine)
//
// (4 MiB at 16 generators), so a commitment is the sum of 32 table points per scalar: one launch, one CTA per row, warp j
// takes scalar j (the last warp the blind and h), lane k window k; a shuffle tree adds the 32 points of a warp, the first
// warp adds the warps' sums and normalises.  The critical path is 5 + log2(n_cols) additions and one inversion; scalars and
// results travel through mapped pinned memory, so the call is one launch and one stream synchronisation.
#pragma once
#include "msm_kernels.cuh"

namespace sbn {

static constexpr int kSmallMaxCols = 16;    // generators of the set, h (and gens_1's G) included
static constexpr int kSmallC = 8, kSmallW = 32, kSmallD = 128;
static constexpr int kSmallMaxRows = 64;    // rows per call through the mapped staging buffer

// One thread per (window k, generator j): B = 2^(8k) P_j, then B, 2B, ..., 128B, each normalised.
__global__ void k_build_small_table(const Affine* __restrict__ orig, int n_cols, Affine* __restrict__ table) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= kSmallW * n_cols) return;
    const int k = t / n_cols, j = t % n_cols;
    const Affine p = load_affine(orig + j);
    Affine* out = table + (size_t)t * kSmallD;
    if (p.is_identity()) {
        for (int d = 0; d < kSmallD; d++) store_affine(out + d, Affine::identity());
        return;
    }
    XYZZ acc = XYZZ::from_affine(p);
    for (int d = 0; d < kSmallC * k; d++) acc = xyzz_dbl<MulCall>(acc);
    const Affine base = xyzz_to_affine<MulCall>(acc);
    acc = XYZZ::from_affine(base);
    store_affine(out, base);
    for (int d = 2; d <= kSmallD; d++) {
        xyzz_add_mixed<MulCall>(acc, base);     // d = 2 takes the doubling branch
        store_affine(out + d - 1, xyzz_to_affine<MulCall>(acc));
    }
}

// blockDim.x = 32 * (R + 1).  Z, blinds, out, inf may be mapped host memory.
__global__ void __launch_bounds__(32 * kSmallMaxCols)
k_small_commit(const Fr* __restrict__ Z, const Fr* __restrict__ blinds, int R, int n_cols, const Affine* __restrict__ table,
               Affine* __restrict__ out, uint8_t* __restrict__ inf) {
    __shared__ XYZZ part[kSmallMaxCols];
    const int row = blockIdx.x, lane = threadIdx.x & 31, j = threadIdx.x >> 5, nw = blockDim.x >> 5;
    XYZZ acc = XYZZ::identity();
    const bool is_blind = j >= R;
    if (!is_blind || blinds) {
        const Fr s = fp_from_mont(is_blind ? load_fr(blinds + row) : load_fr(Z + (size_t)row * R + j));
        const int col = is_blind ? n_cols - 1 : j;
        uint32_t mine = 0;
        bool neg = false;
        for_each_digit<kSmallC>(s, [&](int k, uint32_t dm1, bool negative) {
            if (k == lane) { mine = dm1 + 1; neg = negative; }
        });
        if (mine) {
            Affine p = load_affine(table + ((size_t)(lane * n_cols + col) * kSmallD + (mine - 1)));
            if (neg) p = affine_neg(p);
            acc = XYZZ::from_affine(p);
        }
    }
    for (int stride = 16; stride >= 1; stride >>= 1) {
        XYZZ o = shfl_xyzz(acc, (lane + stride) & 31);
        if (lane >= stride) o = XYZZ::identity();
        xyzz_add_call(&acc, &o);
    }
    if (lane == 0) part[j] = acc;
    __syncthreads();
    if (j != 0) return;
    XYZZ v = lane < nw ? part[lane] : XYZZ::identity();
    for (int stride = kSmallMaxCols / 2; stride >= 1; stride >>= 1) {
        XYZZ o = shfl_xyzz(v, (lane + stride) & 31);
        if (lane >= stride) o = XYZZ::identity();
        xyzz_add_call(&v, &o);
    }
    if (lane == 0) {
        const Affine a = xyzz_to_affine<MulInline>(v);
        store_affine(out + row, a);
        inf[row] = v.is_identity() ? 1 : 0;
    }
}

}  // namespace sbn
