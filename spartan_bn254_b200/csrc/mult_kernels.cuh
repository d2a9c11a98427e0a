// Row commitments as sums of tabulated digit multiples (sm_100a) -- the many-row counterpart of small_kernels.cuh.
//
// The bucket method pays, per (scalar, window) entry, one addition into a bucket and then 2 * 2^(c-1) additions per row to
// weigh the buckets; a batched-affine addition costs 6 products, an XYZZ one 10-14.  With the generators fixed and 180 GB
// of HBM, every digit multiple can be tabulated instead,
//
//     mult[((k * n1 + j) << (c-1)) + d - 1] = d * 2^(k c) * G_j          (k < W = ceil(255 / c), 1 <= d <= 2^(c-1), affine)
//
// (5.4 GB for 1025 generators at c = 13, built once per generator set), and a row commitment becomes a PLAIN SUM of one
// table point per entry: no buckets, no sort, no bucket reduction, and every addition of the sum is a batched-affine one.
// Per chunk of rows:
//   k_mult_entries     thread per scalar: Montgomery -> canonical, signed c-bit digits, entries[row][k][j] = table index |
//                      sign << 31 (NULL for a zero digit and for the padding up to the row stride)
//   r rounds of        k_ba_prefix / k_ba_invert / k_ba_finish (ba_kernels.cuh) over the flat pair array: round 1 reads the
//                      table through the entries, round i the points of round i - 1; the row stride is a multiple of 2^r,
//                      so a pair never straddles two rows
//   k_mult_sum_rows    warp per row: the stride / 2^r points left are added up in XYZZ (mixed additions + a shuffle tree)
//   k_normalize        as everywhere else
// Cost per scalar: W batched-affine additions (6 products each) against W bucket additions (6-10) plus 2 * 2^(c-1) * 14 / n
// for the bucket reduction; W itself shrinks because c may be larger than a bucket set could afford (13 against 11 at 1024
// generators).  The table is read at random (64 B per entry): HBM traffic the bucket method did not have, 1-2 GB per commit.
#pragma once
#include "ba_kernels.cuh"
#include "digits.cuh"
#include "small_kernels.cuh"

namespace sbn {

static constexpr int kMultChunk = 128;       // multiples filled by one thread of the table builder
static constexpr int kMultMaxBits = 16;
static constexpr int kMultSumThreads = 128;

// thread per (k, j, chunk): B = 2^(k c) * base_j, S = (chunk * 128) * B, then 128 times S += B, each stored in affine form
__global__ void k_mult_fill(const Affine* __restrict__ bases, int n1, int c, int W, Affine* __restrict__ mult) {
    const int nchunk = (1 << (c - 1)) / kMultChunk;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)W * n1 * nchunk) return;
    const int chunk = (int)(t % nchunk);
    const size_t kj = t / nchunk;            // k * n1 + j
    const int k = (int)(kj / n1), j = (int)(kj % n1);
    Affine* out = mult + (kj << (c - 1)) + (size_t)chunk * kMultChunk;
    const Affine p = load_affine(bases + j);
    if (p.is_identity()) {
        for (int i = 0; i < kMultChunk; i++) store_affine(out + i, Affine::identity());
        return;
    }
    XYZZ acc = XYZZ::from_affine(p);
    for (int d = 0; d < k * c; d++) acc = xyzz_dbl<MulCall>(acc);
    const Affine B = xyzz_to_affine<MulCall>(acc);
    // S = chunk * B (double-and-add, MSB first), then * 128
    XYZZ S = XYZZ::identity();
    for (int bit = 15; bit >= 0; bit--) {
        S = xyzz_dbl<MulCall>(S);
        if ((chunk >> bit) & 1) xyzz_add_mixed<MulCall>(S, B);
    }
    for (int d = 0; d < 7; d++) S = xyzz_dbl<MulCall>(S);
    for (int i = 0; i < kMultChunk; i++) {
        xyzz_add_mixed<MulCall>(S, B);
        store_affine(out + i, xyzz_to_affine<MulCall>(S));
    }
}

// Signed c-bit digit k of a canonical scalar needs the carry of the digits below it: all W digits of one scalar are
// produced by one thread.  entries[row * stride + k * (R + 1) + j]; positions from W * (R + 1) to stride are NULL.
__global__ void __launch_bounds__(256)
k_mult_entries(const Fr* __restrict__ Z, const Fr* __restrict__ blinds, int R, int n1, int c, int W, uint32_t stride,
               uint32_t* __restrict__ entries) {
    const int row = blockIdx.y;
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;        // 0 .. R: scalar index (R = the blind)
    uint32_t* erow = entries + (size_t)row * stride;
    const uint32_t used = (uint32_t)W * (uint32_t)(R + 1);
    if (j > (uint32_t)R) {                                             // the threads past the scalars write the padding
        const uint32_t p = used + (j - (uint32_t)R - 1);
        if (p < stride) erow[p] = kNullEntry;
        return;
    }
    Fr s;
    bool have = true;
    if (j < (uint32_t)R) s = load_fr(Z + (size_t)row * R + j);
    else if (blinds) s = load_fr(blinds + row);
    else have = false;
    if (have && s.is_zero()) have = false;
    if (have) s = fp_from_mont(s);
    const uint32_t col = j < (uint32_t)R ? j : (uint32_t)(n1 - 1);
    uint32_t carry = 0;
#pragma unroll 1
    for (int k = 0; k < W; k++) {
        uint32_t e = kNullEntry;
        if (have) {
            const uint32_t d = signed_window_digit(s.l, k, c, carry);      // |d| with the sign in bit 31, 0 for a zero digit
            if (d) e = ((((uint32_t)k * (uint32_t)n1 + col) << (c - 1)) + ((d & 0x7fffffffu) - 1)) | (d & 0x80000000u);
        }
        erow[(size_t)k * (R + 1) + j] = e;
    }
}

// warp per row: totals[row] = sum of the row's `cnt` affine points.  Lanes add cnt / 32 points each (mixed additions), a
// shuffle tree adds the lanes -- a tree level costs a full warp's issue slots however few lanes carry a value, so the tree
// is kept to the five levels of one warp.
__global__ void __launch_bounds__(kMultSumThreads)
k_mult_sum_rows(const Affine* __restrict__ pts, uint32_t cnt, int rows, XYZZ* __restrict__ totals) {
    const int lane = threadIdx.x & 31;
    const int row = (int)((blockIdx.x * (unsigned)kMultSumThreads + threadIdx.x) >> 5);
    if (row >= rows) return;
    const Affine* prow = pts + (size_t)row * cnt;
    XYZZ acc = XYZZ::identity();
    for (uint32_t i = lane; i < cnt; i += 32) {
        const Affine p = load_affine(prow + i);
        if (!p.is_identity()) xyzz_add_mixed_call(&acc, &p);
    }
    for (int stride = 16; stride >= 1; stride >>= 1) {
        XYZZ o = shfl_xyzz(acc, (lane + stride) & 31);
        if (lane >= stride) o = XYZZ::identity();
        xyzz_add_call(&acc, &o);
    }
    if (lane == 0) store_xyzz(totals + row, acc);
}

}  // namespace sbn
