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
static constexpr int kMultMaxBits = 17;      // 64 GB at 1025 generators; the entry index stays below 2^31
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

// Bit lengths of canonical scalar values, for the small-scalar schedule: the encode-time polynomials (comb_ops: addresses,
// timestamps -- sparse_mlpoly_full.rs:155-196) hold values below 2^21 in most of their rows, and a commit over W windows
// costs W additions per scalar whatever the values are.
__device__ __forceinline__ uint32_t fr_bit_length(const Fr* p) {
    Fr s = load_fr(p);
    if (s.is_zero()) return 0;
    s = fp_from_mont(s);
    uint32_t bits = 0;
#pragma unroll
    for (int k = 0; k < 8; k++)
        if (s.l[k]) bits = 32 * k + (32 - __clz(s.l[k]));
    return bits;
}
// out[0] = the SMALLEST bit length among `count` scalars sampled at a fixed stride (full-size inputs show none below ~240)
__global__ void __launch_bounds__(256)
k_min_bits_sample(const Fr* __restrict__ Z, size_t stride, size_t count, uint32_t* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    uint32_t bits = i < count ? fr_bit_length(Z + i * stride) : 0xffffffffu;
    bits = __reduce_min_sync(0xffffffffu, bits);
    if ((threadIdx.x & 31) == 0) atomicMin(out, bits);
}
// block per row: out[row] = the largest bit length of the row's R scalars
__global__ void __launch_bounds__(256)
k_row_max_bits(const Fr* __restrict__ Z, int R, uint32_t* __restrict__ out) {
    __shared__ uint32_t s_best;
    if (threadIdx.x == 0) s_best = 0;
    __syncthreads();
    const Fr* row = Z + (size_t)blockIdx.x * R;
    uint32_t best = 0;
    for (int j = threadIdx.x; j < R; j += 256) best = max(best, fr_bit_length(row + j));
    best = __reduce_max_sync(0xffffffffu, best);
    if ((threadIdx.x & 31) == 0 && best) atomicMax(&s_best, best);
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = s_best;
}

// ------------------------------------------------------------------------------------------------------------------------
// Position-major ("transposed") layout of the sum tree (round 2).
//
// Measured on B200 (microbench/gather_bench.cu, profiles/r2_gather_microbench.txt): 64 B gathers at random over a 32 GiB
// table run at 22.7 G points/s however few bytes each one moves (the 128 B DRAM fetch per point halves with the L2::64B
// qualifier and the rate stays), but at 49-55 G points/s when the 32 lanes of a warp read from ONE 2 MiB region.  Round 1
// of the sum tree is 2 x 16 gathers per scalar, so with the row-major entry list (a warp = 32 neighbouring (window,
// generator) columns of one row = 32 regions) its two kernels sat at 74 % and 85 % of the random-gather rate, not on the
// multiplier.  A column (k, j) of the table -- the 2^(c-1) multiples of 2^(kc) G_j -- IS one such region (2 MiB at c = 16),
// so the lists are now stored position-major:
//
//     entries[p * rp + r]       p = k (R + 1) + j  the (window, generator) position, r the row, rp = rows padded to 32
//     round t:  pair g = q * rp + r  adds the points at positions 2q and 2q + 1 of row r:  in[2g - r], in[2g - r + rp]
//
// and the 32 lanes of a warp are 32 ROWS at the same position: one table column per operand and warp instruction, and
// every other load and store of the rounds is a contiguous run of 32 elements.  Intermediate points are kept as separate
// x and y arrays so that the prefix pass of rounds >= 2 streams x only.
// ------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ Fq load_fq_tab(const Fq* p) {      // table gather: 64 B DRAM fetch instead of the default 128 B
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a, b;
    asm volatile("ld.global.nc.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "l"(q));
    asm volatile("ld.global.nc.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(q + 1));
    Fq r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}

__global__ void __launch_bounds__(256)
k_mult_entries_t(const Fr* __restrict__ Z, const Fr* __restrict__ blinds, int R, int n1, int c, int W, uint32_t stride, int rows,
                 uint32_t rp, uint32_t* __restrict__ entries) {
    const size_t t = (size_t)blockIdx.x * 256 + threadIdx.x;
    const uint32_t r = (uint32_t)(t % rp);
    const uint32_t j = (uint32_t)(t / rp);                             // 0 .. R: scalar index (R = the blind); above: padding
    const uint32_t used = (uint32_t)W * (uint32_t)(R + 1);
    if (j > (uint32_t)R) {
        const uint32_t p = used + (j - (uint32_t)R - 1);
        if (p < stride) entries[(size_t)p * rp + r] = kNullEntry;
        return;
    }
    Fr s;
    bool have = r < (uint32_t)rows;
    if (have) {
        if (j < (uint32_t)R) s = load_fr(Z + (size_t)r * R + j);
        else if (blinds) s = load_fr(blinds + r);
        else have = false;
    }
    if (have && s.is_zero()) have = false;
    if (have) s = fp_from_mont(s);
    const uint32_t col = j < (uint32_t)R ? j : (uint32_t)(n1 - 1);
    uint32_t carry = 0;
#pragma unroll 1
    for (int k = 0; k < W; k++) {
        uint32_t e = kNullEntry;
        if (have) {
            const uint32_t d = signed_window_digit(s.l, k, c, carry);
            if (d) e = ((((uint32_t)k * (uint32_t)n1 + col) << (c - 1)) + ((d & 0x7fffffffu) - 1)) | (d & 0x80000000u);
        }
        entries[((size_t)k * (R + 1) + j) * rp + r] = e;
    }
}

// Denominator of pair g (operands at iP and iP + rp) from the x coordinates alone; y only on the rare path.
template <bool FIRST>
__device__ __forceinline__ bool bat_denominator(uint32_t ex, uint32_t ey, const Affine* __restrict__ table,
                                                const Fq* __restrict__ inx, const Fq* __restrict__ iny, size_t iP, uint32_t rp,
                                                Fq& d) {
    Fq px, qx;
    if (FIRST) {
        if (ex == kNullEntry || ey == kNullEntry) return false;
        px = load_fq_tab(&table[ex & 0x7fffffffu].x);
        qx = load_fq_tab(&table[ey & 0x7fffffffu].x);
    } else {
        px = load_fq(inx + iP);
        qx = load_fq(inx + iP + rp);
    }
    if (px != qx && !px.is_zero() && !qx.is_zero()) { d = fp_sub(qx, px); return true; }
    Affine P, Q;
    P.x = px; Q.x = qx;
    if (FIRST) {
        P.y = load_fq_tab(&table[ex & 0x7fffffffu].y);
        Q.y = load_fq_tab(&table[ey & 0x7fffffffu].y);
        if ((ex >> 31) && !P.is_identity()) P.y = fp_neg(P.y);
        if ((ey >> 31) && !Q.is_identity()) Q.y = fp_neg(Q.y);
    } else {
        P.y = load_fq(iny + iP);
        Q.y = load_fq(iny + iP + rp);
    }
    return ba_classify(P, Q, d) != BA_NONE;
}

// The entry pair of pair g (NULL, NULL past the end)
__device__ __forceinline__ uint2 bat_entries(const uint32_t* __restrict__ entries, size_t g, size_t npairs, uint32_t rp) {
    if (g >= npairs) return make_uint2(kNullEntry, kNullEntry);
    const size_t iP = 2 * g - g % rp;
    return make_uint2(__ldg(entries + iP), __ldg(entries + iP + rp));
}
// L2 prefetch of the table bytes a pair will gather (YTOO: the whole point, else the x coordinate's sector)
template <bool YTOO>
__device__ __forceinline__ void bat_prefetch(const Affine* __restrict__ table, uint2 e) {
    if (e.x != kNullEntry) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(&table[e.x & 0x7fffffffu].x));
        if (YTOO) asm volatile("prefetch.global.L2 [%0];" ::"l"(&table[e.x & 0x7fffffffu].y));
    }
    if (e.y != kNullEntry) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(&table[e.y & 0x7fffffffu].x));
        if (YTOO) asm volatile("prefetch.global.L2 [%0];" ::"l"(&table[e.y & 0x7fffffffu].y));
    }
}

// PF (round 1): the entries of a thread's pairs are read two iterations ahead and their table points prefetched into L2 one
// iteration ahead, so that three gathers per operand are in flight instead of one -- the pass is bound by the latency of
// its gathers, not by the multiplier (ncu: 21 % issue, 36 % DRAM at 47 % occupancy).
template <bool FIRST, bool PF = false>
__global__ void __launch_bounds__(kBaThreads)
k_bat_prefix(const uint32_t* __restrict__ entries, const Affine* __restrict__ table, const Fq* __restrict__ inx,
             const Fq* __restrict__ iny, size_t npairs, uint32_t rp, int B, Fq* __restrict__ prefix, Fq* __restrict__ other,
             Fq* __restrict__ warp_tot) {
    const size_t base = (size_t)blockIdx.x * kBaThreads * B;
    Fq run = fq_one();
    uint2 e0 = make_uint2(kNullEntry, kNullEntry), e1 = e0, e2 = e0;       // entries of iterations j, j + 1, j + 2
    if (FIRST) {
        e0 = bat_entries(entries, base + threadIdx.x, npairs, rp);
        if (PF) {
            if (B > 1) e1 = bat_entries(entries, base + kBaThreads + threadIdx.x, npairs, rp);
            bat_prefetch<false>(table, e1);
        }
    }
#pragma unroll 1
    for (int j = 0; j < B; j++) {
        const size_t g = base + (size_t)j * kBaThreads + threadIdx.x;
        if (FIRST && PF) {
            if (j + 2 < B) e2 = bat_entries(entries, g + 2 * kBaThreads, npairs, rp);
            else e2 = make_uint2(kNullEntry, kNullEntry);
        }
        if (g < npairs) {
            Fq d;
            if (bat_denominator<FIRST>(e0.x, e0.y, table, inx, iny, 2 * g - g % rp, rp, d)) run = fp_mul(run, d);
            store_fq(prefix + g, run);
        }
        if (FIRST) {
            if (PF) {
                e0 = e1; e1 = e2;
                bat_prefetch<false>(table, e1);
            } else if (j + 1 < B) {
                e0 = bat_entries(entries, g + kBaThreads, npairs, rp);
            }
        }
    }
    // Every thread needs the product of the OTHER threads' totals of its block (times the inverse of the block total, that
    // is its own total's inverse).  One warp computes all 256 of them through shared memory -- lane l owns the totals of
    // threads l, l + 32, ..., l + 224: 8 running products forward, a shuffle scan over the 32 lane totals, 16 products
    // backward -- 35 multiplications of one warp per block instead of 12 of every warp, and one inversion per block instead
    // of one per warp.
    __shared__ uint4 s_tot[kBaThreads * 2], s_exc[kBaThreads * 2];
    s_tot[2 * threadIdx.x] = make_uint4(run.l[0], run.l[1], run.l[2], run.l[3]);
    s_tot[2 * threadIdx.x + 1] = make_uint4(run.l[4], run.l[5], run.l[6], run.l[7]);
    __syncthreads();
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x;
    auto lds = [](const uint4* a, int i) {
        const uint4 u = a[2 * i], v = a[2 * i + 1];
        Fq r;
        r.l[0] = u.x; r.l[1] = u.y; r.l[2] = u.z; r.l[3] = u.w; r.l[4] = v.x; r.l[5] = v.y; r.l[6] = v.z; r.l[7] = v.w;
        return r;
    };
    Fq acc = fq_one();
#pragma unroll 1
    for (int i = 0; i < kBaThreads / 32; i++) {
        const int t = i * 32 + lane;
        s_exc[2 * t] = make_uint4(acc.l[0], acc.l[1], acc.l[2], acc.l[3]);           // product of this lane's totals before t
        s_exc[2 * t + 1] = make_uint4(acc.l[4], acc.l[5], acc.l[6], acc.l[7]);
        acc = fp_mul(acc, lds(s_tot, t));
    }
    Fq pre = acc, suf = acc;
#pragma unroll 1
    for (int off = 1; off < 32; off <<= 1) {
        Fq a = shfl_fq(pre, (lane - off) & 31), b = shfl_fq(suf, (lane + off) & 31);
        if (lane < off) a = fq_one();
        if (lane + off >= 32) b = fq_one();
        pre = fp_mul(pre, a);
        suf = fp_mul(suf, b);
    }
    Fq pe = shfl_fq(pre, (lane - 1) & 31), se = shfl_fq(suf, (lane + 1) & 31);
    if (lane == 0) pe = fq_one();
    if (lane == 31) se = fq_one();
    acc = fp_mul(pe, se);                                                          // the other lanes' totals
    Fq* oth = other + (size_t)blockIdx.x * kBaThreads;
#pragma unroll 1
    for (int i = kBaThreads / 32 - 1; i >= 0; i--) {
        const int t = i * 32 + lane;
        store_fq(oth + t, fp_mul(acc, lds(s_exc, t)));                             // others of thread t
        acc = fp_mul(acc, lds(s_tot, t));
    }
    if (lane == 31) store_fq(warp_tot + blockIdx.x, pre);                          // the block's total
}

// Round 1 walks the blocks in reverse: the prefix pass has just pulled the table points of the LAST blocks through L2, so
// the finish pass meets them there before they are evicted.
template <bool FIRST, int MINB = 3, bool PF = false>
__global__ void __launch_bounds__(kBaThreads, MINB)
k_bat_finish(const uint32_t* __restrict__ entries, const Affine* __restrict__ table, const Fq* __restrict__ inx,
             const Fq* __restrict__ iny, size_t npairs, uint32_t rp, int B, const Fq* __restrict__ prefix,
             const Fq* __restrict__ other, const Fq* __restrict__ warp_inv, Fq* __restrict__ outx, Fq* __restrict__ outy) {
    const size_t blk = FIRST ? (size_t)(gridDim.x - 1 - blockIdx.x) : (size_t)blockIdx.x;
    const size_t base = blk * kBaThreads * B;
    const size_t tid_global = blk * kBaThreads + threadIdx.x;
    uint2 e0 = make_uint2(kNullEntry, kNullEntry), e1 = e0;                // entries of iterations j and j - 1
    if (FIRST) {
        e0 = bat_entries(entries, base + (size_t)(B - 1) * kBaThreads + threadIdx.x, npairs, rp);
        if (PF && B > 1) e1 = bat_entries(entries, base + (size_t)(B - 2) * kBaThreads + threadIdx.x, npairs, rp);
    }
    Fq run = fp_mul(load_fq(warp_inv + blk), load_fq(other + tid_global));   // (own total)^-1: block total^-1 x the others
#pragma unroll 1
    for (int j = B - 1; j >= 0; j--) {
        const size_t g = base + (size_t)j * kBaThreads + threadIdx.x;
        const uint32_t ex = e0.x, ey = e0.y;
        if (FIRST) {       // rotate the entry pipeline before the long body: loads issued here return under the additions
            if (PF) {       // entries of j - 1 arrived an iteration ago: their table points start towards L2 now
                e0 = e1;
                bat_prefetch<true>(table, e0);
                e1 = j >= 2 ? bat_entries(entries, g - 2 * kBaThreads, npairs, rp) : make_uint2(kNullEntry, kNullEntry);
            } else if (j >= 1) {
                e0 = bat_entries(entries, g - kBaThreads, npairs, rp);
            }
        }
        if (g >= npairs) continue;
        const size_t iP = 2 * g - g % rp;
        Affine P, Q;
        if (FIRST) {
            if (ex == kNullEntry) P = Affine::identity();
            else { P.x = load_fq_tab(&table[ex & 0x7fffffffu].x); P.y = load_fq_tab(&table[ex & 0x7fffffffu].y); }
            if (ey == kNullEntry) Q = Affine::identity();
            else { Q.x = load_fq_tab(&table[ey & 0x7fffffffu].x); Q.y = load_fq_tab(&table[ey & 0x7fffffffu].y); }
            if (ex != kNullEntry && (ex >> 31) && !P.is_identity()) P.y = fp_neg(P.y);
            if (ey != kNullEntry && (ey >> 31) && !Q.is_identity()) Q.y = fp_neg(Q.y);
        } else {
            P.x = load_fq(inx + iP); P.y = load_fq(iny + iP);
            Q.x = load_fq(inx + iP + rp); Q.y = load_fq(iny + iP + rp);
        }
        Fq d;
        const int kind = ba_classify(P, Q, d);
        Affine S;
        if (kind == BA_NONE) {
            if (P.is_identity()) S = Q;
            else if (Q.is_identity()) S = P;
            else S = Affine::identity();
        } else {
            Fq inv_d = run;
            if (j > 0) inv_d = fp_mul(run, load_fq(prefix + (g - kBaThreads)));
            run = fp_mul(run, d);
            Fq num;
            if (kind == BA_ADD) {
                num = fp_sub(Q.y, P.y);
            } else {
                const Fq xx = fp_mul(P.x, P.x);
                num = fp_add(fp_dbl(xx), xx);
            }
            const Fq lambda = fp_mul(num, inv_d);
            const Fq x3 = fp_sub(fp_sub(fp_mul(lambda, lambda), P.x), Q.x);
            S.x = x3;
            S.y = fp_sub(fp_mul(lambda, fp_sub(P.x, x3)), P.y);
        }
        store_fq(outx + g, S.x);
        store_fq(outy + g, S.y);
    }
}

// ------------------------------------------------------------------------------------------------------------------------
// Fused rounds ("layout 2").  Ablation of the pipeline above (profiles/r2_ablation_cfg1.txt): of 2.62 ms per cfg1 commit the
// finish passes take 1.42, the prefix passes 0.77 -- for a sixth of the multiplications: standing alone they are a serial
// chain of one product per pair behind two loads, bound by latency, not by the multiplier.  Here the prefix pass of round
// k + 1 rides inside the finish pass of round k: a thread owns ONE ROW and a run of B consecutive pair positions of it, so the
// two sums that form a pair of the next round come out of the same thread one after the other, and their denominator goes
// into the next round's running product on the spot (one more product per two additions, in a kernel that is already on
// the multiplier).  B halves from round to round (32, 16, ..., 1 for six rounds) while the threads stay the same, and the
// walk alternates direction: a round consumes its running products last-to-first, which is the order the next round's
// products are built in.  Only round 1 keeps a prefix kernel of its own (its operands are table gathers).
//   warp w of the grid = (position block w / (rp / 32), rows 32 (w % (rp / 32)) ...); chain of thread = positions pb B ... pb B + B - 1
//   accumulation step t of a chain <-> position  q0 + t (ascending rounds: 1, 3, 5)  or  q0 + B - 1 - t (descending: 2, 4, 6)
//   prefix[q rp + r] = product of the chain's denominators BEFORE that pair (exclusive)
// ------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ Fq lds_fq(const uint4* a, int i) {
    const uint4 u = a[2 * i], v = a[2 * i + 1];
    Fq r;
    r.l[0] = u.x; r.l[1] = u.y; r.l[2] = u.z; r.l[3] = u.w; r.l[4] = v.x; r.l[5] = v.y; r.l[6] = v.z; r.l[7] = v.w;
    return r;
}
__device__ __forceinline__ void sts_fq(uint4* a, int i, const Fq& v) {
    a[2 * i] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    a[2 * i + 1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
// other[t] = product of the other threads' totals of this block, *total = the block's total (see k_bat_prefix)
__device__ __forceinline__ void bat_block_scan(const Fq& run, Fq* __restrict__ other_block, Fq* __restrict__ total, uint4* s_tot,
                                               uint4* s_exc) {
    __syncthreads();                                   // the caller may still be reading the shared arrays it aliases
    sts_fq(s_tot, threadIdx.x, run);
    __syncthreads();
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x;
    Fq acc = fq_one();
#pragma unroll 1
    for (int i = 0; i < kBaThreads / 32; i++) {
        const int t = i * 32 + lane;
        sts_fq(s_exc, t, acc);
        acc = fp_mul(acc, lds_fq(s_tot, t));
    }
    Fq pre = acc, suf = acc;
#pragma unroll 1
    for (int off = 1; off < 32; off <<= 1) {
        Fq a = shfl_fq(pre, (lane - off) & 31), b = shfl_fq(suf, (lane + off) & 31);
        if (lane < off) a = fq_one();
        if (lane + off >= 32) b = fq_one();
        pre = fp_mul(pre, a);
        suf = fp_mul(suf, b);
    }
    Fq pe = shfl_fq(pre, (lane - 1) & 31), se = shfl_fq(suf, (lane + 1) & 31);
    if (lane == 0) pe = fq_one();
    if (lane == 31) se = fq_one();
    acc = fp_mul(pe, se);
#pragma unroll 1
    for (int i = kBaThreads / 32 - 1; i >= 0; i--) {
        const int t = i * 32 + lane;
        store_fq(other_block + t, fp_mul(acc, lds_fq(s_exc, t)));
        acc = fp_mul(acc, lds_fq(s_tot, t));
    }
    if (lane == 31) store_fq(total, pre);
}

struct BatSlot { uint32_t r, pb; bool valid; };
__device__ __forceinline__ BatSlot bat2_slot(uint32_t rp, uint32_t npb) {
    const uint32_t w = blockIdx.x * (kBaThreads / 32) + (threadIdx.x >> 5), groups = rp >> 5;
    BatSlot s;
    s.pb = w / groups;
    s.r = (w % groups) * 32 + (threadIdx.x & 31);
    s.valid = s.pb < npb;
    return s;
}

// Round 1's running products: chain position q0 + t, operands entries[(2q) rp + r], entries[(2q + 1) rp + r]
__global__ void __launch_bounds__(kBaThreads)
k_bat2_prefix1(const uint32_t* __restrict__ entries, const Affine* __restrict__ table, uint32_t npb, uint32_t rp, int B,
               Fq* __restrict__ prefix, Fq* __restrict__ other, Fq* __restrict__ block_tot) {
    __shared__ uint4 s_tot[kBaThreads * 2], s_exc[kBaThreads * 2];
    const BatSlot sl = bat2_slot(rp, npb);
    Fq run = fq_one();
    if (sl.valid) {
        const uint32_t q0 = sl.pb * (uint32_t)B;
        size_t iP = (size_t)2 * q0 * rp + sl.r;
        uint32_t ex = __ldg(entries + iP), ey = __ldg(entries + iP + rp);
#pragma unroll 1
        for (int t = 0; t < B; t++) {
            const size_t g = (size_t)(q0 + t) * rp + sl.r;
            const uint32_t cx = ex, cy = ey;
            if (t + 1 < B) {                           // next pair's entries: in flight under this pair's gathers and product
                iP += (size_t)2 * rp;
                ex = __ldg(entries + iP);
                ey = __ldg(entries + iP + rp);
            }
            store_fq(prefix + g, run);
            Fq d;
            if (bat_denominator<true>(cx, cy, table, nullptr, nullptr, 0, rp, d)) run = fp_mul(run, d);
        }
    }
    bat_block_scan(run, other + (size_t)blockIdx.x * kBaThreads, block_tot + blockIdx.x, s_tot, s_exc);
}

// One round: the additions of this round's pairs (walking the chain last-to-first) and, when EMIT, the running products of
// the next round's pairs.  ASC: this round's chain was accumulated in ascending position order.
template <bool FIRST, bool EMIT, int MINB>
__global__ void __launch_bounds__(kBaThreads, MINB)
k_bat2_round(const uint32_t* __restrict__ entries, const Affine* __restrict__ table, const Fq* __restrict__ inx,
             const Fq* __restrict__ iny, uint32_t npb, uint32_t rp, int B, int asc, const Fq* __restrict__ prefix,
             const Fq* __restrict__ other, const Fq* __restrict__ block_inv, Fq* __restrict__ outx, Fq* __restrict__ outy,
             Fq* __restrict__ prefix_next, Fq* __restrict__ other_next, Fq* __restrict__ block_tot_next) {
    __shared__ uint4 s_a[kBaThreads * 2], s_b[kBaThreads * 2];       // previous sum of the thread (x, y); then the scan's scratch
    const BatSlot sl = bat2_slot(rp, npb);
    Fq run2 = fq_one();
    if (sl.valid) {
        const uint32_t q0 = sl.pb * (uint32_t)B;
        const size_t u = (size_t)blockIdx.x * kBaThreads + threadIdx.x;
        Fq run = fp_mul(load_fq(block_inv + blockIdx.x), load_fq(other + u));      // (own chain total)^-1
#pragma unroll 1
        for (int t = B - 1; t >= 0; t--) {
            const uint32_t q = asc ? q0 + (uint32_t)t : q0 + (uint32_t)(B - 1 - t);
            const size_t g = (size_t)q * rp + sl.r, iP = 2 * g - sl.r;
            Affine P, Q;
            if (FIRST) {
                const uint32_t ex = __ldg(entries + iP), ey = __ldg(entries + iP + rp);
                if (ex == kNullEntry) P = Affine::identity();
                else { P.x = load_fq_tab(&table[ex & 0x7fffffffu].x); P.y = load_fq_tab(&table[ex & 0x7fffffffu].y); }
                if (ey == kNullEntry) Q = Affine::identity();
                else { Q.x = load_fq_tab(&table[ey & 0x7fffffffu].x); Q.y = load_fq_tab(&table[ey & 0x7fffffffu].y); }
                if (ex != kNullEntry && (ex >> 31) && !P.is_identity()) P.y = fp_neg(P.y);
                if (ey != kNullEntry && (ey >> 31) && !Q.is_identity()) Q.y = fp_neg(Q.y);
            } else {
                P.x = load_fq(inx + iP); P.y = load_fq(iny + iP);
                Q.x = load_fq(inx + iP + rp); Q.y = load_fq(iny + iP + rp);
            }
            Fq d;
            const int kind = ba_classify(P, Q, d);
            Affine S;
            if (kind == BA_NONE) {
                if (P.is_identity()) S = Q;
                else if (Q.is_identity()) S = P;
                else S = Affine::identity();
            } else {
                Fq inv_d = run;
                if (t > 0) inv_d = fp_mul(run, load_fq(prefix + g));     // exclusive running product of this chain
                run = fp_mul(run, d);
                Fq num;
                if (kind == BA_ADD) {
                    num = fp_sub(Q.y, P.y);
                } else {
                    const Fq xx = fp_mul(P.x, P.x);
                    num = fp_add(fp_dbl(xx), xx);
                }
                const Fq lambda = fp_mul(num, inv_d);
                const Fq x3 = fp_sub(fp_sub(fp_mul(lambda, lambda), P.x), Q.x);
                S.x = x3;
                S.y = fp_sub(fp_mul(lambda, fp_sub(P.x, x3)), P.y);
            }
            store_fq(outx + g, S.x);
            store_fq(outy + g, S.y);
            if (EMIT) {
                if (t & 1) {                                   // first of the two sums of a next-round pair: park it
                    sts_fq(s_a, threadIdx.x, S.x);
                    sts_fq(s_b, threadIdx.x, S.y);
                } else {                                       // second: positions q and q +- 1 -> pair q / 2 of the next round
                    Affine T;
                    T.x = lds_fq(s_a, threadIdx.x);
                    T.y = lds_fq(s_b, threadIdx.x);
                    // the next round reads its pair as (position 2q', position 2q' + 1): keep that operand order
                    const bool s_is_low = (q & 1) == 0;
                    Fq d2;
                    const int k2 = s_is_low ? ba_classify(S, T, d2) : ba_classify(T, S, d2);
                    store_fq(prefix_next + (size_t)(q >> 1) * rp + sl.r, run2);
                    if (k2 != BA_NONE) run2 = fp_mul(run2, d2);
                }
            }
        }
    }
    if (EMIT) bat_block_scan(run2, other_next + (size_t)blockIdx.x * kBaThreads, block_tot_next + blockIdx.x, s_a, s_b);
}

// warp per row: totals[row] = sum of the row's `cnt` points, point i of row r at [i * rp + r] of the x and y arrays.
// (A block per row with a shared-memory tree -- 9 additions deep instead of 13 -- was measured and is slower, 2.66 against
// 2.62 ms per cfg1 commit: the kernel's cost is the ~256 XYZZ additions per row, not their depth.)
// WPR warps per row (1, 2 or 4; a block is 4 warps): every lane adds cnt / (32 WPR) points, a shuffle tree adds the lanes,
// the warps of a row meet in shared memory.  The additions are inlined -- the out-of-line forms keep the accumulator in
// local memory, which cost 2.68 against 2.56 ms per cfg1 commit.
template <int WPR>
__global__ void __launch_bounds__(kMultSumThreads)
k_mult_sum_rows_t(const Fq* __restrict__ ptsx, const Fq* __restrict__ ptsy, uint32_t cnt, uint32_t rp, int rows,
                  XYZZ* __restrict__ totals) {
    __shared__ XYZZ s_w[kMultSumThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row = (int)(blockIdx.x * (unsigned)(kMultSumThreads / 32 / WPR)) + warp / WPR;
    const int part = warp % WPR;
    XYZZ acc = XYZZ::identity();
    if (row < rows) {
#pragma unroll 1
        for (uint32_t i = (uint32_t)(part * 32 + lane); i < cnt; i += 32 * WPR) {
            Affine p;
            p.x = load_fq(ptsx + (size_t)i * rp + row);
            p.y = load_fq(ptsy + (size_t)i * rp + row);
            if (p.is_identity()) continue;
            xyzz_add_mixed<MulInline>(acc, p);
        }
    }
    // the shuffle tree runs on every warp of the block (rows past the end hold the identity): both lanes of a pair share the
    // products of each addition (small_kernels.cuh, xyzz_pair_level)
#pragma unroll 1
    for (int stride = 16; stride >= 1; stride >>= 1) xyzz_pair_level(&acc, stride);
    if (WPR == 1) {
        if (lane == 0 && row < rows) store_xyzz(totals + row, acc);
        return;
    }
    if (lane == 0) s_w[warp] = acc;
    __syncthreads();
    if (part == 0 && lane == 0 && row < rows) {
#pragma unroll 1
        for (int k = 1; k < WPR; k++) xyzz_add<MulInline, MulCall>(acc, s_w[warp + k]);
        store_xyzz(totals + row, acc);
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
