// CUDA kernels of the batched fixed-base MSM behind Hyrax row commitments (sm_100a).
//
// One Hyrax commit = L_size independent MSMs over the SAME R_size+1 bases (reference
// hyrax.rs:253-267 fans rows out with rayon; commitments.rs:144-154 appends blind*h).  Because the
// bases never change, the window tables 2^(k*c) * G_j are precomputed once (k_build_tables), which
// collapses the per-window bucket sets of Pippenger into ONE bucket set per row:
//
//   K1+K2  k_sort_row      per row: Montgomery Fr -> canonical, signed c-bit digits, counting sort of
//                          (window, base) entries by bucket in shared memory, buckets ranked by size
//   K3     k_accumulate    one thread per (row, bucket): XYZZ mixed additions of table points
//   K4a    k_reduce        per row: sum_b (b+1) * S_b by chunked running sums + shared-memory tree
//   K4b    k_normalize     XYZZ -> affine (one Fq inversion per row)
//
// All group arithmetic is 8x32-bit-limb Montgomery Fq (fp.cuh / ec.cuh); no tensor cores (modular
// integer arithmetic), no floating point.
#pragma once
#include <cuda_runtime.h>
#include "ec.cuh"

namespace sbn {

static constexpr int kMaxWindowBits = 13;
static constexpr int kMinWindowBits = 4;

__host__ __device__ inline int msm_num_windows(int c) { return (254 + c) / c; }   // ceil(255 / c)

// ---------------------------------------------------------------------------------------------
// vectorised 128-bit loads / stores
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ Fr load_fr(const Fr* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    Fr r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ Affine load_affine(const Affine* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2), d = __ldg(q + 3);
    Affine r;
    r.x.l[0] = a.x; r.x.l[1] = a.y; r.x.l[2] = a.z; r.x.l[3] = a.w;
    r.x.l[4] = b.x; r.x.l[5] = b.y; r.x.l[6] = b.z; r.x.l[7] = b.w;
    r.y.l[0] = c.x; r.y.l[1] = c.y; r.y.l[2] = c.z; r.y.l[3] = c.w;
    r.y.l[4] = d.x; r.y.l[5] = d.y; r.y.l[6] = d.z; r.y.l[7] = d.w;
    return r;
}
__device__ __forceinline__ void store_fq(Fq* p, const Fq& v) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
__device__ __forceinline__ Fq load_fq(const Fq* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    Fq r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ void store_xyzz(XYZZ* p, const XYZZ& v) {
    store_fq(&p->X, v.X); store_fq(&p->Y, v.Y); store_fq(&p->ZZ, v.ZZ); store_fq(&p->ZZZ, v.ZZZ);
}
__device__ __forceinline__ XYZZ load_xyzz(const XYZZ* p) {
    XYZZ r;
    r.X = load_fq(&p->X); r.Y = load_fq(&p->Y); r.ZZ = load_fq(&p->ZZ); r.ZZZ = load_fq(&p->ZZZ);
    return r;
}
__device__ __forceinline__ void store_affine(Affine* p, const Affine& v) { store_fq(&p->x, v.x); store_fq(&p->y, v.y); }

// ---------------------------------------------------------------------------------------------
// window tables: table[k * n1 + j] = 2^(k*c) * base_j   (affine; (0,0) = identity)
// ---------------------------------------------------------------------------------------------
__global__ void k_build_tables(const Affine* __restrict__ bases, const uint8_t* __restrict__ inf, int n1, int c, int W,
                               Affine* __restrict__ table) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n1) return;
    Affine p = load_affine(bases + j);
    if (inf && inf[j]) p = Affine::identity();
    store_affine(table + j, p);
    XYZZ acc = XYZZ::from_affine(p);
    for (int k = 1; k < W; k++) {
        for (int d = 0; d < c; d++) acc = xyzz_dbl(acc);
        p = xyzz_to_affine(acc);
        store_affine(table + (size_t)k * n1 + j, p);
        acc = XYZZ::from_affine(p);
    }
}

// ---------------------------------------------------------------------------------------------
// K1 + K2: digits and per-row counting sort
// ---------------------------------------------------------------------------------------------
// Signed digits d_k in [-2^(c-1), 2^(c-1)] with sum d_k 2^(kc) = s; calls f(k, |d|-1, negative).
template <int C, class Fn>
__device__ __forceinline__ void for_each_digit(const Fr& canon, Fn&& f) {
    constexpr int W = (254 + C) / C;
    uint32_t carry = 0;
#pragma unroll
    for (int k = 0; k < W; k++) {
        constexpr uint32_t mask = (1u << C) - 1;
        const int bit = k * C;
        const int limb = bit >> 5, off = bit & 31;
        uint32_t raw = canon.l[limb] >> off;
        if (off + C > 32 && limb + 1 < 8) raw |= canon.l[limb + 1] << (32 - off);
        raw &= mask;
        uint32_t d = raw + carry;
        carry = 0;
        bool negative = false;
        if (d > (1u << (C - 1))) { d = (1u << C) - d; carry = 1; negative = true; }
        if (d != 0) f(k, d - 1, negative);
    }
}

static constexpr int kSortThreads = 256;
static constexpr int kRankBins = 256;

// entries[row * E + ...] : bucket-sorted list of (k * n1 + j) | (negative << 31)
// starts [row * (NB + 1) + b] : first entry of bucket b (starts[NB] = total)
// order  [row * NB + rank]   : bucket ids by decreasing size
template <int C>
__global__ void __launch_bounds__(kSortThreads)
k_sort_row(const Fr* __restrict__ Z, const Fr* __restrict__ blinds, int R, int scalars_are_mont,
           uint32_t E, uint32_t* __restrict__ entries, uint32_t* __restrict__ starts, uint16_t* __restrict__ order) {
    constexpr int NB = 1 << (C - 1);
    constexpr int PER = (NB + kSortThreads - 1) / kSortThreads;   // buckets per thread in the scans
    __shared__ uint32_t counts[NB];
    __shared__ uint32_t cursor[NB];
    __shared__ uint32_t rank_hist[kRankBins];
    __shared__ uint32_t warp_sums[kSortThreads / 32];

    const int row = blockIdx.x;
    const int tid = threadIdx.x;
    const int n1 = R + 1;
    const Fr* zrow = Z + (size_t)row * R;

    for (int b = tid; b < NB; b += kSortThreads) counts[b] = 0;
    if (tid < kRankBins) rank_hist[tid] = 0;
    __syncthreads();

    auto load_scalar = [&](int j, Fr& s) -> bool {
        if (j < R) s = load_fr(zrow + j);
        else if (blinds) s = load_fr(blinds + row);
        else return false;
        if (s.is_zero()) return false;
        if (scalars_are_mont) s = fp_from_mont(s);
        return true;
    };

    // pass 1: histogram
    for (int j = tid; j < n1; j += kSortThreads) {
        Fr s;
        if (!load_scalar(j, s)) continue;
        for_each_digit<C>(s, [&](int, uint32_t bucket, bool) { atomicAdd(&counts[bucket], 1u); });
    }
    __syncthreads();

    // exclusive scan of counts -> cursor, starts
    uint32_t local[PER];
    uint32_t sum = 0;
#pragma unroll
    for (int i = 0; i < PER; i++) {
        int b = tid * PER + i;
        uint32_t v = (b < NB) ? counts[b] : 0;
        local[i] = sum;
        sum += v;
    }
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((tid & 31) >= o) incl += t;
    }
    if ((tid & 31) == 31) warp_sums[tid >> 5] = incl;
    __syncthreads();
    uint32_t base = 0;
    for (int w = 0; w < (tid >> 5); w++) base += warp_sums[w];
    base += incl - sum;
    uint32_t* srow = starts + (size_t)row * (NB + 1);
#pragma unroll
    for (int i = 0; i < PER; i++) {
        int b = tid * PER + i;
        if (b < NB) {
            cursor[b] = base + local[i];
            srow[b] = base + local[i];
            atomicAdd(&rank_hist[min(counts[b], (uint32_t)(kRankBins - 1))], 1u);
        }
    }
    if (tid == kSortThreads - 1) srow[NB] = base + sum;
    __syncthreads();

    // bucket ranking by decreasing size (counting sort on the clamped size)
    uint32_t rank_base = 0;
    if (tid < kRankBins) {
        for (int v = tid + 1; v < kRankBins; v++) rank_base += rank_hist[v];
    }
    __syncthreads();
    if (tid < kRankBins) rank_hist[tid] = rank_base;
    __syncthreads();
    uint16_t* orow = order + (size_t)row * NB;
    for (int b = tid; b < NB; b += kSortThreads) {
        uint32_t pos = atomicAdd(&rank_hist[min(counts[b], (uint32_t)(kRankBins - 1))], 1u);
        orow[pos] = (uint16_t)b;
    }

    // pass 2: scatter
    uint32_t* erow = entries + (size_t)row * E;
    for (int j = tid; j < n1; j += kSortThreads) {
        Fr s;
        if (!load_scalar(j, s)) continue;
        for_each_digit<C>(s, [&](int k, uint32_t bucket, bool negative) {
            uint32_t pos = atomicAdd(&cursor[bucket], 1u);
            erow[pos] = (uint32_t)(k * n1 + j) | (negative ? 0x80000000u : 0u);
        });
    }
}

// ---------------------------------------------------------------------------------------------
// K3: bucket accumulation.  Thread gid -> (rank = gid / rows, row = gid % rows): a warp holds the
// same size-rank of 32 different rows (near-identical trip counts), and the grid walks ranks from
// the fullest buckets to the emptiest (longest-processing-time-first).
// ---------------------------------------------------------------------------------------------
static constexpr int kAccThreads = 128;

__global__ void __launch_bounds__(kAccThreads)
k_accumulate(const Affine* __restrict__ table, const uint32_t* __restrict__ entries,
             const uint32_t* __restrict__ starts, const uint16_t* __restrict__ order,
             XYZZ* __restrict__ buckets, int rows, int nb, uint32_t E) {
    const size_t gid = (size_t)blockIdx.x * kAccThreads + threadIdx.x;
    if (gid >= (size_t)rows * nb) return;
    const int rank = (int)(gid / rows);
    const int row = (int)(gid % rows);
    const int b = order[(size_t)row * nb + rank];
    const uint32_t* srow = starts + (size_t)row * (nb + 1);
    uint32_t i = srow[b];
    const uint32_t end = srow[b + 1];
    const uint32_t* erow = entries + (size_t)row * E;

    XYZZ acc = XYZZ::identity();
    for (; i < end; i++) {
        const uint32_t v = __ldg(erow + i);
        Affine p = load_affine(table + (v & 0x7fffffffu));
        if (p.is_identity()) continue;
        if (v >> 31) p.y = fp_neg(p.y);
        xyzz_add_mixed(acc, p);
    }
    store_xyzz(buckets + (size_t)row * nb + b, acc);
}

// ---------------------------------------------------------------------------------------------
// K4a: per-row bucket reduction  T = sum_b (b + 1) * S_b
// TPR threads per row, each owns m = nb / TPR consecutive buckets: running sums give
// tot = sum (b - lo + 1) S_b and run = sum S_b; the thread contributes tot + lo * run; contributions
// are combined by a shared-memory tree.
// ---------------------------------------------------------------------------------------------
static constexpr int kRedThreads = 128;

__device__ __forceinline__ XYZZ xyzz_small_mul(const XYZZ& p, uint32_t k) {
    XYZZ r = XYZZ::identity();
    if (k == 0) return r;
    for (int bit = 31 - __clz(k); bit >= 0; bit--) {
        r = xyzz_dbl(r);
        if ((k >> bit) & 1) xyzz_add(r, p);
    }
    return r;
}

__global__ void __launch_bounds__(kRedThreads)
k_reduce(const XYZZ* __restrict__ buckets, int rows, int nb, int tpr, XYZZ* __restrict__ row_totals) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    XYZZ* sm = reinterpret_cast<XYZZ*>(smem_raw);
    const int rows_per_block = kRedThreads / tpr;
    const int local_row = threadIdx.x / tpr;
    const int t = threadIdx.x % tpr;
    const int row = blockIdx.x * rows_per_block + local_row;
    const int m = nb / tpr;
    const bool active = row < rows;

    XYZZ tot = XYZZ::identity();
    if (active) {
        const XYZZ* brow = buckets + (size_t)row * nb;
        const int lo = t * m;
        XYZZ run = XYZZ::identity();
        for (int b = lo + m - 1; b >= lo; b--) {
            XYZZ s = load_xyzz(brow + b);
            xyzz_add(run, s);
            xyzz_add(tot, run);
        }
        if (lo) {
            XYZZ w = xyzz_small_mul(run, (uint32_t)lo);
            xyzz_add(tot, w);
        }
    }
    // tree over the tpr threads of each row
    for (int stride = tpr >> 1; stride >= 1; stride >>= 1) {
        __syncthreads();
        if (t >= stride && t < 2 * stride) sm[threadIdx.x] = tot;
        __syncthreads();
        if (t < stride) {
            XYZZ o = sm[threadIdx.x + stride];
            xyzz_add(tot, o);
        }
    }
    if (active && t == 0) store_xyzz(row_totals + row, tot);
}

// ---------------------------------------------------------------------------------------------
// K4b: XYZZ -> affine (+ identity flag)
// ---------------------------------------------------------------------------------------------
__global__ void k_normalize(const XYZZ* __restrict__ in, int n, Affine* __restrict__ out, uint8_t* __restrict__ inf) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    XYZZ p = load_xyzz(in + i);
    Affine a = xyzz_to_affine(p);
    store_affine(out + i, a);
    if (inf) inf[i] = p.is_identity() ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------
// batched scalar multiplication  out[i] = s[i] * P[i or 0]  (double-and-add, MSB first)
// ---------------------------------------------------------------------------------------------
__global__ void k_scalar_mul(const Affine* __restrict__ P, const uint8_t* __restrict__ Pinf, int p_stride,
                             const Fr* __restrict__ s, int s_stride, int n,
                             Affine* __restrict__ out, uint8_t* __restrict__ inf) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Affine p = load_affine(P + (size_t)i * p_stride);
    if (Pinf && Pinf[(size_t)i * p_stride]) p = Affine::identity();
    Fr k = fp_from_mont(load_fr(s + (size_t)i * s_stride));
    XYZZ acc = XYZZ::identity();
    if (!p.is_identity()) {
        bool started = false;
        for (int w = 7; w >= 0; w--) {
            uint32_t word = k.l[w];
            if (!started && word == 0) continue;
            for (int bit = 31; bit >= 0; bit--) {
                if (started) acc = xyzz_dbl(acc);
                if ((word >> bit) & 1) { xyzz_add_mixed(acc, p); started = true; }
            }
        }
    }
    Affine a = xyzz_to_affine(acc);
    store_affine(out + i, a);
    if (inf) inf[i] = acc.is_identity() ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------
// Fr form conversion
// ---------------------------------------------------------------------------------------------
__global__ void k_fr_convert(const Fr* __restrict__ in, int n, int to_mont, Fr* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr v = load_fr(in + i);
    v = to_mont ? fp_to_mont(v) : fp_from_mont(v);
    uint4* q = reinterpret_cast<uint4*>(out + i);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}

// ---------------------------------------------------------------------------------------------
// integer-pipe microbenchmarks (roofline denominator)
// ---------------------------------------------------------------------------------------------
template <int KIND>
__global__ void k_microbench(uint32_t* out, int iters, uint32_t seed) {
    uint32_t a[8], b = seed | 1u;
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 2654435761u + i + seed;
    if (KIND == 0) {
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int u = 0; u < 4; u++)
#pragma unroll
                for (int i = 0; i < 8; i++) asm volatile("mad.lo.u32 %0, %0, %1, %0;" : "+r"(a[i]) : "r"(b));
        }
    } else if (KIND == 1) {
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int u = 0; u < 4; u++)
#pragma unroll
                for (int i = 0; i < 8; i++) asm volatile("mad.hi.u32 %0, %0, %1, %0;" : "+r"(a[i]) : "r"(b));
        }
    } else if (KIND == 2) {
        uint64_t w[8];
#pragma unroll
        for (int i = 0; i < 8; i++) w[i] = ((uint64_t)a[i] << 32) | a[7 - i];
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int u = 0; u < 4; u++)
#pragma unroll
                for (int i = 0; i < 8; i++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(b));
        }
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] ^= (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32);
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) r ^= a[i];
    if (r == 0x12345678u) out[0] = r;   // never true in practice; keeps the chain alive
}

__global__ void k_microbench_fqmul(Fq* out, int iters, uint32_t seed) {
    Fq a, b;
#pragma unroll
    for (int i = 0; i < 8; i++) { a.l[i] = (threadIdx.x + 1) * 2654435761u + i; b.l[i] = seed + 77 * i; }
    a.l[7] &= 0x0fffffffu; b.l[7] &= 0x0fffffffu;
    Fq c = a, d = b;
    for (int it = 0; it < iters; it++) {
        a = fp_mul(a, b);
        c = fp_mul(c, d);
        b = fp_mul(b, a);
        d = fp_mul(d, c);
    }
    Fq r = fp_add(fp_add(a, b), fp_add(c, d));
    if (r.l[0] == 0x12345678u && r.l[1] == 0x9abcdef0u) out[0] = r;
}

}  // namespace sbn
