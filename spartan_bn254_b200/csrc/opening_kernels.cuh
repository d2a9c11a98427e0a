// Kernels of the Hyrax opening path (sm_100a): variable-base MSM over folded generators, the `bound`
// vector-matrix product, bullet-reduction folds and the R1CS-sat sumcheck round.
//
//   reference hyrax.rs:311-324          DensePolynomial::bound            -> k_bound_partial / k_fr_colsum
//   reference nizk/bullet.rs:57-59,70-76 Gamma, L, R                      -> k_msm_naive / k_points_sum / k_combine
//   reference nizk/bullet.rs:85-102     fold of G, a, b                   -> k_fold_points / k_fold_scalars
//   reference sumcheck.rs:501-530       cubic round evaluation            -> k_sumcheck_eval
//   reference hyrax.rs:195-203          bound_poly_var_top                -> k_bind_top
//
// These steps are round-sequential (Fiat-Shamir) and small (n <= 8192 per round), so they are written
// for low latency and small code (out-of-line field products), not for peak integer throughput.
#pragma once
#include "msm_kernels.cuh"

namespace sbn {

static constexpr int kSmallThreads = 64;

__device__ __forceinline__ void store_fr(Fr* p, const Fr& v) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
__device__ __noinline__ Fr fr_mul_call(const Fr& a, const Fr& b) { return fp_mul(a, b); }

// k * p for a canonical (non-Montgomery) scalar k, MSB-first double-and-add
__device__ __forceinline__ XYZZ xyzz_scalar_mul(const Affine& p, const Fr& k) {
    XYZZ acc = XYZZ::identity();
    if (p.is_identity()) return acc;
    bool started = false;
    for (int w = 7; w >= 0; w--) {
        const uint32_t word = k.l[w];
        if (!started && word == 0) continue;
        for (int bit = 31; bit >= 0; bit--) {
            if (started) acc = xyzz_dbl<MulCall>(acc);
            if ((word >> bit) & 1) { xyzz_add_mixed<MulCall>(acc, p); started = true; }
        }
    }
    return acc;
}

// block-wide sum of one XYZZ per thread (blockDim == kSmallThreads); result valid in thread 0
__device__ __forceinline__ XYZZ block_sum_xyzz(XYZZ v, XYZZ* sm) {
    for (int stride = kSmallThreads >> 1; stride >= 1; stride >>= 1) {
        sm[threadIdx.x] = v;
        __syncthreads();
        XYZZ o = (threadIdx.x < stride) ? sm[threadIdx.x + stride] : XYZZ::identity();
        __syncthreads();
        xyzz_add_call(&v, &o);
    }
    return v;
}

// partial[set * gridDim.x + block] = sum over the block's points of s_j * P_j.
// set = blockIdx.y selects (points + set * p_off, scalars + set * s_off): the L and R products of one
// bullet round run in a single launch.
__global__ void __launch_bounds__(kSmallThreads)
k_msm_naive(const Affine* __restrict__ points, const uint8_t* __restrict__ inf, long p_off,
            const Fr* __restrict__ scalars, long s_off, int n, int scalars_are_mont, XYZZ* __restrict__ partial) {
    __shared__ XYZZ sm[kSmallThreads];
    const int set = blockIdx.y;
    const int j = blockIdx.x * kSmallThreads + threadIdx.x;
    XYZZ acc = XYZZ::identity();
    if (j < n) {
        Affine p = load_affine(points + set * p_off + j);
        if (inf && inf[set * p_off + j]) p = Affine::identity();
        Fr k = load_fr(scalars + set * s_off + j);
        if (scalars_are_mont) k = fp_from_mont(k);
        acc = xyzz_scalar_mul(p, k);
    }
    acc = block_sum_xyzz(acc, sm);
    if (threadIdx.x == 0) store_xyzz(partial + (size_t)set * gridDim.x + blockIdx.x, acc);
}

// out[set * out_stride] = sum_{i < count} in[set * count + i]   (one block per set)
__global__ void __launch_bounds__(kSmallThreads)
k_points_sum(const XYZZ* __restrict__ in, int count, XYZZ* __restrict__ out, int out_stride) {
    __shared__ XYZZ sm[kSmallThreads];
    const XYZZ* src = in + (size_t)blockIdx.x * count;
    XYZZ acc = XYZZ::identity();
    for (int i = threadIdx.x; i < count; i += kSmallThreads) {
        XYZZ v = load_xyzz(src + i);
        xyzz_add_call(&acc, &v);
    }
    acc = block_sum_xyzz(acc, sm);
    if (threadIdx.x == 0) store_xyzz(out + (size_t)blockIdx.x * out_stride, acc);
}

// terms[(i / per) * group_stride + i % per] = s[i] * P[i] for a handful of (point, Montgomery scalar) pairs, one thread each, one warp
// per pair so the pairs run on different schedulers
__global__ void k_scalar_mul_terms(const Affine* __restrict__ P, const Fr* __restrict__ s, int n, XYZZ* __restrict__ terms,
                                   int group_stride, int per) {
    const int i = blockIdx.x;
    if (i >= n || threadIdx.x != 0) return;
    Affine p = load_affine(P + i);
    Fr k = fp_from_mont(load_fr(s + i));
    const int dst = per > 0 ? (i / per) * group_stride + i % per : i;
    store_xyzz(terms + dst, xyzz_scalar_mul(p, k));
}

// out[g] = affine( sum_{t < per_group} terms[g * per_group + t] )
__global__ void k_combine(const XYZZ* __restrict__ terms, int groups, int per_group, Affine* __restrict__ out,
                          uint8_t* __restrict__ inf) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= groups) return;
    XYZZ acc = XYZZ::identity();
    for (int t = 0; t < per_group; t++) {
        XYZZ v = load_xyzz(terms + (size_t)g * per_group + t);
        xyzz_add_call(&acc, &v);
    }
    Affine a = xyzz_to_affine<MulCall>(acc);
    store_affine(out + g, a);
    inf[g] = acc.is_identity() ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------
// Fr reductions
// ---------------------------------------------------------------------------------------------
// first: the first shuffle distance (16 for a full warp; lanes >= 2 * first must hold zero)
__device__ __forceinline__ Fr warp_sum_fr(Fr v, int first = 16) {
#pragma unroll 1
    for (int off = first; off >= 1; off >>= 1) {
        Fr o;
#pragma unroll
        for (int k = 0; k < 8; k++) o.l[k] = __shfl_down_sync(0xffffffffu, v.l[k], off);
        v = fp_add(v, o);
    }
    return v;      // lane 0 holds the sum
}

// Sum over the block; thread 0 holds it.  Shuffle trees inside the warps, one shared-memory hand-over, a shuffle tree over the
// warps' sums: two barriers instead of the fourteen of a shared-memory tree -- these reductions sit in round-sequential kernels
// (a sumcheck round evaluation does three of them), where every barrier is latency on the Fiat-Shamir critical path.
// threads: a multiple of 32, at most 1024; sm: at least threads / 32 elements; callers separate consecutive sums by a barrier.
__device__ __forceinline__ Fr block_sum_fr(Fr v, Fr* sm, int threads) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = threads >> 5;
    v = warp_sum_fr(v);
    if (nw == 1) return v;
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    if (warp != 0) return v;
    v = lane < nw ? sm[lane] : Fr::zero();
    int first = 1;
    while (2 * first < nw) first *= 2;
    return warp_sum_fr(v, first);
}

static constexpr int kDotThreads = 128;

// partial[set * gridDim.x + block] = sum_j a[set * a_off + j] * b[set * b_off + j]
__global__ void __launch_bounds__(kDotThreads)
k_fr_dot(const Fr* __restrict__ a, long a_off, const Fr* __restrict__ b, long b_off, int n, Fr* __restrict__ partial) {
    __shared__ Fr sm[kDotThreads];
    const int set = blockIdx.y;
    Fr acc = Fr::zero();
    for (int j = blockIdx.x * kDotThreads + threadIdx.x; j < n; j += gridDim.x * kDotThreads)
        acc = fp_add(acc, fr_mul_call(load_fr(a + set * a_off + j), load_fr(b + set * b_off + j)));
    acc = block_sum_fr(acc, sm, kDotThreads);
    if (threadIdx.x == 0) store_fr(partial + (size_t)set * gridDim.x + blockIdx.x, acc);
}

// out[set * out_stride] = sum_{i < count} in[set * count + i]
__global__ void __launch_bounds__(kDotThreads)
k_fr_sum(const Fr* __restrict__ in, int count, Fr* __restrict__ out, int out_stride) {
    __shared__ Fr sm[kDotThreads];
    Fr acc = Fr::zero();
    for (int i = threadIdx.x; i < count; i += kDotThreads) acc = fp_add(acc, load_fr(in + (size_t)blockIdx.x * count + i));
    acc = block_sum_fr(acc, sm, kDotThreads);
    if (threadIdx.x == 0) store_fr(out + (size_t)blockIdx.x * out_stride, acc);
}

// ---------------------------------------------------------------------------------------------
// bullet folds (nizk/bullet.rs:85-102), n2 = n / 2
// ---------------------------------------------------------------------------------------------
// a[i] <- u a[i] + u^-1 a[n2 + i];  b[i] <- u^-1 b[i] + u b[n2 + i]
__global__ void k_fold_scalars(Fr* __restrict__ a, Fr* __restrict__ b, int n2, const Fr* __restrict__ uu /* u, u_inv */) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n2) return;
    const Fr u = load_fr(uu), ui = load_fr(uu + 1);
    Fr na = fp_add(fr_mul_call(u, load_fr(a + i)), fr_mul_call(ui, load_fr(a + n2 + i)));
    Fr nb = fp_add(fr_mul_call(ui, load_fr(b + i)), fr_mul_call(u, load_fr(b + n2 + i)));
    store_fr(a + i, na);
    store_fr(b + i, nb);
}

// G[i] <- u^-1 G[i] + u G[n2 + i]  (joint double-and-add over the two canonical scalars, with G_L + G_R
// precomputed: one doubling and at most one addition per bit)
__global__ void __launch_bounds__(kSmallThreads)
k_fold_points(Affine* __restrict__ G, uint8_t* __restrict__ Ginf, int n2, const Fr* __restrict__ uu /* u, u_inv (Montgomery) */) {
    const int i = blockIdx.x * kSmallThreads + threadIdx.x;
    if (i >= n2) return;
    Affine gl = load_affine(G + i), gr = load_affine(G + n2 + i);
    if (Ginf[i]) gl = Affine::identity();
    if (Ginf[n2 + i]) gr = Affine::identity();
    const Fr ku = fp_from_mont(load_fr(uu));         // multiplies G_R
    const Fr kui = fp_from_mont(load_fr(uu + 1));    // multiplies G_L
    XYZZ both_x = XYZZ::from_affine(gl);
    if (!gr.is_identity()) xyzz_add_mixed<MulCall>(both_x, gr);
    const Affine both = xyzz_to_affine<MulCall>(both_x);
    XYZZ acc = XYZZ::identity();
    bool started = false;
    for (int w = 7; w >= 0; w--) {
        const uint32_t wl = kui.l[w], wr = ku.l[w];
        if (!started && (wl | wr) == 0) continue;
        for (int bit = 31; bit >= 0; bit--) {
            if (started) acc = xyzz_dbl<MulCall>(acc);
            const uint32_t sel = ((wl >> bit) & 1) | (((wr >> bit) & 1) << 1);
            if (sel) {
                const Affine& q = sel == 1 ? gl : (sel == 2 ? gr : both);
                if (!q.is_identity()) xyzz_add_mixed<MulCall>(acc, q);
                started = true;
            }
        }
    }
    const Affine r = xyzz_to_affine<MulCall>(acc);
    store_affine(G + i, r);
    Ginf[i] = acc.is_identity() ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------
// Table-based bullet rounds.  Folding the generators (bullet.rs:85-89, two full scalar multiplications per
// element and round) is avoided: after i rounds  G_i[j] = sum_t coef[t] * G[j + t*m]  (m = n / 2^i) with
// coef[t] = prod_r (bit_{i-1-r}(t) ? u_r : u_r^-1), so
//     MSM(a_L, G_R) = sum_{k : k mod m >= m/2} a[k mod m - m/2] * coef[k div m] * G[k]
//     MSM(a_R, G_L) = sum_{k : k mod m <  m/2} a[k mod m + m/2] * coef[k div m] * G[k]
// are two rows over the ORIGINAL, table-resident generators; c * Q = (c * q) * g1 and blind * H are two more
// columns of the same rows.  Only scalars are folded.
// ---------------------------------------------------------------------------------------------
// rows: 2 x (n + 1) Montgomery scalars; dots = (c_L, c_R)
__global__ void k_bullet_expand(const Fr* __restrict__ a, const Fr* __restrict__ coef, int n, int m,
                                const Fr* __restrict__ dots, const Fr* __restrict__ qs, Fr* __restrict__ rows) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > n) return;
    Fr l = Fr::zero(), r = Fr::zero();
    if (k == n) {
        const Fr q = load_fr(qs);
        l = fr_mul_call(load_fr(dots), q);
        r = fr_mul_call(load_fr(dots + 1), q);
    } else {
        const int h = m >> 1, jm = k % m;
        const Fr c = load_fr(coef + k / m);
        if (jm >= h) l = fr_mul_call(load_fr(a + jm - h), c);
        else r = fr_mul_call(load_fr(a + jm + h), c);
    }
    store_fr(rows + k, l);
    store_fr(rows + (size_t)(n + 1) + k, r);
}
// row[k < n] = vec[k]; row[n] = dot ? dot * qs : 0
__global__ void k_bullet_row_single(const Fr* __restrict__ vec, int n, const Fr* __restrict__ dot, const Fr* __restrict__ qs,
                                    Fr* __restrict__ row) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > n) return;
    Fr v = Fr::zero();
    if (k < n) v = load_fr(vec + k);
    else if (dot) v = fr_mul_call(load_fr(dot), load_fr(qs));
    store_fr(row + k, v);
}
// row[k < n] = vec[k] * scale; row[n] = 0
__global__ void k_bullet_row_scaled(const Fr* __restrict__ vec, int n, const Fr* __restrict__ scale, Fr* __restrict__ row) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > n) return;
    store_fr(row + k, k < n ? fr_mul_call(load_fr(vec + k), load_fr(scale)) : Fr::zero());
}
// out[2t] = in[t] * u^-1, out[2t + 1] = in[t] * u
__global__ void k_coef_update(const Fr* __restrict__ in, int len, const Fr* __restrict__ uu, Fr* __restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= len) return;
    const Fr c = load_fr(in + t);
    store_fr(out + 2 * t, fr_mul_call(c, load_fr(uu + 1)));
    store_fr(out + 2 * t + 1, fr_mul_call(c, load_fr(uu)));
}

// ---------------------------------------------------------------------------------------------
// bound (hyrax.rs:311-324): LZ[i] = sum_j L[j] Z[j * R + i].  grid = (R / 128, slices): each block sums a
// slice of the rows for 128 adjacent columns (coalesced 32 B loads); k_fr_colsum adds the slices.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_bound_partial(const Fr* __restrict__ Z, const Fr* __restrict__ Lv, int L_size, int R_size, int rows_per_slice,
                Fr* __restrict__ partial /* slices x R */) {
    const int col = blockIdx.x * 128 + threadIdx.x;
    if (col >= R_size) return;
    const int j0 = blockIdx.y * rows_per_slice;
    const int j1 = min(L_size, j0 + rows_per_slice);
    Fr acc = Fr::zero();
    for (int j = j0; j < j1; j++) acc = fp_add(acc, fp_mul(load_fr(Lv + j), load_fr(Z + (size_t)j * R_size + col)));
    store_fr(partial + (size_t)blockIdx.y * R_size + col, acc);
}
__global__ void k_fr_colsum(const Fr* __restrict__ partial, int slices, int R_size, Fr* __restrict__ out) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= R_size) return;
    Fr acc = Fr::zero();
    for (int s = 0; s < slices; s++) acc = fp_add(acc, load_fr(partial + (size_t)s * R_size + col));
    store_fr(out + col, acc);
}

// ---------------------------------------------------------------------------------------------
// R1CS-sat sumcheck round (sumcheck.rs:501-530): e_t = sum_i tau_t (Az_t Bz_t - Cz_t) at t = 0, 2, 3
// with x_t = lo + t (hi - lo); partial[(e * gridDim.x) + block]
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kDotThreads)
k_sumcheck_eval(const Fr* __restrict__ T0, const Fr* __restrict__ T1, const Fr* __restrict__ T2, const Fr* __restrict__ T3,
                int half, Fr* __restrict__ partial) {
    __shared__ Fr sm[kDotThreads];
    Fr e0 = Fr::zero(), e2 = Fr::zero(), e3 = Fr::zero();
    const Fr* T[4] = {T0, T1, T2, T3};
    for (int i = blockIdx.x * kDotThreads + threadIdx.x; i < half; i += gridDim.x * kDotThreads) {
        Fr v0[4], v2[4], v3[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const Fr lo = load_fr(T[k] + i), hi = load_fr(T[k] + half + i);
            v0[k] = lo;
            v2[k] = fp_sub(fp_add(hi, hi), lo);            // -lo + 2 hi
            v3[k] = fp_sub(fp_add(v2[k], hi), lo);         // -2 lo + 3 hi
        }
        e0 = fp_add(e0, fr_mul_call(v0[0], fp_sub(fr_mul_call(v0[1], v0[2]), v0[3])));
        e2 = fp_add(e2, fr_mul_call(v2[0], fp_sub(fr_mul_call(v2[1], v2[2]), v2[3])));
        e3 = fp_add(e3, fr_mul_call(v3[0], fp_sub(fr_mul_call(v3[1], v3[2]), v3[3])));
    }
    e0 = block_sum_fr(e0, sm, kDotThreads);
    if (threadIdx.x == 0) store_fr(partial + blockIdx.x, e0);
    __syncthreads();
    e2 = block_sum_fr(e2, sm, kDotThreads);
    if (threadIdx.x == 0) store_fr(partial + gridDim.x + blockIdx.x, e2);
    __syncthreads();
    e3 = block_sum_fr(e3, sm, kDotThreads);
    if (threadIdx.x == 0) store_fr(partial + 2 * gridDim.x + blockIdx.x, e3);
}

// quadratic variant (sumcheck.rs:690-699): e_t = sum_i z_t * ABC_t at t = 0, 2; partial[e * gridDim.x + block]
__global__ void __launch_bounds__(kDotThreads)
k_sumcheck_eval_quad(const Fr* __restrict__ T0, const Fr* __restrict__ T1, int half, Fr* __restrict__ partial) {
    __shared__ Fr sm[kDotThreads];
    Fr e0 = Fr::zero(), e2 = Fr::zero();
    for (int i = blockIdx.x * kDotThreads + threadIdx.x; i < half; i += gridDim.x * kDotThreads) {
        const Fr z0 = load_fr(T0 + i), z1 = load_fr(T0 + half + i), a0 = load_fr(T1 + i), a1 = load_fr(T1 + half + i);
        e0 = fp_add(e0, fr_mul_call(z0, a0));
        e2 = fp_add(e2, fr_mul_call(fp_sub(fp_add(z1, z1), z0), fp_sub(fp_add(a1, a1), a0)));
    }
    e0 = block_sum_fr(e0, sm, kDotThreads);
    if (threadIdx.x == 0) store_fr(partial + blockIdx.x, e0);
    __syncthreads();
    e2 = block_sum_fr(e2, sm, kDotThreads);
    if (threadIdx.x == 0) store_fr(partial + gridDim.x + blockIdx.x, e2);
}

// bound_poly_var_top on four tables at once: T[i] <- T[i] + r (T[half + i] - T[i])
__global__ void k_bind_top(Fr* __restrict__ T0, Fr* __restrict__ T1, Fr* __restrict__ T2, Fr* __restrict__ T3, int half,
                           const Fr* __restrict__ r) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= half) return;
    const Fr rr = load_fr(r);
    Fr* T[4] = {T0, T1, T2, T3};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (!T[k]) continue;                  // the quadratic variant binds two tables
        const Fr lo = load_fr(T[k] + i), hi = load_fr(T[k] + half + i);
        store_fr(T[k] + i, fp_add(lo, fr_mul_call(rr, fp_sub(hi, lo))));
    }
}

}  // namespace sbn
