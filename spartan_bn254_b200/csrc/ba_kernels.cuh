// Batched-affine pre-reduction of the bucket lists (sm_100a).
//
// k_accumulate pays 10 Fq products for every XYZZ mixed addition.  An AFFINE addition costs one inversion plus
// 3 products, and inversions can be shared: with Montgomery's trick a batch of n denominators costs one inversion and
// 3(n - 1) products.  The bucket-sorted entry list of a row is therefore first folded pairwise, in `rounds` passes:
//
//   round 1   pts1[g] = table[entries[2g]] + table[entries[2g+1]]
//   round k   ptsk[g] = pts(k-1)[2g] + pts(k-1)[2g+1]
//
// The sort aligns every bucket's first entry to 2^rounds slots and pads with NULL entries, so a pair never straddles two
// buckets and the rows of a chunk concatenate into one flat array per round.  After the last round bucket b owns
// ceil(n_b / 2^rounds) affine points, which k_accumulate_pts sums in XYZZ as before.
//
// One round = three launches over the flat pair array:
//   k_ba_prefix  every thread walks B pairs (interleaved over the block, so loads coalesce), forms the denominators
//                d_j (x2 - x1; 2*y1 for a doubling; 1 for a pair with nothing to add) and their running products,
//                which it stores; a shuffle scan gives every thread the product O_t of the OTHER lanes' totals and the
//                warp the product T of all of them.
//   k_ba_invert  one thread per warp of the previous launch: T^-1 (safegcd division steps, fp.cuh).
//   k_ba_finish  running = T^-1 * O_t = (own total)^-1; walking the pairs backwards yields each 1/d_j with two
//                products, then lambda, x3, y3 with three more: 6 products per addition in total, plus
//                (11 + 1) / B for the scans and ~70 / (32 B) for the inversion.
//
// P + P, P + (-P), identity operands and NULL padding are explicit cases (the reference's generators repeat the same
// point -- group.rs:110-132 -- so doublings are common, not exceptional).
#pragma once
#include "msm_kernels.cuh"

namespace sbn {

static constexpr uint32_t kNullEntry = 0xffffffffu;
static constexpr int kBaThreads = 256;

enum BaKind : int { BA_NONE = 0, BA_ADD = 1, BA_DBL = 2 };

__device__ __forceinline__ Fq fq_one() { return Fq::one(); }

__device__ __forceinline__ Fq shfl_fq(const Fq& v, int src) {
    Fq r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = __shfl_sync(0xffffffffu, v.l[i], src);
    return r;
}

// Operands of pair g.  FIRST: entries -> table points (sign bit = negate y); otherwise the previous round's points.
template <bool FIRST>
__device__ __forceinline__ void ba_load_pair(const uint32_t* __restrict__ entries, const Affine* __restrict__ table,
                                             const Affine* __restrict__ in, size_t g, Affine& P, Affine& Q) {
    if (FIRST) {
        const uint2 e = __ldg(reinterpret_cast<const uint2*>(entries) + g);
        P = e.x == kNullEntry ? Affine::identity() : load_affine(table + (e.x & 0x7fffffffu));
        Q = e.y == kNullEntry ? Affine::identity() : load_affine(table + (e.y & 0x7fffffffu));
        if (e.x != kNullEntry && (e.x >> 31) && !P.is_identity()) P.y = fp_neg(P.y);
        if (e.y != kNullEntry && (e.y >> 31) && !Q.is_identity()) Q.y = fp_neg(Q.y);
    } else {
        P = load_affine(in + 2 * g);
        Q = load_affine(in + 2 * g + 1);
    }
}

// Denominator of the pair's addition and what kind of addition it is.
__device__ __forceinline__ int ba_classify(const Affine& P, const Affine& Q, Fq& d) {
    if (P.is_identity() || Q.is_identity()) return BA_NONE;
    if (P.x != Q.x) { d = fp_sub(Q.x, P.x); return BA_ADD; }
    if (P.y == Q.y) { d = fp_dbl(P.y); return BA_DBL; }      // y != 0 on a prime-order curve
    return BA_NONE;                                           // P + (-P)
}

// Denominator of pair g from the x coordinates alone; y is fetched only when the x coordinates coincide (doubling or
// cancellation) or are zero (the (0, 0) encoding of the identity).  Returns false when the pair needs no inversion.
template <bool FIRST>
__device__ __forceinline__ bool ba_denominator(const uint32_t* __restrict__ entries, const Affine* __restrict__ table,
                                               const Affine* __restrict__ in, size_t g, Fq& d) {
    const Affine *pp, *qp;
    bool pneg = false, qneg = false;
    if (FIRST) {
        const uint2 e = __ldg(reinterpret_cast<const uint2*>(entries) + g);
        if (e.x == kNullEntry || e.y == kNullEntry) return false;
        pp = table + (e.x & 0x7fffffffu);
        qp = table + (e.y & 0x7fffffffu);
        pneg = e.x >> 31;
        qneg = e.y >> 31;
    } else {
        pp = in + 2 * g;
        qp = in + 2 * g + 1;
    }
    const Fq px = load_fq(&pp->x), qx = load_fq(&qp->x);
    if (px != qx && !px.is_zero() && !qx.is_zero()) { d = fp_sub(qx, px); return true; }
    // rare path: full classification
    Affine P, Q;
    P.x = px; P.y = load_fq(&pp->y);
    Q.x = qx; Q.y = load_fq(&qp->y);
    if (pneg && !P.is_identity()) P.y = fp_neg(P.y);
    if (qneg && !Q.is_identity()) Q.y = fp_neg(Q.y);
    return ba_classify(P, Q, d) != BA_NONE;
}

template <bool FIRST>
__global__ void __launch_bounds__(kBaThreads)
k_ba_prefix(const uint32_t* __restrict__ entries, const Affine* __restrict__ table, const Affine* __restrict__ in,
            size_t npairs, int B, Fq* __restrict__ prefix, Fq* __restrict__ other, Fq* __restrict__ warp_tot) {
    const size_t base = (size_t)blockIdx.x * kBaThreads * B;
    Fq run = fq_one();
#pragma unroll 1
    for (int j = 0; j < B; j++) {
        const size_t g = base + (size_t)j * kBaThreads + threadIdx.x;
        if (g < npairs) {
            Fq d;
            if (ba_denominator<FIRST>(entries, table, in, g, d)) run = fp_mul(run, d);
            store_fq(prefix + g, run);
        }
    }
    // product of the other lanes' totals, and of all of them
    const int lane = threadIdx.x & 31;
    Fq pre = run, suf = run;
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
    const size_t tid_global = (size_t)blockIdx.x * kBaThreads + threadIdx.x;
    store_fq(other + tid_global, fp_mul(pe, se));
    if (lane == 31) store_fq(warp_tot + (tid_global >> 5), pre);
}

__global__ void __launch_bounds__(64)
k_ba_invert(const Fq* __restrict__ warp_tot, size_t n, Fq* __restrict__ inv) {
    const size_t i = (size_t)blockIdx.x * 64 + threadIdx.x;
    if (i >= n) return;
    store_fq(inv + i, fp_inv_fast(load_fq(warp_tot + i)));
}

// L2 prefetch of the two table points the next pair of this thread will gather (round 1 only): the gathers are random 64 B
// reads over a table of tens of GB, and a warp that waits for them holds 80 registers per thread doing nothing.
__device__ __forceinline__ void ba_prefetch_pair(const uint32_t* __restrict__ entries, const Affine* __restrict__ table, size_t g) {
    const uint2 e = __ldg(reinterpret_cast<const uint2*>(entries) + g);
    if (e.x != kNullEntry) asm volatile("prefetch.global.L2 [%0];" ::"l"(table + (e.x & 0x7fffffffu)));
    if (e.y != kNullEntry) asm volatile("prefetch.global.L2 [%0];" ::"l"(table + (e.y & 0x7fffffffu)));
}

// MINB = resident CTAs per SM the register allocation aims at (3: 80 registers, 4: 64 with a few spilled words);
// PF = prefetch the next pair's table points (round 1).
template <bool FIRST, int MINB = 3, bool PF = false>
__global__ void __launch_bounds__(kBaThreads, MINB)
k_ba_finish(const uint32_t* __restrict__ entries, const Affine* __restrict__ table, const Affine* __restrict__ in,
            size_t npairs, int B, const Fq* __restrict__ prefix, const Fq* __restrict__ other,
            const Fq* __restrict__ warp_inv, Affine* __restrict__ out) {
    const size_t base = (size_t)blockIdx.x * kBaThreads * B;
    const size_t tid_global = (size_t)blockIdx.x * kBaThreads + threadIdx.x;
    if (FIRST && PF) {
        const size_t g = base + (size_t)(B - 1) * kBaThreads + threadIdx.x;
        if (g < npairs) ba_prefetch_pair(entries, table, g);
    }
    Fq run = fp_mul(load_fq(warp_inv + (tid_global >> 5)), load_fq(other + tid_global));   // (own total)^-1
#pragma unroll 1
    for (int j = B - 1; j >= 0; j--) {
        const size_t g = base + (size_t)j * kBaThreads + threadIdx.x;
        if (FIRST && PF) {
            if (j > 0 && g - kBaThreads < npairs) ba_prefetch_pair(entries, table, g - kBaThreads);
        }
        if (g >= npairs) continue;
        Affine P, Q;
        ba_load_pair<FIRST>(entries, table, in, g, P, Q);
        Fq d;
        const int kind = ba_classify(P, Q, d);
        Affine R;
        if (kind == BA_NONE) {
            // identity + X = X;  P + (-P) = identity
            if (P.is_identity()) R = Q;
            else if (Q.is_identity()) R = P;
            else R = Affine::identity();
        } else {
            Fq inv_d = run;
            if (j > 0) inv_d = fp_mul(run, load_fq(prefix + (g - kBaThreads)));    // running product before this pair
            run = fp_mul(run, d);
            Fq num;
            if (kind == BA_ADD) {
                num = fp_sub(Q.y, P.y);
            } else {
                const Fq xx = fp_mul(P.x, P.x);
                num = fp_add(fp_dbl(xx), xx);
            }
            const Fq lambda = fp_mul(num, inv_d);
            Fq x3 = fp_sub(fp_sub(fp_mul(lambda, lambda), P.x), Q.x);
            R.x = x3;
            R.y = fp_sub(fp_mul(lambda, fp_sub(P.x, x3)), P.y);
        }
        store_affine(out + g, R);
    }
}

// K3 on pre-reduced lists: a task is a run of at most `cap` affine points of one bucket.
__global__ void __launch_bounds__(kAccThreads)
k_accumulate_pts(const Affine* __restrict__ pts, size_t row_stride, const uint32_t* __restrict__ tstart,
                 const Task* __restrict__ tasks, XYZZ* __restrict__ partials, int rows, int nb, uint32_t max_tasks) {
    const size_t gid = (size_t)blockIdx.x * kAccThreads + threadIdx.x;
    if (gid >= (size_t)rows * max_tasks) return;
    const uint32_t rank = (uint32_t)(gid / rows);
    const int row = (int)(gid % rows);
    if (rank >= tstart[(size_t)row * (nb + 1) + nb]) return;
    const Task t = tasks[(size_t)row * max_tasks + rank];
    const Affine* e = pts + (size_t)row * row_stride + t.start;
    const uint32_t len = t.len_slot >> 24;
    XYZZ acc = XYZZ::identity();
    for (uint32_t i = 0; i < len; i++) {
        const Affine p = load_affine(e + i);
        if (p.is_identity()) continue;
        xyzz_add_mixed(acc, p);
    }
    store_xyzz(partials + (size_t)row * max_tasks + (t.len_slot & 0xffffffu), acc);
}

}  // namespace sbn
