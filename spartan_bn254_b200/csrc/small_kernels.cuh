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

// One level of a shuffle tree, with BOTH lanes of a pair at work.  In the plain tree lane i < stride adds the value of lane
// i + stride while that lane idles, and for a lone warp a point addition is a dependency chain of 14 products (~9 us).  Here
// the two lanes exchange their values and split the products of add-2008-s between them -- 8 deep instead of 14:
//   both: U, S of "their" operand and one of ZZ1 ZZ2 / ZZZ1 ZZZ2 | exchange U, S | P^2 on one, R^2 on the other | exchange |
//   P^3 and ZZ1 ZZ2 P^2 on one, U1 P^2 on the other | exchange | S1 P^3, R (Q - X3) on one, ZZZ1 ZZZ2 P^3 on the other.
// The sum lands in the lower lane of the pair (lane & stride == 0); the upper lane's value is dead afterwards, as in the
// plain tree.  The identity, P + P and P - P are resolved after the fact from the operands (rare: table points of distinct
// generators), so the shuffles stay convergent.  All 32 lanes must call it.
__device__ __forceinline__ Fq shfl_xor_fq(const Fq& v, int mask) {
    Fq r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = __shfl_xor_sync(0xffffffffu, v.l[i], mask);
    return r;
}
__device__ __forceinline__ Fq sel_fq(bool c, const Fq& a, const Fq& b) {
    Fq r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = c ? a.l[i] : b.l[i];
    return r;
}
__device__ __noinline__ void xyzz_pair_level(XYZZ* accp, int stride) {
    XYZZ& acc = *accp;
    const bool hi = ((threadIdx.x & 31) & stride) != 0;
    XYZZ oth;
    oth.X = shfl_xor_fq(acc.X, stride); oth.Y = shfl_xor_fq(acc.Y, stride);
    oth.ZZ = shfl_xor_fq(acc.ZZ, stride); oth.ZZZ = shfl_xor_fq(acc.ZZZ, stride);
    // (p1, p2) = (lower lane's value, upper lane's value) on both lanes
    XYZZ p1, p2;
    p1.X = sel_fq(hi, oth.X, acc.X); p1.Y = sel_fq(hi, oth.Y, acc.Y); p1.ZZ = sel_fq(hi, oth.ZZ, acc.ZZ); p1.ZZZ = sel_fq(hi, oth.ZZZ, acc.ZZZ);
    p2.X = sel_fq(hi, acc.X, oth.X); p2.Y = sel_fq(hi, acc.Y, oth.Y); p2.ZZ = sel_fq(hi, acc.ZZ, oth.ZZ); p2.ZZZ = sel_fq(hi, acc.ZZZ, oth.ZZZ);
    // lower lane: U1 = X1 ZZ2, S1 = Y1 ZZZ2, A = ZZ1 ZZ2;  upper lane: U2 = X2 ZZ1, S2 = Y2 ZZZ1, B = ZZZ1 ZZZ2
    const Fq m1 = fp_mul(sel_fq(hi, p2.X, p1.X), sel_fq(hi, p1.ZZ, p2.ZZ));
    const Fq m2 = fp_mul(sel_fq(hi, p2.Y, p1.Y), sel_fq(hi, p1.ZZZ, p2.ZZZ));
    const Fq m3 = fp_mul(sel_fq(hi, p1.ZZZ, p1.ZZ), sel_fq(hi, p2.ZZZ, p2.ZZ));
    const Fq o1 = shfl_xor_fq(m1, stride), o2 = shfl_xor_fq(m2, stride);
    const Fq U1 = sel_fq(hi, o1, m1), U2 = sel_fq(hi, m1, o1), S1 = sel_fq(hi, o2, m2), S2 = sel_fq(hi, m2, o2);
    const Fq P = fp_sub(U2, U1), R = fp_sub(S2, S1);
    // lower: PP = P^2; upper: RR = R^2
    const Fq sq_in = sel_fq(hi, R, P);
    const Fq m4 = fp_mul(sq_in, sq_in);
    const Fq o4 = shfl_xor_fq(m4, stride);
    const Fq PP = sel_fq(hi, o4, m4), RR = sel_fq(hi, m4, o4);
    // lower: PPP = P PP, then ZZ3 = A PP; upper: Q = U1 PP
    const Fq m5 = fp_mul(sel_fq(hi, U1, P), PP);
    const Fq o5 = shfl_xor_fq(m5, stride);
    const Fq PPP = sel_fq(hi, o5, m5), Q = sel_fq(hi, m5, o5);
    const Fq X3 = fp_sub(fp_sub(RR, PPP), fp_dbl(Q));
    // lower: ZZ3 = A PP, Y3a = S1 PPP, Y3b = R (Q - X3); upper: ZZZ3 = B PPP (the other two slots idle)
    const Fq m6 = fp_mul(m3, sel_fq(hi, PPP, PP));              // lower: A PP = ZZ3; upper: B PPP = ZZZ3
    const Fq zzz3 = shfl_xor_fq(m6, stride);                    // the lower lane receives ZZZ3
    XYZZ res;
    res.X = X3;
    res.Y = fp_sub(fp_mul(R, fp_sub(Q, X3)), fp_mul(S1, PPP));
    res.ZZ = m6;
    res.ZZZ = zzz3;
    if (hi) return;                                             // the pair's sum lives in the lower lane
    if (p2.is_identity()) return;                               // acc (= p1) stays
    if (p1.is_identity()) { acc = p2; return; }
    if (P.is_zero()) {
        if (R.is_zero()) acc = xyzz_dbl<MulCall>(p1);
        else acc = XYZZ::identity();
        return;
    }
    acc = res;
}

// blockDim.x = 32 * (R + 1).  Z, blinds, out, inf, raw may be mapped host memory.  raw != nullptr: the row's sum is handed
// back as it is (XYZZ) and the caller normalises it -- the inversion is a 30 us dependency chain for a lone warp and a few
// microseconds for a host core.
__global__ void __launch_bounds__(32 * kSmallMaxCols)
k_small_commit(const Fr* __restrict__ Z, const Fr* __restrict__ blinds, int R, int n_cols, const Affine* __restrict__ table,
               Affine* __restrict__ out, uint8_t* __restrict__ inf, XYZZ* __restrict__ raw) {
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
    for (int stride = 16; stride >= 1; stride >>= 1) xyzz_pair_level(&acc, stride);
    if (lane == 0) part[j] = acc;
    __syncthreads();
    if (j != 0) return;
    XYZZ v = lane < nw ? part[lane] : XYZZ::identity();
    int top = 1;                                     // tree levels for nw values only: a level is several us of one warp's latency
    while (2 * top < nw) top *= 2;
    for (int stride = nw > 1 ? top : 0; stride >= 1; stride >>= 1) xyzz_pair_level(&v, stride);
    if (lane == 0) {
        if (raw) { store_xyzz(raw + row, v); return; }
        const Affine a = xyzz_to_affine<MulInline>(v);
        store_affine(out + row, a);
        inf[row] = v.is_identity() ? 1 : 0;
    }
}

// ---------------------------------------------------------------------------------------------
// Few rows over a LONG tabulated generator set (the opening's Cx, the two rows of every bullet round, delta).
// One or two rows through the bucket pipeline are a chain of ~110 dependent point additions (the longest bucket, then the
// two-level bucket reduction) that no amount of SMs shortens: ~1 ms per call, ~60 calls per proof.  With every digit
// multiple in HBM (2 GiB at 8192 generators -- this is what 180 GB are for) a row is a plain SUM of 32 table points per
// scalar, and sums parallelise: stage 1, thread = 4 windows of one scalar (3 mixed additions), shuffle tree over the warp,
// shared-memory tree over the block's 16 warps -> one partial per 64 scalars; stage 2, one block per row adds the
// partials.  Critical path ~20 additions; ~7 M products per row.
// ---------------------------------------------------------------------------------------------
// A shuffle-tree level costs a warp its full issue slots however few lanes carry a value, so the tree work of a block is
// 5 levels x 14 products x its warps.  With 4 windows per thread (round 1h: 512 threads, 16 warps a block) the trees were
// three quarters of the kernel's multiplier time: 230-360 us per two-row sum over 8193 generators, whatever the round of the
// bullet reduction (ncu, scripts/exp_bullet.py).  8 windows per thread: half the warps, the same 64 scalars a block, each
// thread a chain of 8 inlined mixed additions with the next table point in flight under the current one; two blocks per
// SM so that the 258 blocks of a two-row sum over 8193 generators are one wave.  (16 windows per thread and 4 warps a block
// is no faster: 86 k instructions per warp at two warps per scheduler is a latency chain of the same length.)  What is left
// of a bullet round (0.30 ms, was 0.39) is depth: 8 + 17 dependent point additions at 6-9 us each for a lone warp, a 32 us
// normalisation and six small launches.
static constexpr int kTabThreads = 256;
static constexpr int kTabWinPerThread = 8;                                   // 4 lanes per scalar
static constexpr int kTabScalarsPerBlock = kTabThreads * kTabWinPerThread / kSmallW;   // 64
static constexpr int kTabMaxRows = 4;

__device__ __noinline__ void xyzz_add_mixed_call(XYZZ* acc, const Affine* q) { xyzz_add_mixed<MulCall>(*acc, *q); }

// block tree: every warp's lane-0 value -> value of the whole block in thread 0 (blockDim.x = kTabThreads)
__device__ __forceinline__ XYZZ tab_block_sum(XYZZ acc, XYZZ* part /* shared, kTabThreads / 32 */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int stride = 16; stride >= 1; stride >>= 1) xyzz_pair_level(&acc, stride);
    if (lane == 0) part[warp] = acc;
    __syncthreads();
    if (warp != 0) return XYZZ::identity();
    XYZZ v = lane < kTabThreads / 32 ? part[lane] : XYZZ::identity();
    for (int stride = kTabThreads / 64; stride >= 1; stride >>= 1) xyzz_pair_level(&v, stride);
    return v;
}

// grid = (ceil((R + 1) / 64), rows).  partial[row * gridDim.x + blockIdx.x] = sum over the block's 64 scalars.
__global__ void __launch_bounds__(kTabThreads, 2)
k_tab_commit_partial(const Fr* __restrict__ Z, const Fr* __restrict__ blinds, int R, int n_cols, const Affine* __restrict__ table,
                     XYZZ* __restrict__ partial) {
    __shared__ XYZZ part[kTabThreads / 32];
    const int row = blockIdx.y;
    constexpr int lanes_per_scalar = kSmallW / kTabWinPerThread;              // 4
    const int sidx = blockIdx.x * kTabScalarsPerBlock + threadIdx.x / lanes_per_scalar;
    const int q = threadIdx.x % lanes_per_scalar;                             // windows 4q .. 4q + 3
    XYZZ acc = XYZZ::identity();
    if (sidx < R || (sidx == R && blinds)) {
        const Fr s = fp_from_mont(sidx < R ? load_fr(Z + (size_t)row * R + sidx) : load_fr(blinds + row));
        const int col = sidx < R ? sidx : n_cols - 1;
        int dig[kTabWinPerThread] = {};
        for_each_digit<kSmallC>(s, [&](int k, uint32_t dm1, bool negative) {
            if ((k / kTabWinPerThread) == q) dig[k % kTabWinPerThread] = negative ? -(int)(dm1 + 1) : (int)(dm1 + 1);
        });
        Affine cur = Affine::identity();
#pragma unroll 1
        for (int i = 0; i < kTabWinPerThread; i++) {
            if (dig[i] == 0) continue;
            const int k = q * kTabWinPerThread + i;
            const int d = dig[i] < 0 ? -dig[i] : dig[i];
            Affine nxt = load_affine(table + ((size_t)(k * n_cols + col) * kSmallD + (d - 1)));      // in flight under the addition below
            if (!cur.is_identity()) xyzz_add_mixed<MulInline>(acc, cur);
            if (dig[i] < 0 && !nxt.is_identity()) nxt = affine_neg(nxt);
            cur = nxt;
        }
        if (!cur.is_identity()) xyzz_add_mixed<MulInline>(acc, cur);
    }
    const XYZZ v = tab_block_sum(acc, part);
    if (threadIdx.x == 0) store_xyzz(partial + (size_t)row * gridDim.x + blockIdx.x, v);
}

// grid = rows: totals[row] = sum of the row's nblk partials
__global__ void __launch_bounds__(kTabThreads)
k_tab_commit_final(const XYZZ* __restrict__ partial, int nblk, XYZZ* __restrict__ totals) {
    __shared__ XYZZ part[kTabThreads / 32];
    const int row = blockIdx.x;
    XYZZ acc = XYZZ::identity();
    for (int i = threadIdx.x; i < nblk; i += kTabThreads) {
        const XYZZ v = load_xyzz(partial + (size_t)row * nblk + i);
        xyzz_add_call(&acc, &v);
    }
    const XYZZ v = tab_block_sum(acc, part);
    if (threadIdx.x == 0) store_xyzz(totals + row, v);
}

}  // namespace sbn
