// Fiat-Shamir on the device for the product-layer sumchecks (sm_100a).
//
// SumcheckInstanceProof::prove_cubic_batched (reference sumcheck.rs:165-330) is round-sequential: per round the batched
// evaluation (:201-271), the combination with the batching coefficients (:273-275), UniPoly::from_evals (unipoly.rs:28-59),
// the transcript append (unipoly.rs:119-127), the challenge (transcript.rs:56-67: 64 squeezed bytes reduced mod r) and the
// bind (:293-306).  A keyless-scale proof runs 441 such rounds and with the transcript on the host each one is a launch, a
// stream synchronisation, a microsecond of Keccak and another launch: ~60 us, of which the GPU computes for ~10.  Here the
// Merlin transcript (merlin 3.0: STROBE-128 over Keccak-f[1600]; same operations, byte for byte, as csrc/host/merlin.hpp)
// lives in device memory for the length of a layer:
//   k_bsc_round        between a round's evaluation kernel and its bind kernel: adds up the evaluation's per-block partial
//                      sums and runs the transcript step on one warp -- no host round trip
//   k_bsc_tail         ONE block runs every remaining round of a layer once its tables are short (<= 2 * kBscTailHalf
//                      entries): evaluation (a warp per instance), transcript (warp 0), bind (all threads), no launches
// Keccak-f runs on 25 lanes of a warp, one 64-bit lane of the state each: theta, rho-pi and chi are nine 64-bit shuffles per
// round.  The STROBE framing of a round (operation headers, labels, lengths, where the rate fills up and F runs) does not
// depend on the message bytes, and every round of a layer after the first starts at the same cursor (a challenge leaves
// pos = 64, pos_begin = 0): the host lays the framing out once per layer as a BscFrame (bsc_make_frame in sbn254.cu, a
// recording run of the same operations as csrc/host/merlin.hpp) and the device XORs the framed stream into the state chunk
// by chunk.  A single warp is a latency machine -- one Montgomery product is ~500 cycles -- so the field arithmetic of the
// step is arranged as eight products deep (products of the lanes' instances, the four coefficients' canonical forms and the
// two halves of the challenge in parallel lanes, Horner for the new claim).
#pragma once
#include "prodtree_kernels.cuh"

namespace sbn {
namespace dmerlin {

static constexpr int kRate = 166;
static constexpr uint8_t FLAG_I = 1, FLAG_A = 2, FLAG_C = 4, FLAG_M = 16, FLAG_K = 32;

__constant__ uint64_t kRC[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL, 0x000000000000808bULL,
    0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008aULL, 0x0000000000000088ULL,
    0x0000000080008009ULL, 0x000000008000000aULL, 0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL,
    0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
__constant__ uint8_t kRho[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};

__device__ __forceinline__ uint64_t shfl64(uint64_t v, int src) {
    const uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)v, src), hi = __shfl_sync(0xffffffffu, (uint32_t)(v >> 32), src);
    return ((uint64_t)hi << 32) | lo;
}

// Keccak-f[1600] over the 25 words at st64 (whole warp; lanes 25..31 shadow lanes 0..6 and their results are dropped)
__device__ __noinline__ void permute_warp(uint64_t* st64) {
    const int lane = threadIdx.x & 31;
    const int l = lane < 25 ? lane : lane - 25;
    const int x = l % 5, y = l / 5;
    uint64_t a = st64[l];
    const int rho = kRho[l];
    const int src_pi = ((x + 3 * y) % 5) + 5 * x;          // pi moves (xs, ys) to (ys, 2 xs + 3 ys): the source of (x, y)
    const int c1 = (x + 1) % 5 + 5 * y, c2 = (x + 2) % 5 + 5 * y;
    const int tm = (x + 4) % 5, tp = (x + 1) % 5;
#pragma unroll 1
    for (int round = 0; round < 24; round++) {
        const uint64_t p = a ^ shfl64(a, (l + 5) % 25) ^ shfl64(a, (l + 10) % 25) ^ shfl64(a, (l + 15) % 25) ^ shfl64(a, (l + 20) % 25);
        const uint64_t cm = shfl64(p, tm), cp = shfl64(p, tp);
        a ^= cm ^ ((cp << 1) | (cp >> 63));
        const uint64_t r = rho ? ((a << rho) | (a >> (64 - rho))) : a;
        const uint64_t b = shfl64(r, src_pi);
        const uint64_t b1 = shfl64(b, c1), b2 = shfl64(b, c2);
        a = b ^ (~b1 & b2);
        if (l == 0) a ^= kRC[round];
    }
    __syncwarp();
    if (lane < 25) st64[lane] = a;
    __syncwarp();
}


}  // namespace dmerlin

// Framing of one round's transcript traffic (see above).  stream: every byte the round absorbs, in order, with the four
// 32-byte coefficient slots zero; chunk c = stream[off, off + len) goes into the state at pos, and after it F runs with
// f_begin as the pos_begin byte (0xffff: no F -- only a trailing chunk can have that).  After the last F the 64 challenge
// bytes are squeezed from position 0, which leaves pos = 64, pos_begin = 0, cur_flags = I | A | C.
struct alignas(16) BscFrame {
    uint8_t stream[272];
    uint16_t coeff_off[4];
    uint16_t nchunks, total;
    uint16_t off[6], len[6], pos[6], f_begin[6];
};

// Device image of a layer's Fiat-Shamir state: the 203-byte Merlin state as csrc/host/merlin.hpp lays it out (200 state
// bytes, pos, pos_begin, cur_flags; padded to 208), the running claim, and the two framings (first round, later rounds).
struct BscState {
    uint8_t merlin[208];
    Fr claim;
    BscFrame frame[2];
};
struct BscConst { Fr two_inv, six_inv; };

struct BscShared {
    alignas(16) uint64_t st64[26];
    alignas(16) uint8_t stream[272];
    alignas(16) Fr canon[4];
    alignas(16) Fr poly[4];
    alignas(16) Fr comb[3];
    alignas(16) Fr r, claim;
    alignas(16) Fr prod[3 * 32];
};

// Plain (coherent, generic-address) load: the tail kernel re-reads tables it has just bound, and its evaluations sit in shared
// memory -- load_fr's ld.global.nc is for neither.
__device__ __forceinline__ Fr ld_fr(const Fr* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    const uint4 a = q[0], b = q[1];
    Fr r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}


// One round's transcript step, by one warp.  ev: the round's 3 n evaluations (e0, e2, e3 per instance; shared or global),
// sh.claim: the running claim (in / out).  The evaluations are combined with the batching coefficients (sumcheck.rs:273-275),
// interpolated to the cubic's coefficients (unipoly.rs:28-59), the coefficients enter the transcript (unipoly.rs:119-127), the
// challenge is squeezed (transcript.rs:56-67: 64 bytes, little-endian, reduced mod r) and the claim moves to poly(r).
// Leaves r in sh.r and the coefficients (lowest degree first) in sh.poly; lanes < 4 write them to poly_out.
__device__ __noinline__ void bsc_round_transcript(BscShared& sh, const BscFrame* __restrict__ fr, const Fr* ev, const Fr* __restrict__ cf,
                                                  int n, const BscConst& kc, Fr* __restrict__ poly_out) {
    const int lane = threadIdx.x & 31;
    uint8_t* st = reinterpret_cast<uint8_t*>(sh.st64);
    // the framed stream of this round, coefficient slots still zero (17 x 16 bytes)
    if (lane < 17) reinterpret_cast<uint4*>(sh.stream)[lane] = reinterpret_cast<const uint4*>(fr->stream)[lane];
    if (lane < n) {
        const Fr c = load_fr(cf + lane);
        const Fr a0 = ld_fr(ev + 3 * lane), a1 = ld_fr(ev + 3 * lane + 1), a2 = ld_fr(ev + 3 * lane + 2);
        sh.prod[3 * lane] = fp_mul(a0, c);
        sh.prod[3 * lane + 1] = fp_mul(a1, c);
        sh.prod[3 * lane + 2] = fp_mul(a2, c);
    }
    __syncwarp();
    if (lane < 3) {
        Fr acc = Fr::zero();
#pragma unroll 1
        for (int i = 0; i < n; i++) acc = fp_add(acc, sh.prod[3 * i + lane]);
        sh.comb[lane] = acc;
    }
    __syncwarp();
    if (lane == 0) {
        // evaluations at 0, 1, 2, 3 -> coefficients: a = (e3 - 3 e2 + 3 e1 - e0) / 6, b = (2 e0 - 5 e1 + 4 e2 - e3) / 2,
        // c = e1 - d - a - b, d = e0
        const Fr e0 = sh.comb[0], e1 = fp_sub(sh.claim, sh.comb[0]), e2 = sh.comb[1], e3 = sh.comb[2];
        const Fr e1x2 = fp_dbl(e1), e2x2 = fp_dbl(e2);
        const Fr e1x3 = fp_add(e1x2, e1), e2x3 = fp_add(e2x2, e2);
        const Fr a = fp_mul(kc.six_inv, fp_sub(fp_add(fp_sub(e3, e2x3), e1x3), e0));
        const Fr e1x5 = fp_add(fp_dbl(e1x2), e1), e2x4 = fp_dbl(e2x2);
        const Fr b = fp_mul(kc.two_inv, fp_sub(fp_add(fp_sub(fp_dbl(e0), e1x5), e2x4), e3));
        sh.poly[0] = e0;
        sh.poly[1] = fp_sub(fp_sub(fp_sub(e1, e0), a), b);
        sh.poly[2] = b;
        sh.poly[3] = a;
    }
    __syncwarp();
    if (lane < 4) {
        const Fr pk = sh.poly[lane];
        store_fr(poly_out + lane, pk);
        sh.canon[lane] = fp_from_mont(pk);
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 4; k++) sh.stream[fr->coeff_off[k] + lane] = reinterpret_cast<const uint8_t*>(&sh.canon[k])[lane];
    __syncwarp();
    const int nchunks = fr->nchunks;
#pragma unroll 1
    for (int c = 0; c < nchunks; c++) {
        const int off = fr->off[c], len = fr->len[c], pos = fr->pos[c], fb = fr->f_begin[c];
        for (int i = lane; i < len; i += 32) st[pos + i] ^= sh.stream[off + i];
        __syncwarp();
        if (fb != 0xffff) {
            if (lane == 0) {
                st[pos + len] ^= (uint8_t)fb;
                st[pos + len + 1] ^= 0x04;
                st[dmerlin::kRate + 1] ^= 0x80;
            }
            __syncwarp();
            dmerlin::permute_warp(sh.st64);
        }
    }
    // squeeze 64 bytes from position 0; lanes 0 / 1 take the low / high 256 bits: lo R + hi R^2 is the Montgomery form of
    // lo + hi 2^256 (montmul by R^2 and R^3)
    Fr v = Fr::zero();
    if (lane < 2) {
        Fr w, kk;
#pragma unroll
        for (int t = 0; t < 8; t++) {
            w.l[t] = reinterpret_cast<const uint32_t*>(st)[8 * lane + t];
            kk.l[t] = lane == 0 ? FrParams::R2(t) : FrParams::R3(t);
        }
        v = fp_mul(kk, w);
    }
    __syncwarp();
    if (lane < 8) sh.st64[lane] = 0;
    Fr hi;
#pragma unroll
    for (int t = 0; t < 8; t++) hi.l[t] = __shfl_sync(0xffffffffu, v.l[t], 1);
    if (lane == 0) {
        const Fr r = fp_add(v, hi);
        sh.r = r;
        Fr acc = fp_add(fp_mul(sh.poly[3], r), sh.poly[2]);        // Horner: ((a r + b) r + c) r + d
        acc = fp_add(fp_mul(acc, r), sh.poly[1]);
        sh.claim = fp_add(fp_mul(acc, r), sh.poly[0]);
        st[200] = 64;                                              // cursor after a challenge
        st[201] = 0;
        st[202] = 0x07;
    }
    __syncwarp();
}

__device__ __forceinline__ void bsc_load_state(BscShared& sh, const BscState* state) {
    const int lane = threadIdx.x & 31;
    if (lane < 26) sh.st64[lane] = reinterpret_cast<const uint64_t*>(state->merlin)[lane];
    if (lane == 0) sh.claim = ld_fr(&state->claim);
    __syncwarp();
}
__device__ __forceinline__ void bsc_store_state(BscShared& sh, BscState* state) {
    const int lane = threadIdx.x & 31;
    __syncwarp();
    if (lane < 26) reinterpret_cast<uint64_t*>(state->merlin)[lane] = sh.st64[lane];
    if (lane == 0) store_fr(&state->claim, sh.claim);
}

// Between the evaluation and the bind of round j.  partial: the evaluation kernel's per-block sums,
// partial[(3 i + e) nblk + b]; first: this is the layer's first round (framing 0).  rcur: where the bind reads the challenge.
__global__ void __launch_bounds__(128)
k_bsc_round(BscState* __restrict__ state, const Fr* __restrict__ partial, int nblk, const Fr* __restrict__ cf, int n, int first,
            BscConst kc, Fr* __restrict__ poly_out, Fr* __restrict__ r_out, Fr* __restrict__ rcur) {
    __shared__ BscShared sh;
    __shared__ __align__(16) Fr s_ev[3 * 32];
    if ((int)threadIdx.x < 3 * n) {
        Fr acc = Fr::zero();
        const Fr* p = partial + (size_t)threadIdx.x * nblk;
#pragma unroll 1
        for (int b = 0; b < nblk; b++) acc = fp_add(acc, ld_fr(p + b));
        s_ev[threadIdx.x] = acc;
    }
    __syncthreads();
    if (threadIdx.x >= 32) return;
    bsc_load_state(sh, state);
    bsc_round_transcript(sh, &state->frame[first ? 0 : 1], s_ev, cf, n, kc, poly_out);
    if (threadIdx.x == 0) {
        store_fr(r_out, sh.r);
        store_fr(rcur, sh.r);
    }
    bsc_store_state(sh, state);
}

// One block of 32 n threads: every remaining round of a layer whose tables have `len` <= 2 * kBscTailHalf entries left.
// Warp i evaluates instance i (one element per lane, shuffle tree), warp 0 runs the transcript step, all threads bind every
// table; the final values (entry 0 of every table) go to fin.  first: the layer's first round is among them.
static constexpr int kBscTailHalf = 128;
static constexpr int kBscTailMaxInst = 32;
template <int MAXT>      // 512 (n <= 16: up to 128 registers, nothing spills) or 1024
__global__ void __launch_bounds__(MAXT)
k_bsc_tail(BscState* __restrict__ state, const CubicTriple* __restrict__ triples, Fr* const* __restrict__ tables, int ntables, int n,
           int len, int first, const Fr* __restrict__ cf, BscConst kc, Fr* __restrict__ polys, Fr* __restrict__ r_out,
           Fr* __restrict__ fin) {
    __shared__ BscShared sh;
    __shared__ __align__(16) Fr s_ev[3 * kBscTailMaxInst];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp == 0) bsc_load_state(sh, state);
    const CubicTriple t = triples[warp];
    int round = 0;
    for (int half = len / 2; half >= 1; half >>= 1, round++) {
        {
            Fr e0 = Fr::zero(), e2 = Fr::zero(), e3 = Fr::zero();
#pragma unroll 1
            for (int i = lane; i < half; i += 32) {
                const Fr a0 = ld_fr(t.A + i), a1 = ld_fr(t.A + half + i);
                const Fr b0 = ld_fr(t.B + i), b1 = ld_fr(t.B + half + i);
                const Fr c0 = ld_fr(t.C + i), c1 = ld_fr(t.C + half + i);
                const Fr a2 = fp_sub(fp_add(a1, a1), a0), b2 = fp_sub(fp_add(b1, b1), b0), c2 = fp_sub(fp_add(c1, c1), c0);
                const Fr a3 = fp_sub(fp_add(a2, a1), a0), b3 = fp_sub(fp_add(b2, b1), b0), c3 = fp_sub(fp_add(c2, c1), c0);
                // three independent chains of two inlined products: a lone warp needs the overlap
                e0 = fp_add(e0, fp_mul(fp_mul(a0, b0), c0));
                e2 = fp_add(e2, fp_mul(fp_mul(a2, b2), c2));
                e3 = fp_add(e3, fp_mul(fp_mul(a3, b3), c3));
            }
            const int first_off = half >= 32 ? 16 : half / 2;       // lanes >= half hold zero: their tree levels are skipped
            e0 = warp_sum_fr(e0, first_off);
            e2 = warp_sum_fr(e2, first_off);
            e3 = warp_sum_fr(e3, first_off);
            if (lane == 0) { s_ev[3 * warp] = e0; s_ev[3 * warp + 1] = e2; s_ev[3 * warp + 2] = e3; }
        }
        __syncthreads();
        if (warp == 0) {
            bsc_round_transcript(sh, &state->frame[(first && round == 0) ? 0 : 1], s_ev, cf, n, kc, polys + 4 * round);
            if (lane == 0) store_fr(r_out + round, sh.r);
        }
        __syncthreads();
        const Fr r = sh.r;
        for (int w = threadIdx.x; w < ntables * half; w += blockDim.x) {
            Fr* T = tables[w / half];
            const int i = w % half;
            const Fr lo = ld_fr(T + i), hi = ld_fr(T + half + i);
            store_fr(T + i, fp_add(lo, fr_mul_call(r, fp_sub(hi, lo))));
        }
        __syncthreads();
    }
    if (warp == 0) bsc_store_state(sh, state);
    for (int k = threadIdx.x; k < ntables; k += blockDim.x) store_fr(fin + k, ld_fr(tables[k]));
}

}  // namespace sbn
