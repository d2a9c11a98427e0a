// BN254 prime-field arithmetic for sm_100a: 8 x 32-bit limbs, Montgomery form (R = 2^256).
//
// The multiply is a word-serial (CIOS) Montgomery product built from 32x32->64 multiply-adds
// (`mad.lo.cc.u32` / `madc.hi.cc.u32` pairs, which ptxas fuses into IMAD.WIDE.U32[.X] carry chains).
// Two interleaved accumulators hold the even- and odd-limb partial products so that every 64-bit
// product lands on an aligned limb pair and each accumulator is a single unbroken carry chain.
//
// Layout in memory matches ark-ff's Fp256<MontBackend> (4 x u64 little-endian limbs == 8 x u32 LE
// limbs), i.e. what `Scalar(pub Fr)` (reference scalar.rs:15) and G1Affine coordinates hold.
//
// Every function is __host__ __device__: on the host the PTX carry-flag primitives are emulated with
// a thread-local carry bit so the exact same algorithm is unit-tested on CPU (tests/host_fp_test.cu)
// before it ever reaches a GPU.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SBN_HD __host__ __device__ __forceinline__
#define SBN_D __device__ __forceinline__
#else
#define SBN_HD inline
#define SBN_D inline
#endif

namespace sbn {

// ------------------------------------------------------------------------------------------------
// carry-flag primitives
// ------------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
SBN_D uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
SBN_D uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
SBN_D uint32_t addc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
SBN_D uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
SBN_D uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
SBN_D uint32_t subc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
SBN_D uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
SBN_D uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
SBN_D uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
SBN_D uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
SBN_D uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
SBN_D uint32_t mul_hi(uint32_t a, uint32_t b) { return __umulhi(a, b); }
#else
// host emulation of the PTX condition-code register
inline uint32_t& cc_flag() { static thread_local uint32_t cc = 0; return cc; }
inline uint32_t add_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b; cc_flag() = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t addc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b + cc_flag(); cc_flag() = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t addc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b + cc_flag(); return (uint32_t)t; }
inline uint32_t sub_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b; cc_flag() = (uint32_t)(t >> 32) & 1; return (uint32_t)t; }
inline uint32_t subc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b - cc_flag(); cc_flag() = (uint32_t)(t >> 32) & 1; return (uint32_t)t; }
inline uint32_t subc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b - cc_flag(); return (uint32_t)t; }
inline uint32_t mul_lo(uint32_t a, uint32_t b) { return (uint32_t)((uint64_t)a * b); }
inline uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return add_cc(mul_lo(a, b), c); }
inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return addc_cc(mul_lo(a, b), c); }
inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return addc_cc(mul_hi(a, b), c); }
inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return addc(mul_hi(a, b), c); }
#endif
// NOTE: PTX sub.cc sets CC.CF to the *borrow*; subc consumes it as a borrow -- the emulation matches.

// ------------------------------------------------------------------------------------------------
// field parameters (as constexpr functions so unrolled loops fold them into IMAD immediates)
// ------------------------------------------------------------------------------------------------
struct FqParams {   // base field of BN254 (coordinates of G1)
    static SBN_HD constexpr uint32_t P(int i) {
        constexpr uint32_t v[8] = {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
        return v[i];
    }
    static SBN_HD constexpr uint32_t R1(int i) {  // 2^256 mod p
        constexpr uint32_t v[8] = {0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u, 0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
        return v[i];
    }
    static SBN_HD constexpr uint32_t R2(int i) {  // 2^512 mod p
        constexpr uint32_t v[8] = {0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u, 0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u};
        return v[i];
    }
    static constexpr uint32_t INV = 0xe4866389u;  // -p^-1 mod 2^32
    static SBN_HD constexpr int32_t P30(int i) {  // p in nine 30-bit limbs (fp_inv_plain)
        constexpr int32_t v[9] = {0x187cfd47, 0x3082305b, 0x71ca8d3, 0x205aa45a, 0x1585d97, 0x116da06, 0x1a029b85, 0x139cb84c, 0x3064};
        return v[i];
    }
    static constexpr uint32_t INV30 = 0x1b799c77u;  // p^-1 mod 2^30
    static SBN_HD constexpr uint32_t R3(int i) {  // 2^768 mod p
        constexpr uint32_t v[8] = {0xda1530dfu, 0xb1cd6dafu, 0xa7283db6u, 0x62f210e6u, 0x0ada0afbu, 0xef7f0b0cu, 0x2d592544u, 0x20fd6e90u};
        return v[i];
    }
};
struct FrParams {   // scalar field of BN254
    static SBN_HD constexpr uint32_t P(int i) {
        constexpr uint32_t v[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
        return v[i];
    }
    static SBN_HD constexpr uint32_t R1(int i) {
        constexpr uint32_t v[8] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u, 0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
        return v[i];
    }
    static SBN_HD constexpr uint32_t R2(int i) {
        constexpr uint32_t v[8] = {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u, 0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u};
        return v[i];
    }
    static constexpr uint32_t INV = 0xefffffffu;
    static SBN_HD constexpr int32_t P30(int i) {
        constexpr int32_t v[9] = {0x30000001, 0xf87d64f, 0x1b970914, 0xcfa121e, 0x1585d28, 0x116da06, 0x1a029b85, 0x139cb84c, 0x3064};
        return v[i];
    }
    static constexpr uint32_t INV30 = 0x10000001u;
    static SBN_HD constexpr uint32_t R3(int i) {
        constexpr uint32_t v[8] = {0xb4bf0040u, 0x5e94d8e1u, 0x1cfbb6b8u, 0x2a489cbeu, 0xa19fcfedu, 0x893cc664u, 0x7fcc657cu, 0x0cf8594bu};
        return v[i];
    }
};

// ------------------------------------------------------------------------------------------------
// field element
// ------------------------------------------------------------------------------------------------
template <class F>
struct Fp {
    uint32_t l[8];

    static SBN_HD Fp zero() { Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.l[i] = 0;
        return r; }
    static SBN_HD Fp one() { Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.l[i] = F::R1(i);
        return r; }
    SBN_HD bool is_zero() const {
        uint32_t t = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) t |= l[i];
        return t == 0;
    }
    SBN_HD bool operator==(const Fp& o) const {
        uint32_t t = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) t |= l[i] ^ o.l[i];
        return t == 0;
    }
    SBN_HD bool operator!=(const Fp& o) const { return !(*this == o); }
};

// r = a - p if a >= p else a   (a < 2p)
template <class F>
SBN_HD void fp_reduce_once(Fp<F>& a) {
    uint32_t t[8];
    t[0] = sub_cc(a.l[0], F::P(0));
#pragma unroll
    for (int i = 1; i < 8; i++) t[i] = subc_cc(a.l[i], F::P(i));
    uint32_t borrow = subc(0u, 0u);  // 0 - 0 - CF -> 0xffffffff when a < p
#pragma unroll
    for (int i = 0; i < 8; i++) a.l[i] = borrow ? a.l[i] : t[i];
}

template <class F>
SBN_HD Fp<F> fp_add(const Fp<F>& a, const Fp<F>& b) {
    Fp<F> r;
    r.l[0] = add_cc(a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < 7; i++) r.l[i] = addc_cc(a.l[i], b.l[i]);
    r.l[7] = addc(a.l[7], b.l[7]);   // a,b < p < 2^254: no carry out of limb 7
    fp_reduce_once(r);
    return r;
}

template <class F>
SBN_HD Fp<F> fp_sub(const Fp<F>& a, const Fp<F>& b) {
    Fp<F> r;
    r.l[0] = sub_cc(a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) r.l[i] = subc_cc(a.l[i], b.l[i]);
    uint32_t borrow = subc(0u, 0u);   // all-ones when a < b
    r.l[0] = add_cc(r.l[0], F::P(0) & borrow);
#pragma unroll
    for (int i = 1; i < 7; i++) r.l[i] = addc_cc(r.l[i], F::P(i) & borrow);
    r.l[7] = addc(r.l[7], F::P(7) & borrow);
    return r;
}

template <class F>
SBN_HD Fp<F> fp_neg(const Fp<F>& a) {
    return fp_sub(Fp<F>::zero(), a);
}

template <class F>
SBN_HD Fp<F> fp_dbl(const Fp<F>& a) { return fp_add(a, a); }

// ------------------------------------------------------------------------------------------------
// Montgomery multiplication.  Requires a < 2p; b may be any 256-bit value.  Returns a*b/R mod p, < p.
//
// Invariant between word iterations: T = E + O * 2^32 with E = sum e[k] 2^(32k), O = sum o[k] 2^(32k).
// Products of even-indexed limbs of `a` (and p) go to E, odd-indexed ones to O, so each 64-bit
// product is added to an aligned (lo, hi) limb pair.  After the reduction step e[0] == 0 and the
// division by 2^32 is a role swap: E' = O + e[1],  O' = E >> 64.
// Bound: T < a + p <= 3p < 2^256 after every iteration, so neither accumulator overflows.
// ------------------------------------------------------------------------------------------------
#if !defined(__CUDA_ARCH__) && defined(SBN_HOST_FAST_FP)
// Host-side product for the library's own host code (transcript challenges, UniPoly arithmetic of the in-library sumcheck
// loops, point compression): the same CIOS on 4 x 64-bit limbs with unsigned __int128.  The 32-bit form below, with the PTX
// carry flag emulated through a thread-local, is what the host TESTS compile (they exist to exercise the DEVICE algorithm);
// it costs ~1 us per product, which 441 sumcheck rounds and 4096 compressed points per proof turn into milliseconds.
template <class F>
inline Fp<F> fp_mul_host64(const Fp<F>& a, const Fp<F>& b) {
    typedef unsigned __int128 u128;
    uint64_t A[4], B[4], P[4], t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        A[i] = (uint64_t)a.l[2 * i] | ((uint64_t)a.l[2 * i + 1] << 32);
        B[i] = (uint64_t)b.l[2 * i] | ((uint64_t)b.l[2 * i + 1] << 32);
        P[i] = (uint64_t)F::P(2 * i) | ((uint64_t)F::P(2 * i + 1) << 32);
    }
    // -p^-1 mod 2^64 from the 32-bit constant: one Newton step (x <- x (2 + p x) for the negated inverse)
    uint64_t inv = (uint64_t)F::INV;
    inv = inv * (2 + P[0] * inv);
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) {
            c += (u128)A[j] * B[i] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (uint64_t)c;
        t[5] = (uint64_t)(c >> 64);
        const uint64_t m = t[0] * inv;
        c = (u128)m * P[0] + t[0];
        c >>= 64;
        for (int j = 1; j < 4; j++) {
            c += (u128)m * P[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
    }
    // one conditional subtraction: the result is below 2p for a < 2p (b any 256-bit value), as in the 32-bit form
    uint64_t r[4], borrow = 0;
    for (int i = 0; i < 4; i++) {
        const u128 d = (u128)t[i] - P[i] - borrow;
        r[i] = (uint64_t)d;
        borrow = (uint64_t)(d >> 64) & 1;
    }
    const bool ge = t[4] != 0 || borrow == 0;
    Fp<F> out;
    for (int i = 0; i < 4; i++) {
        const uint64_t v = ge ? r[i] : t[i];
        out.l[2 * i] = (uint32_t)v;
        out.l[2 * i + 1] = (uint32_t)(v >> 32);
    }
    return out;
}
#endif

template <class F>
SBN_HD Fp<F> fp_mul(const Fp<F>& a, const Fp<F>& b) {
#if !defined(__CUDA_ARCH__) && defined(SBN_HOST_FAST_FP)
    return fp_mul_host64(a, b);
#else
    // Both accumulators start at zero and every word iteration (including the first) has the same
    // shape: with a special-cased first iteration ptxas splits the a*b_i products into
    // IMAD + IMAD.HI + IADD3 instead of IMAD.WIDE.U32.X (IMAD.HI is half rate on sm_100).
    uint32_t e[8], o[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { e[i] = 0; o[i] = 0; }

#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t bi = b.l[i];
        uint32_t* x = (i & 1) ? o : e;   // even-aligned accumulator of this iteration (old O)
        uint32_t* y = (i & 1) ? e : o;   // odd-aligned accumulator (old E, shifted down by 64 bits)

        x[0] = add_cc(x[0], y[1]);                       // E' = O + e[1]; carry joins O' at weight 2^32
#pragma unroll
        for (int j = 0; j < 6; j += 2) {                 // O' = (E >> 64) + a_odd * bi + carry
            y[j] = madc_lo_cc(a.l[j + 1], bi, y[j + 2]);
            y[j + 1] = madc_hi_cc(a.l[j + 1], bi, y[j + 3]);
        }
        y[6] = madc_lo_cc(a.l[7], bi, 0u);
        y[7] = madc_hi_cc(a.l[7], bi, 0u);               // (no carry out: T < 3p; .cc form lets ptxas fuse the pair)

        x[0] = mad_lo_cc(a.l[0], bi, x[0]);              // E' += a_even * bi
        x[1] = madc_hi_cc(a.l[0], bi, x[1]);
#pragma unroll
        for (int j = 2; j < 8; j += 2) {
            x[j] = madc_lo_cc(a.l[j], bi, x[j]);
            x[j + 1] = madc_hi_cc(a.l[j], bi, x[j + 1]);
        }
        y[7] = addc(y[7], 0u);

        const uint32_t m = mul_lo(x[0], F::INV);
        y[0] = mad_lo_cc(F::P(1), m, y[0]);              // O' += p_odd * m
        y[1] = madc_hi_cc(F::P(1), m, y[1]);
#pragma unroll
        for (int j = 2; j < 8; j += 2) {
            y[j] = madc_lo_cc(F::P(j + 1), m, y[j]);
            y[j + 1] = madc_hi_cc(F::P(j + 1), m, y[j + 1]);
        }
        x[0] = mad_lo_cc(F::P(0), m, x[0]);              // E' += p_even * m  (x[0] becomes 0)
        x[1] = madc_hi_cc(F::P(0), m, x[1]);
#pragma unroll
        for (int j = 2; j < 8; j += 2) {
            x[j] = madc_lo_cc(F::P(j), m, x[j]);
            x[j + 1] = madc_hi_cc(F::P(j), m, x[j + 1]);
        }
        y[7] = addc(y[7], 0u);
    }

    // after i = 7 (odd): E lives in `o`, O lives in `e`.  result = (E >> 32) + O
    Fp<F> r;
    r.l[0] = add_cc(e[0], o[1]);
#pragma unroll
    for (int k = 1; k < 7; k++) r.l[k] = addc_cc(e[k], o[k + 1]);
    r.l[7] = addc(e[7], 0u);
    fp_reduce_once(r);
    return r;
#endif
}

template <class F>
SBN_HD Fp<F> fp_sqr(const Fp<F>& a) { return fp_mul(a, a); }

// Montgomery -> canonical: a * 1 / R
template <class F>
SBN_HD Fp<F> fp_from_mont(const Fp<F>& a) {
    Fp<F> one = Fp<F>::zero();
    one.l[0] = 1;
    return fp_mul(a, one);
}
// canonical (any 256-bit value) -> Montgomery: v * R2 / R
template <class F>
SBN_HD Fp<F> fp_to_mont(const Fp<F>& v) {
    Fp<F> r2;
#pragma unroll
    for (int i = 0; i < 8; i++) r2.l[i] = F::R2(i);
    return fp_mul(r2, v);
}

// a^(p-2) by square-and-multiply over the constant exponent (variable time in nothing secret)
template <class F>
SBN_HD Fp<F> fp_inv(const Fp<F>& a) {
    Fp<F> acc = Fp<F>::one();
    // exponent p - 2, MSB first
    for (int i = 7; i >= 0; i--) {
        uint32_t w = F::P(i);
        if (i == 0) w -= 2;  // low limb of both moduli is >= 2, no borrow
        for (int b = 31; b >= 0; b--) {
            acc = fp_sqr(acc);
            if ((w >> b) & 1) acc = fp_mul(acc, a);
        }
    }
    return acc;
}

// ------------------------------------------------------------------------------------------------
// Inversion by the Bernstein-Yang "safegcd" division steps (2019/266), in the 30-bit signed-limb form libsecp256k1
// popularised: 20 batches of 30 branch-free division steps on the low words of (f, g) = (p, x), each batch summarised by a
// 2x2 integer matrix that is then applied to the full-width (f, g) and, modulo p, to (d, e).  After 600 steps g = 0,
// f = +-1 and d = +-x^-1.  About 20 k instructions against ~108 k for Fermat's x^(p-2) (380 Montgomery products), and no
// data-dependent branch -- the 32 lanes of a warp invert 32 different values in lockstep.  Works on plain integers mod p.
// ------------------------------------------------------------------------------------------------
template <class F>
SBN_HD void fp_inv_plain(const uint32_t x[8], uint32_t out[8]) {
    const int32_t M30 = (int32_t)((1u << 30) - 1);
    int32_t d[9], e[9], f[9], g[9];
#pragma unroll
    for (int i = 0; i < 9; i++) { d[i] = 0; e[i] = 0; f[i] = F::P30(i); }
    e[0] = 1;
    // 8 x 32 bits -> 9 x 30 bits
#pragma unroll
    for (int i = 0; i < 9; i++) {
        const int bit = 30 * i, w = bit >> 5, off = bit & 31;
        uint32_t v = w < 8 ? x[w] >> off : 0u;
        if (off > 2 && w + 1 < 8) v |= x[w + 1] << (32 - off);
        g[i] = (int32_t)(v & (uint32_t)M30);
    }
    int32_t zeta = -1;   // -(delta + 1/2), delta = 1/2 initially
#pragma unroll 1
    for (int batch = 0; batch < 20; batch++) {
        // 30 division steps on the low limbs
        uint32_t u = 1, v = 0, q = 0, r = 1;
        uint32_t fl = (uint32_t)f[0], gl = (uint32_t)g[0];
#pragma unroll 6
        for (int i = 0; i < 30; i++) {
            uint32_t c1 = (uint32_t)(zeta >> 31);
            const uint32_t c2 = 0u - (gl & 1u);
            const uint32_t xx = (fl ^ c1) - c1, yy = (u ^ c1) - c1, zz = (v ^ c1) - c1;
            gl += xx & c2; q += yy & c2; r += zz & c2;
            c1 &= c2;
            zeta = (int32_t)(((uint32_t)zeta ^ c1) - 1u);
            fl += gl & c1; u += q & c1; v += r & c1;
            gl >>= 1; u <<= 1; v <<= 1;
        }
        const int64_t tu = (int32_t)u, tv = (int32_t)v, tq = (int32_t)q, tr = (int32_t)r;
        // (d, e) <- t * (d, e) / 2^30 mod p
        {
            const int32_t sd = d[8] >> 31, se = e[8] >> 31;
            int32_t md = ((int32_t)tu & sd) + ((int32_t)tv & se), me = ((int32_t)tq & sd) + ((int32_t)tr & se);
            int64_t cd = tu * d[0] + tv * e[0], ce = tq * d[0] + tr * e[0];
            md -= (int32_t)((F::INV30 * (uint32_t)cd + (uint32_t)md) & (uint32_t)M30);
            me -= (int32_t)((F::INV30 * (uint32_t)ce + (uint32_t)me) & (uint32_t)M30);
            cd += (int64_t)F::P30(0) * md;
            ce += (int64_t)F::P30(0) * me;
            cd >>= 30; ce >>= 30;
#pragma unroll
            for (int i = 1; i < 9; i++) {
                cd += tu * d[i] + tv * e[i] + (int64_t)F::P30(i) * md;
                ce += tq * d[i] + tr * e[i] + (int64_t)F::P30(i) * me;
                d[i - 1] = (int32_t)cd & M30; cd >>= 30;
                e[i - 1] = (int32_t)ce & M30; ce >>= 30;
            }
            d[8] = (int32_t)cd;
            e[8] = (int32_t)ce;
        }
        // (f, g) <- t * (f, g) / 2^30
        {
            int64_t cf = tu * f[0] + tv * g[0], cg = tq * f[0] + tr * g[0];
            cf >>= 30; cg >>= 30;
#pragma unroll
            for (int i = 1; i < 9; i++) {
                cf += tu * f[i] + tv * g[i];
                cg += tq * f[i] + tr * g[i];
                f[i - 1] = (int32_t)cf & M30; cf >>= 30;
                g[i - 1] = (int32_t)cg & M30; cg >>= 30;
            }
            f[8] = (int32_t)cf;
            g[8] = (int32_t)cg;
        }
    }
    // d in (-2p, p), times the sign of f; bring it to [0, p)
    {
        int32_t cond_add = d[8] >> 31;
#pragma unroll
        for (int i = 0; i < 9; i++) d[i] += F::P30(i) & cond_add;
        const int32_t cond_neg = f[8] >> 31;
#pragma unroll
        for (int i = 0; i < 9; i++) d[i] = (d[i] ^ cond_neg) - cond_neg;
#pragma unroll
        for (int i = 0; i < 8; i++) { d[i + 1] += d[i] >> 30; d[i] &= M30; }
        cond_add = d[8] >> 31;
#pragma unroll
        for (int i = 0; i < 9; i++) d[i] += F::P30(i) & cond_add;
#pragma unroll
        for (int i = 0; i < 8; i++) { d[i + 1] += d[i] >> 30; d[i] &= M30; }
    }
    // 9 x 30 bits -> 8 x 32 bits
#pragma unroll
    for (int w = 0; w < 8; w++) {
        const int bit = 32 * w, i = bit / 30, off = bit % 30;
        uint32_t v = (uint32_t)d[i] >> off;
        v |= (uint32_t)d[i + 1] << (30 - off);
        if (30 - off + 30 < 32 && i + 2 < 9) v |= (uint32_t)d[i + 2] << (60 - off);
        out[w] = v;
    }
}

// Montgomery-form inverse: a = xR  ->  x^-1 R = (xR)^-1 * R^2 = montmul((xR)^-1, R^3).  a == 0 returns 0.
template <class F>
SBN_HD Fp<F> fp_inv_fast(const Fp<F>& a) {
    Fp<F> t;
    fp_inv_plain<F>(a.l, t.l);
    Fp<F> r3;
#pragma unroll
    for (int i = 0; i < 8; i++) r3.l[i] = F::R3(i);
    return fp_mul(t, r3);
}

typedef Fp<FqParams> Fq;
typedef Fp<FrParams> Fr;

}  // namespace sbn
