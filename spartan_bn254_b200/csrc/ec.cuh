// BN254 G1 (y^2 = x^3 + 3) point arithmetic in extended Jacobian "XYZZ" coordinates
// (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2), all coordinates Montgomery Fq.
//
// Bucket accumulation meets P+P and P+(-P) constantly with the reference's generators (about two
// thirds of MultiCommitGens::new outputs are the same point G -- reference group.rs:110-132 falls
// through to Scalar::one()), so every addition here handles doubling / cancellation / identity
// explicitly instead of assuming distinct inputs.
#pragma once
#include "fp.cuh"

namespace sbn {

// Multiplication policy: the bucket-accumulation hot loop inlines the Montgomery product; the
// latency-bound helper kernels (reduce, normalise, tables, scalar-mul) call one out-of-line copy so
// their code stays inside the instruction cache (an inlined XYZZ add is ~60 KB of SASS).
struct MulInline {
    static SBN_HD Fq mul(const Fq& a, const Fq& b) { return fp_mul(a, b); }
};
struct MulCall {
#if defined(__CUDACC__)
    static __host__ __device__ __noinline__ Fq mul(const Fq& a, const Fq& b) { return fp_mul(a, b); }
#else
    static Fq mul(const Fq& a, const Fq& b) { return fp_mul(a, b); }
#endif
};

struct Affine {   // 64 B, layout of sbn_g1a in include/sbn254.h; (0,0) encodes the identity
    Fq x, y;
    SBN_HD bool is_identity() const { return x.is_zero() && y.is_zero(); }
    static SBN_HD Affine identity() { Affine a; a.x = Fq::zero(); a.y = Fq::zero(); return a; }
};

struct XYZZ {     // 128 B; ZZ == 0 encodes the identity
    Fq X, Y, ZZ, ZZZ;
    SBN_HD bool is_identity() const { return ZZ.is_zero(); }
    static SBN_HD XYZZ identity() {
        XYZZ r; r.X = Fq::zero(); r.Y = Fq::zero(); r.ZZ = Fq::zero(); r.ZZZ = Fq::zero(); return r;
    }
    static SBN_HD XYZZ from_affine(const Affine& a) {
        XYZZ r;
        if (a.is_identity()) return identity();
        r.X = a.x; r.Y = a.y; r.ZZ = Fq::one(); r.ZZZ = Fq::one();
        return r;
    }
};

// 2 * (x, y) for an affine input (mdbl-2008-s-1, a = 0)
template <class MP = MulInline>
SBN_HD XYZZ xyzz_dbl_affine(const Affine& p) {
    XYZZ r;
    Fq U = fp_dbl(p.y);
    Fq V = MP::mul(U, U);
    Fq W = MP::mul(U, V);
    Fq S = MP::mul(p.x, V);
    Fq xx = MP::mul(p.x, p.x);
    Fq M = fp_add(fp_dbl(xx), xx);
    r.X = fp_sub(MP::mul(M, M), fp_dbl(S));
    r.Y = fp_sub(MP::mul(M, fp_sub(S, r.X)), MP::mul(W, p.y));
    r.ZZ = V;
    r.ZZZ = W;
    return r;
}

// 2 * P (dbl-2008-s-1, a = 0).  y == 0 cannot happen on a prime-order curve.
template <class MP = MulInline>
SBN_HD XYZZ xyzz_dbl(const XYZZ& p) {
    if (p.is_identity()) return p;
    XYZZ r;
    Fq U = fp_dbl(p.Y);
    Fq V = MP::mul(U, U);
    Fq W = MP::mul(U, V);
    Fq S = MP::mul(p.X, V);
    Fq xx = MP::mul(p.X, p.X);
    Fq M = fp_add(fp_dbl(xx), xx);
    r.X = fp_sub(MP::mul(M, M), fp_dbl(S));
    r.Y = fp_sub(MP::mul(M, fp_sub(S, r.X)), MP::mul(W, p.Y));
    r.ZZ = MP::mul(V, p.ZZ);
    r.ZZZ = MP::mul(W, p.ZZZ);
    return r;
}

// acc += q  (mixed addition madd-2008-s: 8M + 2S), q affine and not the identity
template <class MP = MulInline>
SBN_HD void xyzz_add_mixed(XYZZ& acc, const Affine& q) {
    if (acc.is_identity()) { acc.X = q.x; acc.Y = q.y; acc.ZZ = Fq::one(); acc.ZZZ = Fq::one(); return; }
    Fq U2 = MP::mul(q.x, acc.ZZ);
    Fq S2 = MP::mul(q.y, acc.ZZZ);
    Fq P = fp_sub(U2, acc.X);
    Fq R = fp_sub(S2, acc.Y);
    if (P.is_zero()) {
        if (R.is_zero()) acc = xyzz_dbl_affine<MP>(q);   // same point
        else acc = XYZZ::identity();                  // opposite points
        return;
    }
    Fq PP = MP::mul(P, P);
    Fq PPP = MP::mul(P, PP);
    Fq Q = MP::mul(acc.X, PP);
    Fq X3 = fp_sub(fp_sub(MP::mul(R, R), PPP), fp_dbl(Q));
    Fq Y3 = fp_sub(MP::mul(R, fp_sub(Q, X3)), MP::mul(acc.Y, PPP));
    acc.X = X3;
    acc.Y = Y3;
    acc.ZZ = MP::mul(acc.ZZ, PP);
    acc.ZZZ = MP::mul(acc.ZZZ, PPP);
}

// acc += q  (add-2008-s: 12M + 2S)
// MP: policy of the common path; MPD: policy of the rare doubling path (kept out of line by callers
// that care about code size)
template <class MP = MulInline, class MPD = MP>
SBN_HD void xyzz_add(XYZZ& acc, const XYZZ& q) {
    if (q.is_identity()) return;
    if (acc.is_identity()) { acc = q; return; }
    Fq U1 = MP::mul(acc.X, q.ZZ);
    Fq U2 = MP::mul(q.X, acc.ZZ);
    Fq S1 = MP::mul(acc.Y, q.ZZZ);
    Fq S2 = MP::mul(q.Y, acc.ZZZ);
    Fq P = fp_sub(U2, U1);
    Fq R = fp_sub(S2, S1);
    if (P.is_zero()) {
        if (R.is_zero()) acc = xyzz_dbl<MPD>(acc);
        else acc = XYZZ::identity();
        return;
    }
    Fq PP = MP::mul(P, P);
    Fq PPP = MP::mul(P, PP);
    Fq Q = MP::mul(U1, PP);
    Fq X3 = fp_sub(fp_sub(MP::mul(R, R), PPP), fp_dbl(Q));
    Fq Y3 = fp_sub(MP::mul(R, fp_sub(Q, X3)), MP::mul(S1, PPP));
    acc.X = X3;
    acc.Y = Y3;
    acc.ZZ = MP::mul(MP::mul(acc.ZZ, q.ZZ), PP);
    acc.ZZZ = MP::mul(MP::mul(acc.ZZZ, q.ZZZ), PPP);
}

SBN_HD Affine affine_neg(const Affine& a) {
    Affine r;
    r.x = a.x;
    r.y = a.y.is_zero() ? a.y : fp_neg(a.y);   // keeps (0,0) as the identity encoding
    return r;
}

// a^(p-2) with the multiplication policy M
template <class MP>
SBN_HD Fq fq_inv(const Fq& a) {
    Fq acc = Fq::one();
    for (int i = 7; i >= 0; i--) {
        uint32_t w = FqParams::P(i);
        if (i == 0) w -= 2;
        for (int b = 31; b >= 0; b--) {
            acc = MP::mul(acc, acc);
            if ((w >> b) & 1) acc = MP::mul(acc, a);
        }
    }
    return acc;
}

// XYZZ -> affine with one field inversion: I = 1/(ZZ*ZZZ)  =>  1/ZZ = ZZZ * I,  1/ZZZ = ZZ * I.
template <class MP = MulInline>
SBN_HD Affine xyzz_to_affine(const XYZZ& p) {
    if (p.is_identity()) return Affine::identity();
    Fq I = fp_inv_fast(MP::mul(p.ZZ, p.ZZZ));      // safegcd division steps: ~5x fewer instructions than x^(p-2)
    Affine r;
    r.x = MP::mul(p.X, MP::mul(p.ZZZ, I));
    r.y = MP::mul(p.Y, MP::mul(p.ZZ, I));
    return r;
}

}  // namespace sbn
