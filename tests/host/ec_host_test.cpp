// Host-side unit test of the device point arithmetic (spartan_bn254_b200/csrc/ec.cuh) against the
// C oracle: mixed add, full add, doubling, the P+P / P+(-P) / identity special cases, to_affine.
#include <cstdio>
#include <cstring>
#include <vector>
#include "../../spartan_bn254_b200/csrc/ec.cuh"
#include "../../oracle/bn254_oracle.h"
using namespace sbn;

static uint64_t st = 99;
static uint64_t splitmix() {
    uint64_t z = (st += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static Affine from_o(const og1a& p, uint8_t inf) { Affine a; if (inf) return Affine::identity(); memcpy(a.x.l, p.x.l, 32); memcpy(a.y.l, p.y.l, 32); return a; }
static bool same(const Affine& a, const og1a& p, uint8_t inf) {
    if (inf) return a.is_identity();
    return !memcmp(a.x.l, p.x.l, 32) && !memcmp(a.y.l, p.y.l, 32);
}

int main() {
    og1a g; orc_g1_generator(&g);
    const int N = 64;
    std::vector<og1a> pts(N); std::vector<uint8_t> inf(N);
    for (int i = 0; i < N; i++) {
        uint64_t c[4] = {splitmix(), splitmix(), splitmix(), splitmix() & 0x0fffffffffffffffULL};
        ofp s; orc_fp_from_u64x4(1, c, &s);
        orc_g1_scalar_mul(&g, 0, &s, &pts[i], &inf[i]);
    }
    int fails = 0;
    // running mixed-add chain with deliberate repeats (P+P) and cancellations (P + -P)
    XYZZ acc = XYZZ::identity();
    og1a oacc; uint8_t oinf = 1; memset(&oacc, 0, sizeof oacc);
    for (int it = 0; it < 400; it++) {
        int idx = splitmix() % N;
        Affine q = from_o(pts[idx], 0);
        og1a oq = pts[idx];
        int mode = it % 7;
        if (mode == 3) { // add the current accumulator's own affine value -> doubling path
            if (!oinf) { q = from_o(oacc, 0); oq = oacc; }
        } else if (mode == 5) { // add the negation -> identity path
            if (!oinf) { q = affine_neg(from_o(oacc, 0)); memcpy(oq.x.l, q.x.l, 32); memcpy(oq.y.l, q.y.l, 32); }
        } else if (mode == 6) { q = affine_neg(q); memcpy(oq.y.l, q.y.l, 32); }
        xyzz_add_mixed(acc, q);
        orc_g1_add_affine(&oacc, oinf, &oq, 0, &oacc, &oinf);
        Affine a = xyzz_to_affine(acc);
        if (!same(a, oacc, oinf)) { if (fails < 5) printf("mixed chain mismatch it=%d mode=%d\n", it, mode); fails++; }
    }
    // full adds: (sum of first k) + (sum of next k), plus self-add and cancel
    for (int it = 0; it < 100; it++) {
        XYZZ A = XYZZ::identity(), B = XYZZ::identity();
        og1a oa, ob; uint8_t ia = 1, ib = 1; memset(&oa, 0, sizeof oa); memset(&ob, 0, sizeof ob);
        int ka = splitmix() % 4, kb = splitmix() % 4;
        for (int k = 0; k < ka; k++) { int idx = splitmix() % N; xyzz_add_mixed(A, from_o(pts[idx], 0)); orc_g1_add_affine(&oa, ia, &pts[idx], 0, &oa, &ia); }
        for (int k = 0; k < kb; k++) { int idx = splitmix() % N; xyzz_add_mixed(B, from_o(pts[idx], 0)); orc_g1_add_affine(&ob, ib, &pts[idx], 0, &ob, &ib); }
        if (it % 5 == 1) { B = A; ob = oa; ib = ia; B = xyzz_dbl(B); B = A; }           // A + A with different-looking reps below
        if (it % 5 == 2 && !ia) { B = XYZZ::from_affine(affine_neg(xyzz_to_affine(A))); Affine nb = xyzz_to_affine(B); memcpy(ob.x.l, nb.x.l, 32); memcpy(ob.y.l, nb.y.l, 32); ib = 0; }
        XYZZ S = A; xyzz_add(S, B);
        og1a os; uint8_t is;
        orc_g1_add_affine(&oa, ia, &ob, ib, &os, &is);
        if (!same(xyzz_to_affine(S), os, is)) { if (fails < 5) printf("full add mismatch it=%d\n", it); fails++; }
        XYZZ D = xyzz_dbl(A);
        orc_g1_add_affine(&oa, ia, &oa, ia, &os, &is);
        if (!same(xyzz_to_affine(D), os, is)) { if (fails < 5) printf("dbl mismatch it=%d\n", it); fails++; }
    }
    printf("ec: %s\n", fails ? "FAIL" : "ok");
    return fails ? 1 : 0;
}
