// Host-side unit test of the device field arithmetic (spartan_bn254_b200/csrc/fp.cuh): the PTX
// carry-flag primitives are emulated on the CPU, so the exact multiply/add/sub algorithm that runs
// on the GPU is checked limb-for-limb against the C oracle (oracle/bn254_oracle.c).
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include "../../spartan_bn254_b200/csrc/fp.cuh"
#include "../../oracle/bn254_oracle.h"

using namespace sbn;

static uint64_t sm_state = 12345;
static uint64_t splitmix() {
    uint64_t z = (sm_state += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

template <class F>
static void rand_elem(int mod, Fp<F>& out, int kind) {
    uint64_t c[4];
    for (int i = 0; i < 4; i++) c[i] = splitmix();
    c[3] &= 0x1fffffffffffffffULL;  // < 2^253 < modulus
    if (kind == 1) { c[0] = c[1] = c[2] = c[3] = 0; }
    if (kind == 2) { c[0] = 1; c[1] = c[2] = c[3] = 0; }
    if (kind == 3) {  // p - 1
        for (int i = 0; i < 4; i++) c[i] = (uint64_t)F::P(2 * i) | ((uint64_t)F::P(2 * i + 1) << 32);
        c[0] -= 1;
    }
    if (kind == 4) { c[0] = c[1] = c[2] = ~0ULL; c[3] = 0x0fffffffffffffffULL; }
    (void)mod;
    memcpy(out.l, c, 32);
}

template <class F>
static int run(int mod, const char* name) {
    int fails = 0;
    for (int it = 0; it < 200000; it++) {
        Fp<F> a, b;
        rand_elem<F>(mod, a, it < 25 ? it / 5 : 0);
        rand_elem<F>(mod, b, it < 25 ? it % 5 : 0);
        ofp oa, ob, om, os, od;
        memcpy(oa.l, a.l, 32);
        memcpy(ob.l, b.l, 32);
        orc_fp_mul(mod, &oa, &ob, &om);
        orc_fp_add(mod, &oa, &ob, &os);
        orc_fp_sub(mod, &oa, &ob, &od);
        Fp<F> m = fp_mul(a, b), s = fp_add(a, b), d = fp_sub(a, b);
        if (memcmp(m.l, om.l, 32) || memcmp(s.l, os.l, 32) || memcmp(d.l, od.l, 32)) {
            if (fails < 5) printf("%s mismatch at it=%d (mul %d add %d sub %d)\n", name, it, memcmp(m.l, om.l, 32) != 0, memcmp(s.l, os.l, 32) != 0, memcmp(d.l, od.l, 32) != 0);
            fails++;
        }
    }
    // unreduced second operand (used by to_mont of arbitrary 256-bit input)
    for (int it = 0; it < 1000; it++) {
        Fp<F> a, v;
        rand_elem<F>(mod, a, 0);
        uint64_t c[4];
        for (int i = 0; i < 4; i++) c[i] = splitmix();
        if (it == 0) c[0] = c[1] = c[2] = c[3] = ~0ULL;
        memcpy(v.l, c, 32);
        ofp oa, ov, om;
        memcpy(oa.l, a.l, 32);
        memcpy(ov.l, v.l, 32);
        orc_fp_mul(mod, &oa, &ov, &om);
        Fp<F> m = fp_mul(a, v);
        if (memcmp(m.l, om.l, 32)) { if (fails < 5) printf("%s unreduced-b mismatch\n", name); fails++; }
    }
    // inverse
    for (int it = 0; it < 20; it++) {
        Fp<F> a;
        rand_elem<F>(mod, a, 0);
        Fp<F> inv = fp_inv(a), prod = fp_mul(a, inv), one = Fp<F>::one();
        if (!(prod == one)) { printf("%s inverse mismatch\n", name); fails++; }
    }
    // safegcd inverse (fp_inv_fast) against the oracle's inverse and the Fermat chain, including 0, 1, p - 1
    for (int it = 0; it < 5000; it++) {
        Fp<F> a;
        rand_elem<F>(mod, a, it < 5 ? it : 0);
        ofp oa, oi;
        memcpy(oa.l, a.l, 32);
        const int nonzero = orc_fp_inv(mod, &oa, &oi);
        Fp<F> fast = fp_inv_fast(a);
        if (!nonzero) { if (!fast.is_zero()) { printf("%s fast inverse of 0 != 0\n", name); fails++; } continue; }
        if (memcmp(fast.l, oi.l, 32)) { if (fails < 5) printf("%s fast inverse mismatch at it=%d\n", name, it); fails++; }
        if (it < 50 && !(fast == fp_inv(a))) { printf("%s fast inverse != Fermat\n", name); fails++; }
    }
    // mont round trip
    {
        Fp<F> a;
        rand_elem<F>(mod, a, 0);
        Fp<F> back = fp_to_mont(fp_from_mont(a));
        if (!(back == a)) { printf("%s mont round trip mismatch\n", name); fails++; }
    }
    printf("%s: %s\n", name, fails ? "FAIL" : "ok");
    return fails;
}

int main() {
    int f = run<FqParams>(0, "Fq") + run<FrParams>(1, "Fr");
    return f ? 1 : 0;
}
