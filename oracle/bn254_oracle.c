/* TEST INFRASTRUCTURE ONLY -- see bn254_oracle.h.  Plain-C CPU restatement of the reference's
 * algorithm for the Hyrax commit / opening path.  All curve / field / MSM arithmetic in the
 * reference lives in third-party crates absent from /root/reference (ark-ff / ark-ec / ark-bn254
 * ^0.5, merlin 3.0, sha3 0.10 -- Cargo.toml:9-29, no Cargo.lock); their *published* algorithms are
 * restated here (Montgomery CIOS Fp256, Jacobian short-Weierstrass a=0, signed-window Pippenger,
 * FIPS-202 Keccak, STROBE-128) and parity is anchored on the reference's own call sites:
 *
 *   group.rs:110-132   from_uniform_bytes            -> gen_scalar_from_chunk()
 *   group.rs:135-140   compress                      -> orc_g1_compress()
 *   group.rs:143-175   vartime_multiscalar_mul / msm_affine -> orc_msm()
 *   commitments.rs:31-62   MultiCommitGens::new      -> orc_gen_scalars(), orc_multi_commit_gens()
 *   commitments.rs:144-154 <[Scalar]>::commit        -> commit_row()
 *   hyrax.rs:253-281   DensePolynomial::commit_inner -> orc_hyrax_commit()
 *   hyrax.rs:311-324   DensePolynomial::bound        -> orc_bound()
 *   hyrax.rs:355-369   EqPolynomial::evals           -> orc_eq_evals()
 *   hyrax.rs:195-203   bound_poly_var_top            -> orc_bind_top()
 *   nizk/bullet.rs:24-126 BulletReductionProof::prove-> orc_bullet_prove()
 *   sumcheck.rs:501-530 cubic round evaluation       -> orc_sumcheck_cubic_eval()
 *   transcript.rs:56-67 challenge_scalar             -> orc_transcript_challenge_scalar()
 *
 * PARITY STATUS: "parity unpinned" by reference fixtures (none exist); checked
 * against oracle/pymodel.py (independent big-int model) and public BN254 / Merlin test vectors. */
#include "bn254_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>
#include <time.h>

/* minimal parallel-for over [0,n) with dynamic chunking (the image's gcc has no libgomp) */
typedef void (*pf_body)(long i, void* ctx);
typedef struct { pf_body body; void* ctx; long n; long chunk; long next; pthread_mutex_t mu; } pf_job;
static void* pf_worker(void* arg) {
    pf_job* j = (pf_job*)arg;
    for (;;) {
        pthread_mutex_lock(&j->mu);
        long lo = j->next;
        j->next += j->chunk;
        pthread_mutex_unlock(&j->mu);
        if (lo >= j->n) break;
        long hi = lo + j->chunk < j->n ? lo + j->chunk : j->n;
        for (long i = lo; i < hi; i++) j->body(i, j->ctx);
    }
    return NULL;
}
static int host_threads(int threads) {
    if (threads > 0) return threads;
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}
static void parallel_for(long n, long chunk, int threads, pf_body body, void* ctx) {
    threads = host_threads(threads);
    if (threads > n) threads = (int)(n > 0 ? n : 1);
    pf_job job = {body, ctx, n, chunk > 0 ? chunk : 1, 0, PTHREAD_MUTEX_INITIALIZER};
    if (threads <= 1) { pf_worker(&job); return; }
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)threads);
    for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, pf_worker, &job);
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    free(th);
}

typedef unsigned __int128 u128;

typedef struct {
    uint64_t p[4];
    uint64_t inv;      /* -p^{-1} mod 2^64 */
    uint64_t r1[4];    /* R mod p  (Montgomery one) */
    uint64_t r2[4];    /* R^2 mod p */
} field_t;

static const field_t FIELDS[2] = {
    { /* Fq */
        {0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL},
        0x87d20782e4866389ULL,
        {0xd35d438dc58f0d9dULL, 0x0a78eb28f5c70b3dULL, 0x666ea36f7879462cULL, 0x0e0a77c19a07df2fULL},
        {0xf32cfc5b538afa89ULL, 0xb5e71911d44501fbULL, 0x47ab1eff0a417ff6ULL, 0x06d89f71cab8351fULL},
    },
    { /* Fr */
        {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL},
        0xc2e1f593efffffffULL,
        {0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL, 0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL},
        {0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL},
    },
};
#define FQ 0
#define FR 1

/* ------------------------------------------------------------------ field */
static inline int geq(const uint64_t a[4], const uint64_t b[4]) {
    for (int i = 3; i >= 0; i--) {
        if (a[i] > b[i]) return 1;
        if (a[i] < b[i]) return 0;
    }
    return 1;
}
static inline uint64_t sub4(uint64_t r[4], const uint64_t a[4], const uint64_t b[4]) {
    u128 borrow = 0;
    for (int i = 0; i < 4; i++) {
        u128 t = (u128)a[i] - b[i] - borrow;
        r[i] = (uint64_t)t;
        borrow = (t >> 64) & 1;
    }
    return (uint64_t)borrow;
}
static inline uint64_t add4(uint64_t r[4], const uint64_t a[4], const uint64_t b[4]) {
    u128 c = 0;
    for (int i = 0; i < 4; i++) {
        c += (u128)a[i] + b[i];
        r[i] = (uint64_t)c;
        c >>= 64;
    }
    return (uint64_t)c;
}
static inline int is_zero4(const uint64_t a[4]) { return (a[0] | a[1] | a[2] | a[3]) == 0; }

/* Montgomery product, CIOS (the same word-serial algorithm as before) with the accumulator in five named words so that the
 * compiler keeps it in registers and emits mul / adc chains. */
#define ORC_MAC(lo, hi, a, b, c, d) do { u128 _t = (u128)(a) * (b) + (c) + (d); lo = (uint64_t)_t; hi = (uint64_t)(_t >> 64); } while (0)
static inline __attribute__((always_inline)) void fp_mul(int m, const ofp* a, const ofp* b, ofp* out) {
    const field_t* F = &FIELDS[m];
    const uint64_t p0 = F->p[0], p1 = F->p[1], p2 = F->p[2], p3 = F->p[3], inv = F->inv;
    const uint64_t a0 = a->l[0], a1 = a->l[1], a2 = a->l[2], a3 = a->l[3];
    uint64_t t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0;
    for (int i = 0; i < 4; i++) {
        const uint64_t bi = b->l[i];
        uint64_t c, t5;
        ORC_MAC(t0, c, a0, bi, t0, 0);
        ORC_MAC(t1, c, a1, bi, t1, c);
        ORC_MAC(t2, c, a2, bi, t2, c);
        ORC_MAC(t3, c, a3, bi, t3, c);
        { u128 s = (u128)t4 + c; t4 = (uint64_t)s; t5 = (uint64_t)(s >> 64); }
        const uint64_t mm = t0 * inv;
        uint64_t dump;
        ORC_MAC(dump, c, mm, p0, t0, 0);
        (void)dump;
        ORC_MAC(t0, c, mm, p1, t1, c);
        ORC_MAC(t1, c, mm, p2, t2, c);
        ORC_MAC(t2, c, mm, p3, t3, c);
        { u128 s = (u128)t4 + c; t3 = (uint64_t)s; t4 = t5 + (uint64_t)(s >> 64); }
    }
    uint64_t r[4] = {t0, t1, t2, t3};
    if (t4 || geq(r, F->p)) sub4(r, r, F->p);
    memcpy(out->l, r, 32);
}
static inline __attribute__((always_inline)) void fp_add(int m, const ofp* a, const ofp* b, ofp* out) {
    const field_t* F = &FIELDS[m];
    uint64_t r[4];
    uint64_t c = add4(r, a->l, b->l);
    if (c || geq(r, F->p)) sub4(r, r, F->p);
    memcpy(out->l, r, 32);
}
static inline __attribute__((always_inline)) void fp_sub(int m, const ofp* a, const ofp* b, ofp* out) {
    const field_t* F = &FIELDS[m];
    uint64_t r[4];
    if (sub4(r, a->l, b->l)) add4(r, r, F->p);
    memcpy(out->l, r, 32);
}
static inline __attribute__((always_inline)) void fp_neg(int m, const ofp* a, ofp* out) {
    ofp z = {{0, 0, 0, 0}};
    fp_sub(m, &z, a, out);
}
static inline int fp_is_zero(const ofp* a) { return is_zero4(a->l); }
static inline int fp_eq(const ofp* a, const ofp* b) { return memcmp(a->l, b->l, 32) == 0; }
static void fp_one(int m, ofp* out) { memcpy(out->l, FIELDS[m].r1, 32); }
static void fp_from_canon(int m, const uint64_t c[4], ofp* out) {
    ofp a, r2;
    memcpy(a.l, c, 32);
    memcpy(r2.l, FIELDS[m].r2, 32);
    fp_mul(m, &a, &r2, out);
}
static void fp_to_canon(int m, const ofp* in, uint64_t c[4]) {
    ofp one = {{1, 0, 0, 0}}, o;
    fp_mul(m, in, &one, &o);
    memcpy(c, o.l, 32);
}
static void fp_pow(int m, const ofp* a, const uint64_t e[4], ofp* out) {
    ofp acc, base = *a;
    fp_one(m, &acc);
    for (int i = 0; i < 256; i++) {
        if ((e[i / 64] >> (i % 64)) & 1) fp_mul(m, &acc, &base, &acc);
        fp_mul(m, &base, &base, &base);
    }
    *out = acc;
}
static int fp_inv(int m, const ofp* a, ofp* out) {
    if (fp_is_zero(a)) { memset(out, 0, sizeof *out); return 0; }
    uint64_t e[4], two[4] = {2, 0, 0, 0};
    sub4(e, FIELDS[m].p, two);
    fp_pow(m, a, e, out);
    return 1;
}

void orc_fp_from_u64x4(int mod, const uint64_t c[4], ofp* out) { fp_from_canon(mod, c, out); }
void orc_fp_to_u64x4(int mod, const ofp* in, uint64_t c[4]) { fp_to_canon(mod, in, c); }
void orc_fp_mul(int mod, const ofp* a, const ofp* b, ofp* out) { fp_mul(mod, a, b, out); }
void orc_fp_add(int mod, const ofp* a, const ofp* b, ofp* out) { fp_add(mod, a, b, out); }
void orc_fp_sub(int mod, const ofp* a, const ofp* b, ofp* out) { fp_sub(mod, a, b, out); }
int orc_fp_inv(int mod, const ofp* a, ofp* out) { return fp_inv(mod, a, out); }

/* ------------------------------------------------------------------ G1 (y^2 = x^3 + 3), Jacobian */
static void j_identity(og1j* p) {
    fp_one(FQ, &p->X);
    fp_one(FQ, &p->Y);
    memset(&p->Z, 0, sizeof p->Z);
}
static inline int j_is_identity(const og1j* p) { return fp_is_zero(&p->Z); }

static void j_double(const og1j* p, og1j* out) { /* dbl-2009-l, a = 0 */
    if (j_is_identity(p)) { *out = *p; return; }
    ofp A, B, C, D, E, F, t, X3, Y3, Z3;
    fp_mul(FQ, &p->X, &p->X, &A);
    fp_mul(FQ, &p->Y, &p->Y, &B);
    fp_mul(FQ, &B, &B, &C);
    fp_add(FQ, &p->X, &B, &t);
    fp_mul(FQ, &t, &t, &t);
    fp_sub(FQ, &t, &A, &t);
    fp_sub(FQ, &t, &C, &t);
    fp_add(FQ, &t, &t, &D);
    fp_add(FQ, &A, &A, &E);
    fp_add(FQ, &E, &A, &E);
    fp_mul(FQ, &E, &E, &F);
    fp_sub(FQ, &F, &D, &X3);
    fp_sub(FQ, &X3, &D, &X3);
    fp_sub(FQ, &D, &X3, &t);
    fp_mul(FQ, &E, &t, &Y3);
    fp_add(FQ, &C, &C, &t);
    fp_add(FQ, &t, &t, &t);
    fp_add(FQ, &t, &t, &t);
    fp_sub(FQ, &Y3, &t, &Y3);
    fp_mul(FQ, &p->Y, &p->Z, &Z3);
    fp_add(FQ, &Z3, &Z3, &Z3);
    out->X = X3; out->Y = Y3; out->Z = Z3;
}

static void j_add_mixed(const og1j* p, const og1a* q, og1j* out) { /* madd-2007-bl style */
    if (j_is_identity(p)) { out->X = q->x; out->Y = q->y; fp_one(FQ, &out->Z); return; }
    ofp Z1Z1, U2, S2, H, r, HH, HHH, V, t, X3, Y3, Z3;
    fp_mul(FQ, &p->Z, &p->Z, &Z1Z1);
    fp_mul(FQ, &q->x, &Z1Z1, &U2);
    fp_mul(FQ, &q->y, &p->Z, &S2);
    fp_mul(FQ, &S2, &Z1Z1, &S2);
    fp_sub(FQ, &U2, &p->X, &H);
    fp_sub(FQ, &S2, &p->Y, &r);
    if (fp_is_zero(&H)) {
        if (fp_is_zero(&r)) { j_double(p, out); return; }
        j_identity(out);
        return;
    }
    fp_mul(FQ, &H, &H, &HH);
    fp_mul(FQ, &H, &HH, &HHH);
    fp_mul(FQ, &p->X, &HH, &V);
    fp_mul(FQ, &r, &r, &X3);
    fp_sub(FQ, &X3, &HHH, &X3);
    fp_sub(FQ, &X3, &V, &X3);
    fp_sub(FQ, &X3, &V, &X3);
    fp_sub(FQ, &V, &X3, &t);
    fp_mul(FQ, &r, &t, &Y3);
    fp_mul(FQ, &p->Y, &HHH, &t);
    fp_sub(FQ, &Y3, &t, &Y3);
    fp_mul(FQ, &p->Z, &H, &Z3);
    out->X = X3; out->Y = Y3; out->Z = Z3;
}

static void j_add(const og1j* p, const og1j* q, og1j* out) { /* add-2007-bl style */
    if (j_is_identity(p)) { *out = *q; return; }
    if (j_is_identity(q)) { *out = *p; return; }
    ofp Z1Z1, Z2Z2, U1, U2, S1, S2, H, r, HH, HHH, V, t, X3, Y3, Z3;
    fp_mul(FQ, &p->Z, &p->Z, &Z1Z1);
    fp_mul(FQ, &q->Z, &q->Z, &Z2Z2);
    fp_mul(FQ, &p->X, &Z2Z2, &U1);
    fp_mul(FQ, &q->X, &Z1Z1, &U2);
    fp_mul(FQ, &p->Y, &q->Z, &S1);
    fp_mul(FQ, &S1, &Z2Z2, &S1);
    fp_mul(FQ, &q->Y, &p->Z, &S2);
    fp_mul(FQ, &S2, &Z1Z1, &S2);
    fp_sub(FQ, &U2, &U1, &H);
    fp_sub(FQ, &S2, &S1, &r);
    if (fp_is_zero(&H)) {
        if (fp_is_zero(&r)) { j_double(p, out); return; }
        j_identity(out);
        return;
    }
    fp_mul(FQ, &H, &H, &HH);
    fp_mul(FQ, &H, &HH, &HHH);
    fp_mul(FQ, &U1, &HH, &V);
    fp_mul(FQ, &r, &r, &X3);
    fp_sub(FQ, &X3, &HHH, &X3);
    fp_sub(FQ, &X3, &V, &X3);
    fp_sub(FQ, &X3, &V, &X3);
    fp_sub(FQ, &V, &X3, &t);
    fp_mul(FQ, &r, &t, &Y3);
    fp_mul(FQ, &S1, &HHH, &t);
    fp_sub(FQ, &Y3, &t, &Y3);
    fp_mul(FQ, &p->Z, &q->Z, &Z3);
    fp_mul(FQ, &Z3, &H, &Z3);
    out->X = X3; out->Y = Y3; out->Z = Z3;
}

static void j_to_affine(const og1j* p, og1a* out, uint8_t* inf) {
    if (j_is_identity(p)) { memset(out, 0, sizeof *out); *inf = 1; return; }
    ofp zi, zi2, zi3;
    fp_inv(FQ, &p->Z, &zi);
    fp_mul(FQ, &zi, &zi, &zi2);
    fp_mul(FQ, &zi2, &zi, &zi3);
    fp_mul(FQ, &p->X, &zi2, &out->x);
    fp_mul(FQ, &p->Y, &zi3, &out->y);
    *inf = 0;
}
static void j_from_affine(const og1a* a, uint8_t inf, og1j* out) {
    if (inf) { j_identity(out); return; }
    out->X = a->x; out->Y = a->y; fp_one(FQ, &out->Z);
}
static void a_neg(const og1a* a, og1a* out) { out->x = a->x; fp_neg(FQ, &a->y, &out->y); }

/* k (canonical 4x64) * P, double-and-add MSB first */
static void j_scalar_mul_canon(const og1a* p, uint8_t inf, const uint64_t k[4], og1j* out) {
    og1j acc;
    j_identity(&acc);
    if (!inf) {
        int started = 0;
        for (int i = 255; i >= 0; i--) {
            if (started) j_double(&acc, &acc);
            if ((k[i / 64] >> (i % 64)) & 1) { j_add_mixed(&acc, p, &acc); started = 1; }
        }
    }
    *out = acc;
}

void orc_g1_generator(og1a* out) {
    uint64_t one[4] = {1, 0, 0, 0}, two[4] = {2, 0, 0, 0};
    fp_from_canon(FQ, one, &out->x);
    fp_from_canon(FQ, two, &out->y);
}
int orc_g1_on_curve(const og1a* p, uint8_t inf) {
    if (inf) return 1;
    ofp y2, x3, b;
    uint64_t three[4] = {3, 0, 0, 0};
    fp_from_canon(FQ, three, &b);
    fp_mul(FQ, &p->y, &p->y, &y2);
    fp_mul(FQ, &p->x, &p->x, &x3);
    fp_mul(FQ, &x3, &p->x, &x3);
    fp_add(FQ, &x3, &b, &x3);
    return fp_eq(&y2, &x3);
}
void orc_g1_add_affine(const og1a* a, uint8_t ainf, const og1a* b, uint8_t binf, og1a* out, uint8_t* oinf) {
    og1j ja;
    j_from_affine(a, ainf, &ja);
    if (!binf) j_add_mixed(&ja, b, &ja);
    j_to_affine(&ja, out, oinf);
}
void orc_g1_scalar_mul(const og1a* p, uint8_t inf, const ofp* s, og1a* out, uint8_t* oinf) {
    uint64_t k[4];
    og1j r;
    fp_to_canon(FR, s, k);
    j_scalar_mul_canon(p, inf, k, &r);
    j_to_affine(&r, out, oinf);
}
void orc_g1_compress(const og1a* p, uint8_t inf, uint8_t out[32]) {
    memset(out, 0, 32);
    if (inf) { out[31] |= 0x40; return; }
    uint64_t x[4], y[4], ny[4];
    ofp negy;
    fp_to_canon(FQ, &p->x, x);
    fp_to_canon(FQ, &p->y, y);
    fp_neg(FQ, &p->y, &negy);
    fp_to_canon(FQ, &negy, ny);
    for (int i = 0; i < 4; i++)
        for (int b = 0; b < 8; b++) out[8 * i + b] = (uint8_t)(x[i] >> (8 * b));
    /* y > -y  <=> !(ny >= y) */
    if (!geq(ny, y)) out[31] |= 0x80;
}

/* ------------------------------------------------------------------ MSM */
static void msm_naive(const og1a* pts, const uint8_t* inf, const ofp* sc, size_t n, og1j* out) {
    og1j acc, t;
    j_identity(&acc);
    for (size_t i = 0; i < n; i++) {
        uint64_t k[4];
        fp_to_canon(FR, &sc[i], k);
        j_scalar_mul_canon(&pts[i], inf ? inf[i] : 0, k, &t);
        j_add(&acc, &t, &acc);
    }
    *out = acc;
}

/* Signed-window Pippenger as published for arkworks' VariableBaseMSM: window c = ln(n)+2 for
 * n >= 32 else 3; digits in [-2^{c-1}, 2^{c-1}); per-window buckets, running-sum reduction,
 * Horner combination with c doublings.  (The result does not depend on any of these choices.) */
static void msm_pippenger(const og1a* pts, const uint8_t* inf, const ofp* sc, size_t n, og1j* out) {
    if (n == 0) { j_identity(out); return; }
    size_t c = 3; /* ark: n < 32 ? 3 : ln_without_floats(n) + 2, ln_without_floats = log2(n) * 69 / 100 */
    if (n >= 32) {
        size_t lg = 0;
        while (((size_t)1 << (lg + 1)) <= n) lg++;
        c = lg * 69 / 100 + 2;
    }
    size_t nwin = (254 + c - 1) / c + 1; /* one spare window absorbs the final carry */
    int32_t* digits = (int32_t*)malloc(sizeof(int32_t) * n * nwin);
    for (size_t i = 0; i < n; i++) {
        uint64_t k[4];
        fp_to_canon(FR, &sc[i], k);
        int64_t carry = 0;
        for (size_t w = 0; w < nwin; w++) {
            size_t bit = w * c;
            uint64_t v = 0;
            if (bit < 256) {
                size_t limb = bit / 64, off = bit % 64;
                v = k[limb] >> off;
                if (off + c > 64 && limb + 1 < 4) v |= k[limb + 1] << (64 - off);
                v &= (((uint64_t)1) << c) - 1;
            }
            int64_t d = (int64_t)v + carry;
            carry = 0;
            if (d >= ((int64_t)1 << (c - 1))) { d -= ((int64_t)1 << c); carry = 1; }
            digits[i * nwin + w] = (int32_t)d;
        }
    }
    size_t nb = (size_t)1 << (c - 1);
    og1j* buckets = (og1j*)malloc(sizeof(og1j) * nb);
    og1j total;
    j_identity(&total);
    for (size_t wi = nwin; wi-- > 0;) {
        for (size_t b = 0; b < nb; b++) j_identity(&buckets[b]);
        for (size_t i = 0; i < n; i++) {
            int32_t d = digits[i * nwin + wi];
            if (d == 0 || (inf && inf[i])) continue;
            if (d > 0) j_add_mixed(&buckets[d - 1], &pts[i], &buckets[d - 1]);
            else { og1a np; a_neg(&pts[i], &np); j_add_mixed(&buckets[-d - 1], &np, &buckets[-d - 1]); }
        }
        og1j run, sum;
        j_identity(&run);
        j_identity(&sum);
        for (size_t b = nb; b-- > 0;) {
            j_add(&run, &buckets[b], &run);
            j_add(&sum, &run, &sum);
        }
        for (size_t k = 0; k < c; k++) j_double(&total, &total);
        j_add(&total, &sum, &total);
    }
    free(buckets);
    free(digits);
    *out = total;
}

void orc_msm(const og1a* pts, const uint8_t* inf, const ofp* sc, size_t n, int algo, og1a* out, uint8_t* oinf) {
    og1j r;
    if (algo == 0) msm_naive(pts, inf, sc, n, &r);
    else msm_pippenger(pts, inf, sc, n, &r);
    j_to_affine(&r, out, oinf);
}

/* ------------------------------------------------------------------ Keccak-f[1600], SHA3-256, SHAKE256 */
static const uint64_t KRC[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
    0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
    0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
    0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
    0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
static const int KROT[24] = {1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 2, 14, 27, 41, 56, 8, 25, 43, 62, 18, 39, 61, 20, 44};
static const int KPIL[24] = {10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4, 15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1};
static inline uint64_t rotl64(uint64_t x, int n) { return (x << n) | (x >> (64 - n)); }
static void keccakf(uint64_t s[25]) {
    for (int round = 0; round < 24; round++) {
        uint64_t bc[5], t;
        for (int i = 0; i < 5; i++) bc[i] = s[i] ^ s[i + 5] ^ s[i + 10] ^ s[i + 15] ^ s[i + 20];
        for (int i = 0; i < 5; i++) {
            t = bc[(i + 4) % 5] ^ rotl64(bc[(i + 1) % 5], 1);
            for (int j = 0; j < 25; j += 5) s[j + i] ^= t;
        }
        t = s[1];
        for (int i = 0; i < 24; i++) {
            int j = KPIL[i];
            uint64_t b = s[j];
            s[j] = rotl64(t, KROT[i]);
            t = b;
        }
        for (int j = 0; j < 25; j += 5) {
            for (int i = 0; i < 5; i++) bc[i] = s[j + i];
            for (int i = 0; i < 5; i++) s[j + i] ^= (~bc[(i + 1) % 5]) & bc[(i + 2) % 5];
        }
        s[0] ^= KRC[round];
    }
}
static void keccakf_bytes(uint8_t st[200]) {
    uint64_t s[25];
    for (int i = 0; i < 25; i++) {
        s[i] = 0;
        for (int b = 0; b < 8; b++) s[i] |= (uint64_t)st[8 * i + b] << (8 * b);
    }
    keccakf(s);
    for (int i = 0; i < 25; i++)
        for (int b = 0; b < 8; b++) st[8 * i + b] = (uint8_t)(s[i] >> (8 * b));
}
static void sponge(const uint8_t* in, size_t len, size_t rate, uint8_t dom, uint8_t* out, size_t outlen) {
    uint8_t st[200];
    memset(st, 0, 200);
    while (len >= rate) {
        for (size_t i = 0; i < rate; i++) st[i] ^= in[i];
        keccakf_bytes(st);
        in += rate;
        len -= rate;
    }
    for (size_t i = 0; i < len; i++) st[i] ^= in[i];
    st[len] ^= dom;
    st[rate - 1] ^= 0x80;
    keccakf_bytes(st);
    while (outlen > 0) {
        size_t k = outlen < rate ? outlen : rate;
        memcpy(out, st, k);
        out += k;
        outlen -= k;
        if (outlen) keccakf_bytes(st);
    }
}
void orc_sha3_256(const uint8_t* in, size_t len, uint8_t out[32]) { sponge(in, len, 136, 0x06, out, 32); }
void orc_shake256(const uint8_t* in, size_t len, uint8_t* out, size_t outlen) { sponge(in, len, 136, 0x1f, out, outlen); }

/* ------------------------------------------------------------------ generators */
static int scalar_from_le32(const uint8_t b[32], uint64_t k[4]) { /* scalar.rs:87-95 */
    for (int i = 0; i < 4; i++) {
        k[i] = 0;
        for (int j = 0; j < 8; j++) k[i] |= (uint64_t)b[8 * i + j] << (8 * j);
    }
    return !geq(k, FIELDS[FR].p);
}
static uint8_t gen_scalar_from_chunk(const uint8_t chunk[64], uint64_t k[4]) { /* group.rs:110-132 */
    uint8_t h[32], buf[72];
    orc_sha3_256(chunk, 64, h);
    if (scalar_from_le32(h, k)) return 0;
    memcpy(buf, "fallback", 8);
    memcpy(buf + 8, chunk, 64);
    orc_sha3_256(buf, 72, h);
    if (scalar_from_le32(h, k)) return 1;
    k[0] = 1; k[1] = k[2] = k[3] = 0;
    return 2;
}
void orc_gen_scalars(const uint8_t* label, size_t label_len, size_t n, ofp* out, uint8_t* kinds) {
    og1a g;
    orc_g1_generator(&g);
    uint8_t* seed = (uint8_t*)malloc(label_len + 32);
    memcpy(seed, label, label_len);
    orc_g1_compress(&g, 0, seed + label_len);
    uint8_t* xof = (uint8_t*)malloc(64 * (n + 1));
    orc_shake256(seed, label_len + 32, xof, 64 * (n + 1));
    for (size_t i = 0; i <= n; i++) {
        uint64_t k[4];
        uint8_t kind = gen_scalar_from_chunk(xof + 64 * i, k);
        fp_from_canon(FR, k, &out[i]);
        if (kinds) kinds[i] = kind;
    }
    free(xof);
    free(seed);
}
typedef struct { const ofp* sc; og1a* out; og1a g; } gens_ctx;
static void gens_body(long i, void* vc) {
    gens_ctx* c = (gens_ctx*)vc;
    uint8_t inf;
    orc_g1_scalar_mul(&c->g, 0, &c->sc[i], &c->out[i], &inf);
}
void orc_multi_commit_gens(const uint8_t* label, size_t label_len, size_t n, og1a* out) {
    ofp* sc = (ofp*)malloc(sizeof(ofp) * (n + 1));
    orc_gen_scalars(label, label_len, n, sc, NULL);
    gens_ctx c = {sc, out, {{{0}}, {{0}}}};
    orc_g1_generator(&c.g);
    parallel_for((long)n + 1, 16, 0, gens_body, &c);
    free(sc);
}

/* ------------------------------------------------------------------ Hyrax */
static void commit_row(const og1a* pts /* R_size+1 incl. h */, const ofp* row, size_t R_size, const ofp* blind,
                       ofp* scratch, og1a* out, uint8_t* inf) { /* commitments.rs:144-154 */
    memcpy(scratch, row, sizeof(ofp) * R_size);
    scratch[R_size] = *blind;
    orc_msm(pts, NULL, scratch, R_size + 1, 1, out, inf);
}
typedef struct { const og1a* pts; const ofp* Z; const ofp* blinds; size_t R_size; og1a* C; uint8_t* inf; } commit_ctx;
static void commit_body(long i, void* vc) {
    commit_ctx* c = (commit_ctx*)vc;
    ofp zero;
    memset(&zero, 0, sizeof zero);
    ofp* scratch = (ofp*)malloc(sizeof(ofp) * (c->R_size + 1));
    commit_row(c->pts, c->Z + (size_t)i * c->R_size, c->R_size, c->blinds ? &c->blinds[i] : &zero, scratch,
               &c->C[i], &c->inf[i]);
    free(scratch);
}
void orc_hyrax_commit(const og1a* G, const og1a* h, const ofp* Z, size_t L_size, size_t R_size,
                      const ofp* blinds, int threads, og1a* C_out, uint8_t* inf_out) {
    og1a* pts = (og1a*)malloc(sizeof(og1a) * (R_size + 1));
    memcpy(pts, G, sizeof(og1a) * R_size);
    pts[R_size] = *h;
    commit_ctx c = {pts, Z, blinds, R_size, C_out, inf_out};
    parallel_for((long)L_size, 1, threads, commit_body, &c);   /* rows are independent: hyrax.rs:259 */
    free(pts);
}

typedef struct { const ofp* Z; const ofp* L; size_t L_size, R_size; ofp* LZ; } bound_ctx;
static void bound_body(long i, void* vc) {
    bound_ctx* c = (bound_ctx*)vc;
    ofp acc, t;
    memset(&acc, 0, sizeof acc);
    for (size_t j = 0; j < c->L_size; j++) {
        fp_mul(FR, &c->L[j], &c->Z[j * c->R_size + (size_t)i], &t);
        fp_add(FR, &acc, &t, &acc);
    }
    c->LZ[i] = acc;
}
void orc_bound(const ofp* Z, const ofp* L, size_t L_size, size_t R_size, int threads, ofp* LZ) {
    bound_ctx c = {Z, L, L_size, R_size, LZ};
    parallel_for((long)R_size, 16, threads, bound_body, &c);
}

void orc_eq_evals(const ofp* r, size_t ell, ofp* ev) {
    size_t n = (size_t)1 << ell;
    for (size_t i = 0; i < n; i++) fp_one(FR, &ev[i]);
    size_t size = 1;
    for (size_t j = 0; j < ell; j++) {
        size *= 2;
        for (size_t i = size; i-- > 0;) {
            if ((i & 1) == 0) continue;
            ofp s = ev[i / 2];
            fp_mul(FR, &s, &r[j], &ev[i]);
            fp_sub(FR, &s, &ev[i], &ev[i - 1]);
        }
    }
}

void orc_bind_top(ofp* Z, size_t len, const ofp* r) {
    size_t n = len / 2;
    for (size_t i = 0; i < n; i++) {
        ofp d;
        fp_sub(FR, &Z[i + n], &Z[i], &d);
        fp_mul(FR, r, &d, &d);
        fp_add(FR, &Z[i], &d, &Z[i]);
    }
}

void orc_sumcheck_cubic_eval(const ofp* A, const ofp* B, const ofp* C, const ofp* D, size_t len2,
                             ofp* e0, ofp* e2, ofp* e3) {
    /* tables: A = tau, B = Az, C = Bz, D = Cz; comb = tau * (Az*Bz - Cz)  (r1csproof.rs comb_func) */
    size_t len = len2 / 2;
    ofp s0, s2, s3;
    memset(&s0, 0, sizeof s0); memset(&s2, 0, sizeof s2); memset(&s3, 0, sizeof s3);
    const ofp* T[4] = {A, B, C, D};
    for (size_t i = 0; i < len; i++) {
        ofp v0[4], v2[4], v3[4], t;
        for (int k = 0; k < 4; k++) {
            v0[k] = T[k][i];
            fp_add(FR, &T[k][len + i], &T[k][len + i], &v2[k]);
            fp_sub(FR, &v2[k], &T[k][i], &v2[k]);
            fp_add(FR, &v2[k], &T[k][len + i], &v3[k]);
            fp_sub(FR, &v3[k], &T[k][i], &v3[k]);
        }
        ofp (*vs[3])[4] = {&v0, &v2, &v3};
        ofp* acc[3] = {&s0, &s2, &s3};
        for (int e = 0; e < 3; e++) {
            ofp* v = *vs[e];
            fp_mul(FR, &v[1], &v[2], &t);
            fp_sub(FR, &t, &v[3], &t);
            fp_mul(FR, &v[0], &t, &t);
            fp_add(FR, acc[e], &t, acc[e]);
        }
    }
    *e0 = s0; *e2 = s2; *e3 = s3;
}

void orc_sumcheck_quad_eval(const ofp* Zt, const ofp* ABC, size_t len2, ofp* e0, ofp* e2) {
    /* sumcheck.rs:690-699 with comb = z * ABC (r1csproof.rs phase 2) */
    size_t len = len2 / 2;
    ofp s0, s2, t, a, b;
    memset(&s0, 0, sizeof s0); memset(&s2, 0, sizeof s2);
    for (size_t i = 0; i < len; i++) {
        fp_mul(FR, &Zt[i], &ABC[i], &t); fp_add(FR, &s0, &t, &s0);
        fp_add(FR, &Zt[len + i], &Zt[len + i], &a); fp_sub(FR, &a, &Zt[i], &a);
        fp_add(FR, &ABC[len + i], &ABC[len + i], &b); fp_sub(FR, &b, &ABC[i], &b);
        fp_mul(FR, &a, &b, &t); fp_add(FR, &s2, &t, &s2);
    }
    *e0 = s0; *e2 = s2;
}

/* ------------------------------------------------------------------ bullet reduction */
static void fr_dot(const ofp* a, const ofp* b, size_t n, ofp* out) {
    ofp acc, t;
    memset(&acc, 0, sizeof acc);
    for (size_t i = 0; i < n; i++) { fp_mul(FR, &a[i], &b[i], &t); fp_add(FR, &acc, &t, &acc); }
    *out = acc;
}
static void j_msm_affine_pts(const og1a* pts, const uint8_t* inf, const ofp* sc, size_t n, og1j* out) {
    if (n >= 16) msm_pippenger(pts, inf, sc, n, out);
    else msm_naive(pts, inf, sc, n, out);
}
static void j_mul_fr(const og1a* p, uint8_t inf, const ofp* s, og1j* out) {
    uint64_t k[4];
    fp_to_canon(FR, s, k);
    j_scalar_mul_canon(p, inf, k, out);
}

typedef struct { og1a* G; uint8_t* Ginf; size_t n; const ofp* u; const ofp* ui; } fold_ctx;
static void fold_body(long i, void* vc) {
    fold_ctx* c = (fold_ctx*)vc;
    og1j gl, gr;
    j_mul_fr(&c->G[i], c->Ginf[i], c->ui, &gl);
    j_mul_fr(&c->G[c->n + i], c->Ginf[c->n + i], c->u, &gr);
    j_add(&gl, &gr, &gl);
    j_to_affine(&gl, &c->G[i], &c->Ginf[i]);
}
void orc_bullet_prove(const og1a* Q, const og1a* G_in, size_t n, const og1a* H, const ofp* a_in, const ofp* b_in,
                      const ofp* blind, const ofp* blinds_L, const ofp* blinds_R, const ofp* u_vec,
                      og1a* L_out, uint8_t* L_inf, og1a* R_out, uint8_t* R_inf,
                      og1a* Gamma, uint8_t* Gamma_inf, ofp* a_hat, ofp* b_hat,
                      og1a* g_hat, uint8_t* g_hat_inf, ofp* blind_hat) {
    og1a* G = (og1a*)malloc(sizeof(og1a) * n);
    uint8_t* Ginf = (uint8_t*)calloc(n, 1);
    ofp* a = (ofp*)malloc(sizeof(ofp) * n);
    ofp* b = (ofp*)malloc(sizeof(ofp) * n);
    memcpy(G, G_in, sizeof(og1a) * n);
    memcpy(a, a_in, sizeof(ofp) * n);
    memcpy(b, b_in, sizeof(ofp) * n);

    og1j acc, t;
    ofp c;
    /* bullet.rs:57-59 */
    j_msm_affine_pts(G, Ginf, a, n, &acc);
    fr_dot(a, b, n, &c);
    j_mul_fr(Q, 0, &c, &t); j_add(&acc, &t, &acc);
    j_mul_fr(H, 0, blind, &t); j_add(&acc, &t, &acc);
    j_to_affine(&acc, Gamma, Gamma_inf);

    ofp blind_Gamma = *blind;
    size_t round = 0;
    while (n > 1) {
        n /= 2;
        ofp cL, cR;
        fr_dot(a, b + n, n, &cL);           /* <a_L, b_R> */
        fr_dot(a + n, b, n, &cR);           /* <a_R, b_L> */
        /* L = MSM(a_L, G_R) + c_L Q + blind_L H ; R = MSM(a_R, G_L) + c_R Q + blind_R H */
        j_msm_affine_pts(G + n, Ginf + n, a, n, &acc);
        j_mul_fr(Q, 0, &cL, &t); j_add(&acc, &t, &acc);
        j_mul_fr(H, 0, &blinds_L[round], &t); j_add(&acc, &t, &acc);
        j_to_affine(&acc, &L_out[round], &L_inf[round]);
        j_msm_affine_pts(G, Ginf, a + n, n, &acc);
        j_mul_fr(Q, 0, &cR, &t); j_add(&acc, &t, &acc);
        j_mul_fr(H, 0, &blinds_R[round], &t); j_add(&acc, &t, &acc);
        j_to_affine(&acc, &R_out[round], &R_inf[round]);

        ofp u = u_vec[round], ui, uu, uiui, t1, t2;
        fp_inv(FR, &u, &ui);
        fold_ctx fc = {G, Ginf, n, &u, &ui};
        parallel_for((long)n, 8, 0, fold_body, &fc);   /* bullet.rs:85-89 */
        for (size_t i = 0; i < n; i++) {       /* bullet.rs:92-102 */
            fp_mul(FR, &u, &a[i], &t1); fp_mul(FR, &ui, &a[n + i], &t2); fp_add(FR, &t1, &t2, &a[i]);
            fp_mul(FR, &ui, &b[i], &t1); fp_mul(FR, &u, &b[n + i], &t2); fp_add(FR, &t1, &t2, &b[i]);
        }
        fp_mul(FR, &u, &u, &uu);
        fp_mul(FR, &ui, &ui, &uiui);
        fp_mul(FR, &uu, &blinds_L[round], &t1);
        fp_mul(FR, &uiui, &blinds_R[round], &t2);
        fp_add(FR, &blind_Gamma, &t1, &blind_Gamma);
        fp_add(FR, &blind_Gamma, &t2, &blind_Gamma);
        round++;
    }
    *a_hat = a[0];
    *b_hat = b[0];
    *g_hat = G[0];
    *g_hat_inf = Ginf[0];
    *blind_hat = blind_Gamma;
    free(G); free(Ginf); free(a); free(b);
}

/* ------------------------------------------------------------------ Merlin (STROBE-128) */
#define STROBE_R 166
enum { FLAG_I = 1, FLAG_A = 2, FLAG_C = 4, FLAG_T = 8, FLAG_M = 16, FLAG_K = 32 };
static void strobe_run_f(orc_transcript* t) {
    t->st[t->pos] ^= t->pos_begin;
    t->st[t->pos + 1] ^= 0x04;
    t->st[STROBE_R + 1] ^= 0x80;
    keccakf_bytes(t->st);
    t->pos = 0;
    t->pos_begin = 0;
}
static void strobe_absorb(orc_transcript* t, const uint8_t* d, size_t n) {
    for (size_t i = 0; i < n; i++) {
        t->st[t->pos] ^= d[i];
        t->pos++;
        if (t->pos == STROBE_R) strobe_run_f(t);
    }
}
static void strobe_squeeze(orc_transcript* t, uint8_t* d, size_t n) {
    for (size_t i = 0; i < n; i++) {
        d[i] = t->st[t->pos];
        t->st[t->pos] = 0;
        t->pos++;
        if (t->pos == STROBE_R) strobe_run_f(t);
    }
}
static void strobe_begin_op(orc_transcript* t, uint8_t flags, int more) {
    if (more) return;
    uint8_t old_begin = t->pos_begin;
    t->pos_begin = t->pos + 1;
    t->cur_flags = flags;
    uint8_t hdr[2] = {old_begin, flags};
    strobe_absorb(t, hdr, 2);
    if ((flags & (FLAG_C | FLAG_K)) && t->pos != 0) strobe_run_f(t);
}
static void strobe_meta_ad(orc_transcript* t, const uint8_t* d, size_t n, int more) {
    strobe_begin_op(t, FLAG_M | FLAG_A, more);
    strobe_absorb(t, d, n);
}
static void strobe_ad(orc_transcript* t, const uint8_t* d, size_t n, int more) {
    strobe_begin_op(t, FLAG_A, more);
    strobe_absorb(t, d, n);
}
static void strobe_prf(orc_transcript* t, uint8_t* d, size_t n, int more) {
    strobe_begin_op(t, FLAG_I | FLAG_A | FLAG_C, more);
    strobe_squeeze(t, d, n);
}
static void strobe_new(orc_transcript* t, const uint8_t* proto, size_t n) {
    memset(t, 0, sizeof *t);
    const uint8_t hdr[6] = {1, STROBE_R + 2, 1, 0, 1, 96};
    memcpy(t->st, hdr, 6);
    memcpy(t->st + 6, "STROBEv1.0.2", 12);
    keccakf_bytes(t->st);
    strobe_meta_ad(t, proto, n, 0);
}
void orc_transcript_append(orc_transcript* t, const uint8_t* label, size_t llen, const uint8_t* msg, size_t mlen) {
    uint8_t len4[4] = {(uint8_t)mlen, (uint8_t)(mlen >> 8), (uint8_t)(mlen >> 16), (uint8_t)(mlen >> 24)};
    strobe_meta_ad(t, label, llen, 0);
    strobe_meta_ad(t, len4, 4, 1);
    strobe_ad(t, msg, mlen, 0);
}
void orc_transcript_new(orc_transcript* t, const uint8_t* label, size_t len) {
    strobe_new(t, (const uint8_t*)"Merlin v1.0", 11);
    orc_transcript_append(t, (const uint8_t*)"dom-sep", 7, label, len);
}
void orc_transcript_challenge(orc_transcript* t, const uint8_t* label, size_t llen, uint8_t* out, size_t outlen) {
    uint8_t len4[4] = {(uint8_t)outlen, (uint8_t)(outlen >> 8), (uint8_t)(outlen >> 16), (uint8_t)(outlen >> 24)};
    strobe_meta_ad(t, label, llen, 0);
    strobe_meta_ad(t, len4, 4, 1);
    strobe_prf(t, out, outlen, 0);
}
void orc_transcript_challenge_scalar(orc_transcript* t, const uint8_t* label, size_t llen, ofp* out) {
    /* transcript.rs:56-67: 64 B, Fr::from_le_bytes_mod_order  => (lo + hi * 2^256) mod r */
    uint8_t buf[64];
    orc_transcript_challenge(t, label, llen, buf, 64);
    uint64_t lo[4], hi[4];
    for (int i = 0; i < 4; i++) {
        lo[i] = hi[i] = 0;
        for (int j = 0; j < 8; j++) {
            lo[i] |= (uint64_t)buf[8 * i + j] << (8 * j);
            hi[i] |= (uint64_t)buf[32 + 8 * i + j] << (8 * j);
        }
    }
    /* Montgomery: mont(lo) = lo*R2*R^-1 = lo*R ; hi*2^256 -> mont = hi*R*R = mul(mul(hi,R2),R2) */
    ofp l, h, r2, a, b;
    memcpy(l.l, lo, 32); memcpy(h.l, hi, 32); memcpy(r2.l, FIELDS[FR].r2, 32);
    fp_mul(FR, &l, &r2, &a);       /* lo (unreduced < 2^256 is fine for CIOS with a final subtract? see below) */
    fp_mul(FR, &h, &r2, &b);
    fp_mul(FR, &b, &r2, &b);
    fp_add(FR, &a, &b, out);
}

/* ------------------------------------------------------------------ Hyrax opening proof with transcripts */
static void t_append(orc_transcript* t, const char* label, const uint8_t* msg, size_t n) {
    orc_transcript_append(t, (const uint8_t*)label, strlen(label), msg, n);
}
static void t_protocol(orc_transcript* t, const char* name) { t_append(t, "protocol-name", (const uint8_t*)name, strlen(name)); } /* transcript.rs:38-40 */
static void t_scalar(orc_transcript* t, const char* label, const ofp* s) { /* transcript.rs:42-44 + scalar.rs:75-84 */
    uint64_t c[4];
    uint8_t b[32];
    fp_to_canon(FR, s, c);
    for (int i = 0; i < 4; i++) for (int j = 0; j < 8; j++) b[8 * i + j] = (uint8_t)(c[i] >> (8 * j));
    t_append(t, label, b, 32);
}
static void t_point(orc_transcript* t, const char* label, const og1a* p, uint8_t inf) { /* transcript.rs:102-108 */
    uint8_t b[32];
    orc_g1_compress(p, inf, b);
    t_append(t, label, b, 32);
}
static void t_challenge(orc_transcript* t, const char* label, ofp* out) {
    orc_transcript_challenge_scalar(t, (const uint8_t*)label, strlen(label), out);
}
static void j_commit2(const og1a* g, const ofp* s, const og1a* h, const ofp* blind, og1j* out) { /* s*g + blind*h */
    og1j t;
    j_mul_fr(g, 0, s, out);
    j_mul_fr(h, 0, blind, &t);
    j_add(out, &t, out);
}

void orc_poly_eval_prove(const ofp* Z, size_t ell, const ofp* blinds, const ofp* r, const ofp* Zr, const ofp* blind_Zr,
                         const og1a* G, const og1a* h, const og1a* G1, orc_transcript* tr, orc_transcript* tape,
                         orc_eval_proof* proof, og1a* C_Zr_prime, uint8_t* C_Zr_prime_inf) {
    t_protocol(tr, "polynomial evaluation proof");                       /* hyrax.rs:75 */
    const size_t lv = ell / 2, rv = ell - lv, L_size = (size_t)1 << lv, n = (size_t)1 << rv;
    ofp* Lv = (ofp*)malloc(sizeof(ofp) * L_size);
    ofp* Rv = (ofp*)malloc(sizeof(ofp) * n);
    orc_eq_evals(r, lv, Lv);
    orc_eq_evals(r + lv, rv, Rv);
    ofp* LZ = (ofp*)malloc(sizeof(ofp) * n);
    orc_bound(Z, Lv, L_size, n, 0, LZ);                                   /* hyrax.rs:100 */
    ofp LZ_blind, zero, t1, t2;
    memset(&zero, 0, sizeof zero);
    LZ_blind = zero;
    if (blinds) for (size_t i = 0; i < L_size; i++) { fp_mul(FR, &blinds[i], &Lv[i], &t1); fp_add(FR, &LZ_blind, &t1, &LZ_blind); }
    const ofp* blind_y = blind_Zr ? blind_Zr : &zero;

    /* DotProductProofLog::prove (nizk/mod.rs:439-522) */
    t_protocol(tr, "dot product proof (log)");
    size_t lg = 0;
    while (((size_t)1 << lg) < n) lg++;
    ofp d, r_delta, r_beta, bl1[ORC_MAX_LG], bl2[ORC_MAX_LG];
    t_challenge(tape, "d", &d);
    t_challenge(tape, "r_delta", &r_delta);
    t_challenge(tape, "r_delta", &r_beta);                                 /* sic: same label, mod.rs:459 */
    for (size_t i = 0; i < lg; i++) t_challenge(tape, "blinds_vec_1", &bl1[i]);
    for (size_t i = 0; i < lg; i++) t_challenge(tape, "blinds_vec_2", &bl2[i]);

    og1j acc, tj;
    og1a Cx, Cy;
    uint8_t Cx_inf, Cy_inf;
    j_msm_affine_pts(G, NULL, LZ, n, &acc);                               /* Cx = x.commit(blind_x, gens_n) */
    j_mul_fr(h, 0, &LZ_blind, &tj); j_add(&acc, &tj, &acc);
    j_to_affine(&acc, &Cx, &Cx_inf);
    t_point(tr, "Cx", &Cx, Cx_inf);
    j_commit2(G1, Zr, h, blind_y, &acc);                                   /* Cy = y.commit(blind_y, gens_1) */
    j_to_affine(&acc, &Cy, &Cy_inf);
    t_point(tr, "Cy", &Cy, Cy_inf);
    for (size_t i = 0; i < n; i++) t_scalar(tr, "a", &Rv[i]);             /* a_vec.append_to_transcript(b"a") */
    ofp rr;
    t_challenge(tr, "r", &rr);
    og1a Q; uint8_t Q_inf;                                                 /* gens_1.scale(r).G[0] */
    orc_g1_scalar_mul(G1, 0, &rr, &Q, &Q_inf);
    ofp blind_Gamma;
    fp_mul(FR, &rr, blind_y, &t1);
    fp_add(FR, &LZ_blind, &t1, &blind_Gamma);

    /* BulletReductionProof::prove (bullet.rs:24-126), challenges from the transcript */
    og1a* Gc = (og1a*)malloc(sizeof(og1a) * n);
    uint8_t* Ginf = (uint8_t*)calloc(n, 1);
    ofp* a = (ofp*)malloc(sizeof(ofp) * n);
    ofp* b = (ofp*)malloc(sizeof(ofp) * n);
    memcpy(Gc, G, sizeof(og1a) * n); memcpy(a, LZ, sizeof(ofp) * n); memcpy(b, Rv, sizeof(ofp) * n);
    proof->lg_n = lg;
    ofp rhat = blind_Gamma;
    size_t m = n;
    for (size_t round = 0; round < lg; round++) {
        m /= 2;
        ofp cL, cR;
        fr_dot(a, b + m, m, &cL);
        fr_dot(a + m, b, m, &cR);
        j_msm_affine_pts(Gc + m, Ginf + m, a, m, &acc);
        j_mul_fr(&Q, Q_inf, &cL, &tj); j_add(&acc, &tj, &acc);
        j_mul_fr(h, 0, &bl1[round], &tj); j_add(&acc, &tj, &acc);
        j_to_affine(&acc, &proof->L[round], &proof->L_inf[round]);
        j_msm_affine_pts(Gc, Ginf, a + m, m, &acc);
        j_mul_fr(&Q, Q_inf, &cR, &tj); j_add(&acc, &tj, &acc);
        j_mul_fr(h, 0, &bl2[round], &tj); j_add(&acc, &tj, &acc);
        j_to_affine(&acc, &proof->R[round], &proof->R_inf[round]);
        t_point(tr, "L", &proof->L[round], proof->L_inf[round]);
        t_point(tr, "R", &proof->R[round], proof->R_inf[round]);
        ofp u, ui, uu, uiui;
        t_challenge(tr, "u", &u);
        fp_inv(FR, &u, &ui);
        fold_ctx fc = {Gc, Ginf, m, &u, &ui};
        parallel_for((long)m, 8, 0, fold_body, &fc);
        for (size_t i = 0; i < m; i++) {
            fp_mul(FR, &u, &a[i], &t1); fp_mul(FR, &ui, &a[m + i], &t2); fp_add(FR, &t1, &t2, &a[i]);
            fp_mul(FR, &ui, &b[i], &t1); fp_mul(FR, &u, &b[m + i], &t2); fp_add(FR, &t1, &t2, &b[i]);
        }
        fp_mul(FR, &u, &u, &uu); fp_mul(FR, &ui, &ui, &uiui);
        fp_mul(FR, &uu, &bl1[round], &t1); fp_mul(FR, &uiui, &bl2[round], &t2);
        fp_add(FR, &rhat, &t1, &rhat); fp_add(FR, &rhat, &t2, &rhat);
    }
    ofp x_hat = a[0], a_hat = b[0], y_hat;
    og1a g_hat = Gc[0]; uint8_t g_hat_inf = Ginf[0];
    fp_mul(FR, &x_hat, &a_hat, &y_hat);
    /* delta = d.commit(r_delta, (g_hat, h)); beta = d.commit(r_beta, gens_1_scaled)   (mod.rs:497-505) */
    j_mul_fr(&g_hat, g_hat_inf, &d, &acc); j_mul_fr(h, 0, &r_delta, &tj); j_add(&acc, &tj, &acc);
    j_to_affine(&acc, &proof->delta, &proof->delta_inf);
    t_point(tr, "delta", &proof->delta, proof->delta_inf);
    j_mul_fr(&Q, Q_inf, &d, &acc); j_mul_fr(h, 0, &r_beta, &tj); j_add(&acc, &tj, &acc);
    j_to_affine(&acc, &proof->beta, &proof->beta_inf);
    t_point(tr, "beta", &proof->beta, proof->beta_inf);
    ofp c;
    t_challenge(tr, "c", &c);
    fp_mul(FR, &c, &y_hat, &t1); fp_add(FR, &d, &t1, &proof->z1);          /* z1 = d + c*y_hat */
    fp_mul(FR, &c, &rhat, &t1); fp_add(FR, &t1, &r_beta, &t1);            /* z2 = a_hat*(c*rhat + r_beta) + r_delta */
    fp_mul(FR, &a_hat, &t1, &t1); fp_add(FR, &t1, &r_delta, &proof->z2);
    *C_Zr_prime = Cy; *C_Zr_prime_inf = Cy_inf;
    free(Lv); free(Rv); free(LZ); free(Gc); free(Ginf); free(a); free(b);
}

int orc_poly_eval_verify(const orc_eval_proof* proof, size_t ell, const ofp* r, const og1a* C_Zr, uint8_t C_Zr_inf,
                         const og1a* comm, const uint8_t* comm_inf, const og1a* G, const og1a* h, const og1a* G1,
                         orc_transcript* tr) {
    t_protocol(tr, "polynomial evaluation proof");                       /* hyrax.rs:126 */
    const size_t lv = ell / 2, rv = ell - lv, L_size = (size_t)1 << lv, n = (size_t)1 << rv;
    ofp* Lv = (ofp*)malloc(sizeof(ofp) * L_size);
    ofp* Rv = (ofp*)malloc(sizeof(ofp) * n);
    orc_eq_evals(r, lv, Lv);
    orc_eq_evals(r + lv, rv, Rv);
    og1j C_LZ;                                                            /* hyrax.rs:133 */
    j_msm_affine_pts(comm, comm_inf, Lv, L_size, &C_LZ);
    og1a Cx; uint8_t Cx_inf;
    j_to_affine(&C_LZ, &Cx, &Cx_inf);
    /* DotProductProofLog::verify (mod.rs:525-567) */
    t_protocol(tr, "dot product proof (log)");
    t_point(tr, "Cx", &Cx, Cx_inf);
    t_point(tr, "Cy", C_Zr, C_Zr_inf);
    for (size_t i = 0; i < n; i++) t_scalar(tr, "a", &Rv[i]);
    ofp rr;
    t_challenge(tr, "r", &rr);
    og1a Q; uint8_t Q_inf;
    orc_g1_scalar_mul(G1, 0, &rr, &Q, &Q_inf);
    og1j Gamma, tj;                                                       /* Gamma = Cx + r*Cy */
    j_mul_fr(C_Zr, C_Zr_inf, &rr, &Gamma);
    j_add(&Gamma, &C_LZ, &Gamma);
    /* BulletReductionProof::verify (bullet.rs:130-173) */
    const size_t lg = proof->lg_n;
    int ok = (((size_t)1 << lg) == n);
    ofp u[ORC_MAX_LG], ui[ORC_MAX_LG], usq[ORC_MAX_LG], uisq[ORC_MAX_LG];
    for (size_t i = 0; i < lg && ok; i++) {
        t_point(tr, "L", &proof->L[i], proof->L_inf[i]);
        t_point(tr, "R", &proof->R[i], proof->R_inf[i]);
        t_challenge(tr, "u", &u[i]);
        fp_inv(FR, &u[i], &ui[i]);
        fp_mul(FR, &u[i], &u[i], &usq[i]);
        fp_inv(FR, &usq[i], &uisq[i]);
    }
    ofp* s = (ofp*)malloc(sizeof(ofp) * n);
    for (size_t i = 0; i < n && ok; i++) {                                /* compute_s */
        fp_one(FR, &s[i]);
        for (size_t j = 0; j < lg; j++) fp_mul(FR, &s[i], ((i >> j) & 1) ? &u[lg - 1 - j] : &ui[lg - 1 - j], &s[i]);
    }
    int result = 0;
    if (ok) {
        og1j g_hat, Gamma_hat, lhs, rhs, t2;
        ofp a_hat, t1;
        j_msm_affine_pts(G, NULL, s, n, &g_hat);
        fr_dot(s, Rv, n, &a_hat);
        j_msm_affine_pts(proof->L, proof->L_inf, usq, lg, &Gamma_hat);
        j_msm_affine_pts(proof->R, proof->R_inf, uisq, lg, &tj);
        j_add(&Gamma_hat, &Gamma, &Gamma_hat);
        j_add(&Gamma_hat, &tj, &Gamma_hat);
        t_point(tr, "delta", &proof->delta, proof->delta_inf);
        t_point(tr, "beta", &proof->beta, proof->beta_inf);
        ofp c;
        t_challenge(tr, "c", &c);
        /* lhs = (c*Gamma_hat + beta)*a_hat + delta ; rhs = z1*(g_hat + a_hat*Q) + z2*h */
        og1a tmp; uint8_t tmp_inf; og1j bj, dj;
        j_to_affine(&Gamma_hat, &tmp, &tmp_inf);
        j_mul_fr(&tmp, tmp_inf, &c, &lhs);
        j_from_affine(&proof->beta, proof->beta_inf, &bj);
        j_add(&lhs, &bj, &lhs);
        j_to_affine(&lhs, &tmp, &tmp_inf);
        j_mul_fr(&tmp, tmp_inf, &a_hat, &lhs);
        j_from_affine(&proof->delta, proof->delta_inf, &dj);
        j_add(&lhs, &dj, &lhs);
        j_mul_fr(&Q, Q_inf, &a_hat, &t2);
        j_add(&t2, &g_hat, &t2);
        j_to_affine(&t2, &tmp, &tmp_inf);
        j_mul_fr(&tmp, tmp_inf, &proof->z1, &rhs);
        j_mul_fr(h, 0, &proof->z2, &t2);
        j_add(&rhs, &t2, &rhs);
        og1a la, ra; uint8_t li, ri;
        j_to_affine(&lhs, &la, &li);
        j_to_affine(&rhs, &ra, &ri);
        result = (li == ri) && (li || (fp_eq(&la.x, &ra.x) && fp_eq(&la.y, &ra.y)));
        (void)t1;
    }
    free(Lv); free(Rv); free(s);
    return result;
}

/* ==================================================================================================================
 * CPU BASELINE of the end-to-end prove (bench.py's cpu_baseline leg for the second half of BASELINE.json's metric).
 *
 * orc_prove_workload() runs, on the host threads, the table-sized phases of SNARK::prove (snark.rs:428-484) for a synthetic
 * R1CS of the keyless shape -- n = 2^s constraints and variables, 3 + 2 + 1 non-zeros per row so that every matrix pads to
 * 4 n entries (the shape bench.py's GPU prove uses) -- in the reference's order, with the reference's algorithms (file:line
 * at every phase), the real field and group arithmetic of this file, and every round's challenge drawn from a Merlin
 * transcript fed with that round's evaluations, so the rounds are as sequential as the prover's.  It is a RESTATEMENT OF THE
 * WORK, not a prover: tables are seeded pseudo-random values (a satisfying assignment changes no operation count), the
 * Sigma-protocols of the ZK sumchecks are represented by their commitments (four short MSMs per round), and nothing is
 * assembled into a proof or verified -- tests/test_snark.py does that for the GPU prover.  The reference itself runs these
 * phases with Rayon over rows / table entries; threads here play that role.
 * ================================================================================================================== */
static double wall_now(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
static inline uint64_t wl_mix(uint64_t z) {
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static inline void wl_scalar(uint64_t seed, uint64_t i, ofp* out) {      /* a valid Montgomery representative (< 2^253 < r) */
    for (int k = 0; k < 4; k++) out->l[k] = wl_mix(seed * 0x100000001B3ULL + 4 * i + (uint64_t)k);
    out->l[3] &= (1ULL << 61) - 1;
}
typedef struct { ofp* v; uint64_t seed; } wl_fill_ctx;
static void wl_fill_body(long i, void* vc) { wl_fill_ctx* c = (wl_fill_ctx*)vc; wl_scalar(c->seed, (uint64_t)i, &c->v[i]); }
static ofp* wl_table(size_t n, uint64_t seed, int threads) {
    ofp* v = (ofp*)malloc(sizeof(ofp) * n);
    wl_fill_ctx c = {v, seed};
    parallel_for((long)n, 4096, threads, wl_fill_body, &c);
    return v;
}

/* --- threaded forms of the table passes (the serial ones above are the oracle's; these split the index range) */
#define WL_MAX_THREADS 256
typedef struct { ofp** T; int nt; size_t half; const ofp* r; } wl_bind_ctx;
static void wl_bind_body(long i, void* vc) {                             /* bound_poly_var_top, hyrax.rs:195-203 */
    wl_bind_ctx* c = (wl_bind_ctx*)vc;
    for (int k = 0; k < c->nt; k++) {
        ofp d;
        fp_sub(FR, &c->T[k][(size_t)i + c->half], &c->T[k][i], &d);
        fp_mul(FR, c->r, &d, &d);
        fp_add(FR, &c->T[k][i], &d, &c->T[k][i]);
    }
}
typedef struct { const ofp* A; const ofp* B; const ofp* C; const ofp* D; size_t half, per; ofp (*part)[3]; int mode; } wl_eval_ctx;
static void wl_eval_body(long blk, void* vc) {
    /* mode 0: tau * (Az * Bz - Cz) at 0, 2, 3 (sumcheck.rs:501-530); mode 1: z * ABC at 0, 2 (sumcheck.rs:690-699);
     * mode 2: A * B * eq at 0, 2, 3 (product layer, sumcheck.rs:236-262) */
    wl_eval_ctx* c = (wl_eval_ctx*)vc;
    size_t lo = (size_t)blk * c->per, hi = lo + c->per;
    if (hi > c->half) hi = c->half;
    ofp s[3], t, u;
    memset(s, 0, sizeof s);
    const ofp* T[4] = {c->A, c->B, c->C, c->D};
    const int nt = c->mode == 0 ? 4 : (c->mode == 1 ? 2 : 3);
    for (size_t i = lo; i < hi; i++) {
        ofp v[3][4];
        for (int k = 0; k < nt; k++) {
            const ofp *a = &T[k][i], *b = &T[k][c->half + i];
            v[0][k] = *a;
            fp_add(FR, b, b, &v[1][k]); fp_sub(FR, &v[1][k], a, &v[1][k]);           /* 2b - a */
            fp_add(FR, &v[1][k], b, &v[2][k]); fp_sub(FR, &v[2][k], a, &v[2][k]);    /* 3b - 2a */
        }
        for (int e = 0; e < 3; e++) {
            if (c->mode == 1 && e == 2) break;
            if (c->mode == 0) { fp_mul(FR, &v[e][1], &v[e][2], &t); fp_sub(FR, &t, &v[e][3], &t); fp_mul(FR, &v[e][0], &t, &u); }
            else if (c->mode == 1) fp_mul(FR, &v[e][0], &v[e][1], &u);
            else { fp_mul(FR, &v[e][0], &v[e][1], &t); fp_mul(FR, &t, &v[e][2], &u); }
            fp_add(FR, &s[e], &u, &s[e]);
        }
    }
    memcpy(c->part[blk], s, sizeof s);
}
/* one sumcheck: `rounds` rounds of evaluate + challenge + bind over nt tables of length len; ncommit short commitments per
 * round stand for comm_poly and the per-round dot-product proof of the ZK variants (sumcheck.rs:533-649) */
static void wl_sumcheck(ofp** T, int nt, size_t len, int mode, int rounds, int ncommit, const og1a* G4, orc_transcript* tr, int threads_in) {
    const int nth = host_threads(threads_in);
    ofp (*part)[3] = (ofp(*)[3])malloc(sizeof(ofp[3]) * (size_t)nth * 4);
    for (int j = 0; j < rounds && len > 1; j++) {
        const size_t half = len / 2;
        const int threads = half >= 8192 ? threads_in : 1;        /* a thread team costs more than a short round (Rayon's pool would not split it either) */
        long nblk = threads == 1 ? 1 : (long)nth * 4;
        if ((size_t)nblk > half) nblk = (long)half;
        wl_eval_ctx ec = {T[0], T[1], nt > 2 ? T[2] : NULL, nt > 3 ? T[3] : NULL, half, (half + (size_t)nblk - 1) / (size_t)nblk, part, mode};
        parallel_for(nblk, 1, threads, wl_eval_body, &ec);
        ofp e[3];
        memset(e, 0, sizeof e);
        for (long b = 0; b < nblk; b++) for (int k = 0; k < 3; k++) fp_add(FR, &e[k], &part[b][k], &e[k]);
        for (int k = 0; k < ncommit; k++) {                   /* 4-point commitments (comm_poly, delta, beta, ...) */
            og1j acc;
            ofp sc[4] = {e[0], e[1], e[2], e[(k + j) % 3]};
            msm_naive(G4, NULL, sc, 4, &acc);
            og1a p; uint8_t inf;
            j_to_affine(&acc, &p, &inf);
            t_point(tr, "comm_poly", &p, inf);
        }
        for (int k = 0; k < 3; k++) t_scalar(tr, "poly", &e[k]);
        ofp r;
        t_challenge(tr, "challenge_nextround", &r);
        wl_bind_ctx bc = {T, nt, half, &r};
        parallel_for((long)half, 2048, threads, wl_bind_body, &bc);
        len = half;
    }
    free(part);
}
typedef struct { ofp** circ; const size_t* clen; const ofp* eq; size_t len; int layer; const og1a* G4; const orc_transcript* tr; } wl_layer_ctx;
static void wl_layer_body(long k, void* vc) {                             /* one circuit's share of a layer's batched sumcheck */
    wl_layer_ctx* c = (wl_layer_ctx*)vc;
    const size_t len = c->len;
    if (c->clen[k] < 2 * len) return;
    size_t off = 0, l2 = c->clen[k];
    while (l2 > 2 * len) { off += l2; l2 /= 2; }             /* the stored layer with 2 len entries: left half, right half */
    ofp* tabs[3];
    for (int t = 0; t < 3; t++) tabs[t] = (ofp*)malloc(sizeof(ofp) * len);
    memcpy(tabs[0], c->circ[k] + off, sizeof(ofp) * len);
    memcpy(tabs[1], c->circ[k] + off + len, sizeof(ofp) * len);
    memcpy(tabs[2], c->eq, sizeof(ofp) * len);
    orc_transcript local = *c->tr;
    wl_sumcheck(tabs, 3, len, 2, c->layer, 0, c->G4, &local, 1);
    for (int t = 0; t < 3; t++) free(tabs[t]);
}
typedef struct { const ofp* eq_r; const ofp* eq_c; const ofp* z; ofp* out; size_t n; int nnz_per_row; uint64_t seed; int mode; } wl_sp_ctx;
static void wl_sparse_body(long i, void* vc) {
    /* mode 0: (M z)[i] = sum val * z[col] (r1cs.rs:275-288); mode 1: sum val * eq_rx[row] * eq_ry[col] per row
     * (R1CSInstance::evaluate, sparse_mlpoly.rs), the row's partial sum stored in out[i] */
    wl_sp_ctx* c = (wl_sp_ctx*)vc;
    ofp acc, val, t;
    memset(&acc, 0, sizeof acc);
    for (int k = 0; k < c->nnz_per_row; k++) {
        const uint64_t h = wl_mix(c->seed + (uint64_t)i * 8 + (uint64_t)k);
        const size_t col = (size_t)(h % c->n);
        wl_scalar(c->seed ^ 0x5555, (uint64_t)i * 8 + (uint64_t)k, &val);
        if (c->mode == 0) fp_mul(FR, &val, &c->z[col], &t);
        else { fp_mul(FR, &val, &c->eq_r[i], &t); fp_mul(FR, &t, &c->eq_c[col], &t); }
        fp_add(FR, &acc, &t, &acc);
    }
    c->out[i] = acc;
}
typedef struct { const ofp* mem; ofp* out; size_t n; uint64_t seed; const ofp* g1; const ofp* g2; int mode; size_t cells; } wl_hash_ctx;
static void wl_hash_body(long i, void* vc) {
    /* mode 0: derefs gather out[i] = mem[addr_i] (AddrTimestamps::deref, sparse_mlpoly_full.rs:245-257);
     * mode 1: hash of a (addr, val, ts) triple, addr * gamma^2 + val * gamma + ts - tau (:745-790) */
    wl_hash_ctx* c = (wl_hash_ctx*)vc;
    const size_t addr = (size_t)(wl_mix(c->seed + (uint64_t)i) % c->cells);
    if (c->mode == 0) { c->out[i] = c->mem[addr]; return; }
    ofp a, t, u;
    memset(&a, 0, sizeof a);
    a.l[0] = addr;
    fp_mul(FR, &a, c->g2, &t);
    fp_mul(FR, &c->mem[addr], c->g1, &u);
    fp_add(FR, &t, &u, &t);
    a.l[0] = (uint64_t)i & 0xfffff;
    fp_mul(FR, &a, c->g1, &u);            /* the timestamp enters Montgomery form through a product, as on the GPU */
    fp_add(FR, &t, &u, &c->out[i]);
}
typedef struct { const ofp* in; ofp* out; size_t half; } wl_prod_ctx;
static void wl_prod_body(long i, void* vc) {                              /* ProductCircuit::new layers, product_tree.rs:39-57 */
    wl_prod_ctx* c = (wl_prod_ctx*)vc;
    fp_mul(FR, &c->in[i], &c->in[(size_t)i + c->half], &c->out[i]);
}
typedef struct { const ofp* a; const ofp* b; size_t per, n; ofp* part; } wl_dot_ctx;
static void wl_dot_body(long blk, void* vc) {                             /* DensePolynomial::evaluate, hyrax.rs:217-222 */
    wl_dot_ctx* c = (wl_dot_ctx*)vc;
    size_t lo = (size_t)blk * c->per, hi = lo + c->per;
    if (hi > c->n) hi = c->n;
    fr_dot(c->a + lo, c->b + lo, hi > lo ? hi - lo : 0, &c->part[blk]);
}
static void wl_opening(const ofp* Z, size_t ell, const og1a* G, const og1a* h, const og1a* G1, int with_blinds, orc_transcript* tr,
                       orc_transcript* tape) {                            /* PolyEvalProof::prove, hyrax.rs:65-116 */
    ofp r[64], Zr, blind;
    for (size_t k = 0; k < ell; k++) t_challenge(tr, "r", &r[k]);
    wl_scalar(99, ell, &Zr);
    wl_scalar(98, ell, &blind);
    ofp* blinds = NULL;
    if (with_blinds) blinds = wl_table((size_t)1 << (ell / 2), 97, 0);
    orc_eval_proof proof;
    og1a cz; uint8_t czi;
    orc_poly_eval_prove(Z, ell, blinds, r, &Zr, &blind, G, h, G1, tr, tape, &proof, &cz, &czi);
    free(blinds);
}

#define WL_PHASES 13
static const char* const WL_NAMES[WL_PHASES] = {
    "witness_commit", "Az_Bz_Cz", "sumcheck_phase1", "eval_tables_phase2", "sumcheck_phase2", "witness_opening",
    "instance_evaluation", "eq_tables_derefs_gather", "derefs_commit", "network_construction", "product_layer_sumchecks",
    "hash_layer_evaluations", "hash_layer_openings"};
const char* orc_prove_workload_phase_name(int i) { return i >= 0 && i < WL_PHASES ? WL_NAMES[i] : ""; }
int orc_prove_workload_phases(void) { return WL_PHASES; }

/* s = log2(constraints) (even or odd, >= 10); derefs_rows_done: 0 = commit every row, otherwise only that many of the
 * non-zero rows are committed and seconds[8] is scaled up to all of them (flagged by the return value 1) -- the caller
 * states it.  gens_seconds (may be NULL) receives the one-off generator derivation time (SNARKGens::new, not part of prove). */
int orc_prove_workload(int s, int threads, size_t derefs_rows_done, double* seconds, double* gens_seconds) {
    if (s < 10 || s > 24) return -1;
    const size_t n = (size_t)1 << s, m = 4 * n;              /* nnz padded per matrix */
    const int ell_w = s, ell_d = s + 5, ell_o = s + 6, ell_m = s + 2;
    for (int i = 0; i < WL_PHASES; i++) seconds[i] = 0;
    double t0 = wall_now();
    /* generators: gens_r1cs_sat over 2^(s - s/2), gens_r1cs_eval over 2^(ell_o - ell_o/2) (its prefixes serve derefs / mem) */
    const size_t Rw = (size_t)1 << (ell_w - ell_w / 2), Ro = (size_t)1 << (ell_o - ell_o / 2);
    og1a* Gw = (og1a*)malloc(sizeof(og1a) * (Rw + 2));
    og1a* Go = (og1a*)malloc(sizeof(og1a) * (Ro + 2));
    orc_multi_commit_gens((const uint8_t*)"gens_r1cs_sat", 13, Rw + 1, Gw);
    orc_multi_commit_gens((const uint8_t*)"gens_r1cs_eval", 14, Ro + 1, Go);
    if (gens_seconds) *gens_seconds = wall_now() - t0;
    orc_transcript tr, tape;
    orc_transcript_new(&tr, (const uint8_t*)"snark", 5);
    orc_transcript_new(&tape, (const uint8_t*)"tape", 4);
    int scaled = 0;

    /* 0. R1CSProof::commit_poly (r1csproof.rs:210-237): witness polynomial, random blinds */
    ofp* z = wl_table(2 * n, 1, threads);                    /* (vars, 1, inputs) padded: z of the second sumcheck */
    {
        const size_t Lw = n / Rw;
        ofp* blinds = wl_table(Lw, 2, threads);
        og1a* C = (og1a*)malloc(sizeof(og1a) * Lw);
        uint8_t* inf = (uint8_t*)malloc(Lw);
        t0 = wall_now();
        orc_hyrax_commit(Gw, &Gw[Rw + 1], z, Lw, Rw, blinds, threads, C, inf);
        for (size_t i = 0; i < Lw; i++) t_point(&tr, "poly_commitment_share", &C[i], inf[i]);
        seconds[0] = wall_now() - t0;
        free(blinds); free(C); free(inf);
    }
    /* 1. Az, Bz, Cz (r1cs.rs:275-288) */
    ofp* T1[4];
    t0 = wall_now();
    {
        const int nnz[3] = {3, 2, 1};
        for (int k = 0; k < 3; k++) {
            T1[k + 1] = (ofp*)malloc(sizeof(ofp) * n);
            wl_sp_ctx c = {NULL, NULL, z, T1[k + 1], 2 * n, nnz[k], 100 + (uint64_t)k, 0};
            parallel_for((long)n, 1024, threads, wl_sparse_body, &c);
        }
    }
    seconds[1] = wall_now() - t0;
    /* 2. first ZK sumcheck (sumcheck.rs:465-649): eq(tau) table, s cubic rounds over 4 tables */
    t0 = wall_now();
    {
        ofp tau[64];
        for (int k = 0; k < s; k++) t_challenge(&tr, "challenge_tau", &tau[k]);
        T1[0] = (ofp*)malloc(sizeof(ofp) * n);
        orc_eq_evals(tau, (size_t)s, T1[0]);
        wl_sumcheck(T1, 4, n, 0, s, 4, Gw, &tr, threads);
    }
    seconds[2] = wall_now() - t0;
    for (int k = 0; k < 4; k++) free(T1[k]);
    /* 3. evaluation tables of the second phase: eq(rx) and r_A A^T eq + r_B B^T eq + r_C C^T eq (r1cs.rs compute_eval_table_sparse) */
    ofp* T2[2];
    t0 = wall_now();
    {
        ofp rx[64];
        for (int k = 0; k < s; k++) t_challenge(&tr, "rx", &rx[k]);
        ofp* eq = (ofp*)malloc(sizeof(ofp) * n);
        orc_eq_evals(rx, (size_t)s, eq);
        T2[1] = (ofp*)calloc(2 * n, sizeof(ofp));
        /* scatter val * eq[row] into column col: as many products as non-zeros; done as a gather over rows here */
        ofp* tmp = (ofp*)malloc(sizeof(ofp) * n);
        for (int k = 0; k < 3; k++) {
            const int nnz[3] = {3, 2, 1};
            wl_sp_ctx c = {NULL, NULL, eq, tmp, n, nnz[k], 200 + (uint64_t)k, 0};
            parallel_for((long)n, 1024, threads, wl_sparse_body, &c);
            for (size_t i = 0; i < n; i++) fp_add(FR, &T2[1][i], &tmp[i], &T2[1][i]);
        }
        free(tmp); free(eq);
    }
    seconds[3] = wall_now() - t0;
    /* 4. second ZK sumcheck (sumcheck.rs:657-811): s + 1 quadratic rounds over (z, ABC) */
    t0 = wall_now();
    T2[0] = (ofp*)malloc(sizeof(ofp) * 2 * n);
    memcpy(T2[0], z, sizeof(ofp) * 2 * n);
    wl_sumcheck(T2, 2, 2 * n, 1, s + 1, 4, Gw, &tr, threads);
    seconds[4] = wall_now() - t0;
    free(T2[0]); free(T2[1]);
    /* 5. opening of the witness commitment (r1csproof.rs:70-121 = hyrax.rs:65-116) */
    t0 = wall_now();
    wl_opening(z, (size_t)ell_w, Gw, &Gw[Rw + 1], &Gw[Rw], 1, &tr, &tape);
    seconds[5] = wall_now() - t0;
    /* 6. instance evaluation at (rx, ry) (snark.rs:455-460): eq tables + one pass over the non-zeros */
    t0 = wall_now();
    ofp *mem_rx = (ofp*)malloc(sizeof(ofp) * 2 * n), *mem_ry = (ofp*)malloc(sizeof(ofp) * 2 * n);
    {
        ofp rr[64];
        for (int k = 0; k < s + 1; k++) t_challenge(&tr, "ry", &rr[k]);
        orc_eq_evals(rr, (size_t)s, mem_rx);
        orc_eq_evals(rr, (size_t)s + 1, mem_ry);
        ofp* tmp = (ofp*)malloc(sizeof(ofp) * n);
        for (int k = 0; k < 3; k++) {
            const int nnz[3] = {3, 2, 1};
            wl_sp_ctx c = {mem_rx, mem_ry, NULL, tmp, 2 * n, nnz[k], 300 + (uint64_t)k, 1};
            parallel_for((long)n, 1024, threads, wl_sparse_body, &c);
        }
        free(tmp);
    }
    seconds[6] = wall_now() - t0;
    /* 7. derefs (sparse_mlpoly_full.rs:245-257, 292-297): 3 row + 3 col gathers of m entries, merged, zero-padded to 8 m */
    t0 = wall_now();
    ofp* derefs = (ofp*)calloc((size_t)1 << ell_d, sizeof(ofp));
    for (int k = 0; k < 6; k++) {
        wl_hash_ctx c = {k < 3 ? mem_rx : mem_ry, derefs + (size_t)k * m, m, 400 + (uint64_t)k, NULL, NULL, 0, k < 3 ? n : 2 * n};
        parallel_for((long)m, 4096, threads, wl_hash_body, &c);
    }
    seconds[7] = wall_now() - t0;
    /* 8. Derefs::commit (sparse_mlpoly_full.rs:301-304 -> hyrax.rs:253-267): zero blinds; the last quarter of the rows is zero */
    {
        const size_t Rd = (size_t)1 << (ell_d - ell_d / 2), Ld = ((size_t)1 << ell_d) / Rd, live = (6 * m + Rd - 1) / Rd;
        size_t done = live;
        if (derefs_rows_done && derefs_rows_done < live) { done = derefs_rows_done; scaled = 1; }
        og1a* C = (og1a*)malloc(sizeof(og1a) * Ld);
        uint8_t* inf = (uint8_t*)malloc(Ld);
        t0 = wall_now();
        orc_hyrax_commit(Go, &Go[Ro + 1], derefs, done, Rd, NULL, threads, C, inf);     /* gens_derefs = a prefix of gens_ops' G */
        for (size_t i = 0; i < done; i++) t_point(&tr, "poly_commitment_share", &C[i], inf[i]);
        seconds[8] = (wall_now() - t0) * ((double)live / (double)done);
        free(C); free(inf);
    }
    /* 9. network construction (sparse_mlpoly_full.rs:745-866): hash layers of both sides + their product circuits */
    const int ncirc = 16;
    ofp* circ[16];
    size_t clen[16];
    t0 = wall_now();
    {
        ofp g1, g2;
        t_challenge(&tr, "challenge_gamma_hash", &g1);
        fp_mul(FR, &g1, &g1, &g2);
        for (int k = 0; k < ncirc; k++) {
            const int side = k / 8, idx = k % 8;                 /* per side: init, 3 reads, 3 writes, audit */
            clen[k] = (idx == 0 || idx == 7) ? 2 * n : m;
            circ[k] = (ofp*)malloc(sizeof(ofp) * 2 * clen[k]);    /* layer 0 followed by the upper layers */
            wl_hash_ctx c = {side ? mem_ry : mem_rx, circ[k], clen[k], 500 + (uint64_t)k, &g1, &g2, 1, side ? 2 * n : n};
            parallel_for((long)clen[k], 4096, threads, wl_hash_body, &c);
            size_t off = 0, len = clen[k];
            while (len > 1) {
                wl_prod_ctx pc = {circ[k] + off, circ[k] + off + len, len / 2};
                parallel_for((long)(len / 2), 4096, threads, wl_prod_body, &pc);
                off += len;
                len /= 2;
            }
        }
    }
    seconds[9] = wall_now() - t0;
    /* 10. ProductLayerProof::prove (product_tree.rs:251-392): per layer, one batched cubic sumcheck (sumcheck.rs:165-330) over
     *     (left, right, eq) of every circuit that has that layer.  The reference batches the circuits into one sumcheck and lets
     *     Rayon split every round; here the circuits of a layer run side by side, one thread each, every one with its own copy
     *     of the transcript -- the same field work, and a thread team per layer instead of one per round. */
    t0 = wall_now();
    {
        int top = 0;
        while (((size_t)1 << top) < m) top++;
        for (int layer = 1; layer < top; layer++) {              /* a layer with 2^layer outputs per circuit */
            const size_t len = (size_t)1 << layer;
            ofp rr[64];
            for (int k = 0; k < layer; k++) t_challenge(&tr, "challenge_r_layer", &rr[k]);
            ofp* eq = (ofp*)malloc(sizeof(ofp) * len);
            orc_eq_evals(rr, (size_t)layer, eq);
            wl_layer_ctx lc = {circ, clen, eq, len, layer, Gw, &tr};
            parallel_for(ncirc, 1, threads, wl_layer_body, &lc);
            free(eq);
        }
    }
    seconds[10] = wall_now() - t0;
    for (int k = 0; k < ncirc; k++) free(circ[k]);
    /* 11. HashLayerProof::prove (sparse_mlpoly_full.rs:907-1100): evaluations of the derefs / comb_ops / comb_mem segments
     *     (hyrax.rs:217-222) and the three joint openings (hyrax.rs:65-116) */
    t0 = wall_now();
    {
        ofp rr[64];
        for (int k = 0; k < s + 2; k++) t_challenge(&tr, "r_ops", &rr[k]);
        ofp* eq = (ofp*)malloc(sizeof(ofp) * m);
        orc_eq_evals(rr, (size_t)s + 2, eq);
        const int nth = host_threads(threads);
        ofp* part = (ofp*)malloc(sizeof(ofp) * (size_t)nth * 4);
        ofp* ops = wl_table((size_t)1 << ell_o, 7, threads);     /* comb_ops: 15 segments of m + padding */
        for (int seg = 0; seg < 6 + 15 + 2; seg++) {
            const ofp* base = seg < 6 ? derefs + (size_t)seg * m : ops + (size_t)((seg - 6) % 15) * m;
            const size_t len = seg < 21 ? m : 2 * n;
            wl_dot_ctx dc = {base, eq, (len + (size_t)nth * 4 - 1) / ((size_t)nth * 4), len, part};
            parallel_for((long)nth * 4, 1, threads, wl_dot_body, &dc);
            ofp e;
            memset(&e, 0, sizeof e);
            for (int b = 0; b < nth * 4; b++) fp_add(FR, &e, &part[b], &e);
            t_scalar(&tr, "eval", &e);
        }
        free(part); free(eq);
        seconds[11] = wall_now() - t0;
        t0 = wall_now();
        const size_t Rd = (size_t)1 << (ell_d - ell_d / 2), Rm = (size_t)1 << (ell_m - ell_m / 2);
        wl_opening(derefs, (size_t)ell_d, Go, &Go[Ro + 1], &Go[Rd], 0, &tr, &tape);
        wl_opening(ops, (size_t)ell_o, Go, &Go[Ro + 1], &Go[Ro], 0, &tr, &tape);
        wl_opening(ops, (size_t)ell_m, Go, &Go[Ro + 1], &Go[Rm], 0, &tr, &tape);          /* comb_mem: 2^(s+2) values */
        free(ops);
    }
    seconds[12] = wall_now() - t0;
    free(derefs); free(mem_rx); free(mem_ry); free(z); free(Gw); free(Go);
    return scaled;
}
