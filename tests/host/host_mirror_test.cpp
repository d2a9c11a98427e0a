// GPU test of the C++ host mirror (spartan_bn254_b200/csrc/host/sbn254_host.hpp): generator derivation,
// DensePolynomial::commit / commit_inner / bound and msm_affine through the C ABI, checked bit-for-bit
// against the C oracle.  Reads like the reference's own tests (commitments.rs:160-186, group.rs:313-321).
#include <cstdio>
#include <cstring>
#include <vector>
#include "../../spartan_bn254_b200/csrc/host/sbn254_host.hpp"
#include "../../oracle/bn254_oracle.h"
using namespace sbn::host;

static uint64_t st = 7;
static uint64_t splitmix() { uint64_t z = (st += 0x9E3779B97F4A7C15ULL); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; return z ^ (z >> 31); }
static sbn_fr rand_fr() { sbn_fr f; for (int i = 0; i < 4; i++) f.l[i] = splitmix(); f.l[3] &= 0x0fffffffffffffffULL; return f; }   // < 2^252 < r: valid Montgomery residue

int main() {
    int fails = 0;
    Context ctx(0);
    // test_msm (group.rs:313-321): MSM([2,3],[G,G]) == 5G == MSM([5],[G])
    {
        GroupElement g = GroupElement::generator();
        uint64_t c2[4] = {2, 0, 0, 0}, c3[4] = {3, 0, 0, 0}, c5[4] = {5, 0, 0, 0};
        std::vector<uint64_t> canon = {2, 0, 0, 0, 3, 0, 0, 0, 5, 0, 0, 0};
        std::vector<sbn_fr> m(3);
        check(sbn_fr_from_canonical(ctx.get(), canon.data(), 3, m.data()), "from_canonical");
        GroupElement a = GroupElement::msm_affine(ctx, {m[0], m[1]}, {g.p, g.p});
        GroupElement b = GroupElement::msm_affine(ctx, {m[2]}, {g.p});
        if (a.inf || b.inf || memcmp(&a.p, &b.p, sizeof a.p)) { printf("msm KAT mismatch\n"); fails++; }
        GroupElement c = GroupElement::msm_affine(ctx, {m[0]}, {g.p, g.p});     // length mismatch -> identity
        if (!c.inf) { printf("length mismatch should give identity\n"); fails++; }
        (void)c2; (void)c3; (void)c5;
    }
    // generators + commit (ell = 9 -> 16 rows x 32 generators), one all-zero row, random blinds
    const size_t ell = 9;
    PolyCommitmentGens gens(ctx, ell, "gens_r1cs_eval");
    const size_t R = gens.gens.gens_n.n, L = (size_t(1) << ell) / R;
    std::vector<og1a> og(R + 2);
    orc_multi_commit_gens((const uint8_t*)"gens_r1cs_eval", 14, R + 1, og.data());
    if (memcmp(og.data(), gens.gens.gens_n.G.data(), R * sizeof(og1a)) || memcmp(&og[R + 1], &gens.gens.gens_n.h, sizeof(og1a)) ||
        memcmp(&og[R], &gens.gens.gens_1.G[0], sizeof(og1a))) { printf("generator mismatch\n"); fails++; }
    std::vector<sbn_fr> Z(L * R), blinds(L);
    for (auto& z : Z) z = rand_fr();
    for (auto& b : blinds) b = rand_fr();
    for (size_t j = 0; j < R; j++) Z[3 * R + j] = sbn_fr{};
    blinds[3] = sbn_fr{};
    DensePolynomial poly(Z);
    auto res = poly.commit(gens, blinds);
    std::vector<og1a> oc(L);
    std::vector<uint8_t> oinf(L);
    orc_hyrax_commit(og.data(), &og[R + 1], (const ofp*)Z.data(), L, R, (const ofp*)blinds.data(), 0, oc.data(), oinf.data());
    if (memcmp(oc.data(), res.first.C.data(), L * sizeof(og1a)) || memcmp(oinf.data(), res.first.inf.data(), L)) { printf("commit mismatch\n"); fails++; }
    if (!res.first.inf[3]) { printf("zero row should commit to the identity\n"); fails++; }
    // zero blinds (random_tape = None)
    auto res0 = poly.commit(gens);
    orc_hyrax_commit(og.data(), &og[R + 1], (const ofp*)Z.data(), L, R, nullptr, 0, oc.data(), oinf.data());
    if (memcmp(oc.data(), res0.first.C.data(), L * sizeof(og1a))) { printf("zero-blind commit mismatch\n"); fails++; }
    // precondition: wrong generator count panics in the reference (commitments.rs:146)
    try { poly.commit_inner(std::vector<sbn_fr>(L * 2), gens.gens.gens_n); printf("missing shape error\n"); fails++; } catch (const std::logic_error&) {}
    // bound
    std::vector<sbn_fr> Lv(L);
    for (auto& v : Lv) v = rand_fr();
    auto lz = poly.bound(ctx, Lv);
    std::vector<ofp> olz(R);
    orc_bound((const ofp*)Z.data(), (const ofp*)Lv.data(), L, R, 0, olz.data());
    if (memcmp(olz.data(), lz.data(), R * sizeof(ofp))) { printf("bound mismatch\n"); fails++; }
    // Pedersen commit of one vector + scale
    GroupElement c1 = gens.gens.gens_n.commit(std::vector<sbn_fr>(Z.begin(), Z.begin() + R), blinds[0]);
    if (c1.inf || memcmp(&c1.p, &res.first.C[0], sizeof c1.p)) { printf("single commit mismatch\n"); fails++; }
    MultiCommitGens scaled = gens.gens.gens_1.scale(blinds[1]);
    og1a sp; uint8_t sinf;
    orc_g1_scalar_mul(&og[R], 0, (const ofp*)&blinds[1], &sp, &sinf);
    if (memcmp(&sp, &scaled.G[0], sizeof sp)) { printf("scale mismatch\n"); fails++; }
    printf("host mirror: %s\n", fails ? "FAIL" : "ok");
    return fails ? 1 : 0;
}
