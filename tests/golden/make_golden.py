"""Generates tests/golden/*.json from oracle/pymodel.py (independent Python big-int model of the
reference behaviour).  The reference itself cannot run here (Rust, no toolchain) and ships no
golden vectors, so these pin OUR two oracles and the GPU against each other and against public
constants (2G of EIP-196, Merlin's published test vector).  Re-run:  python tests/golden/make_golden.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import pymodel as pm  # noqa: E402


def hx(v):
    return "0x%064x" % v


def pt(p):
    return None if p is None else [hx(p[0]), hx(p[1])]


out = {}
out["constants"] = {
    "p": hx(pm.P), "r": hx(pm.R),
    "two_G": pt(pm.mul(2, pm.G)), "three_G": pt(pm.mul(3, pm.G)),
    "minus_G": pt(pm.mul(pm.R - 1, pm.G)), "r_G_is_identity": pm.mul(pm.R, pm.G) is None,
    "msm_2_3_on_G_G": pt(pm.msm([2, 3], [pm.G, pm.G])), "five_G": pt(pm.mul(5, pm.G)),
    "compress_G": pm.compress(pm.G).hex(), "compress_2G": pm.compress(pm.mul(2, pm.G)).hex(),
    "compress_minus_G": pm.compress(pm.neg(pm.G)).hex(), "compress_identity": pm.compress(None).hex(),
    "merlin_test_vector": "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615",
}

gens = {}
for label in (b"gens_r1cs_sat", b"gens_r1cs_eval", b"test", b"test-gens"):
    sc = pm.gen_scalars(16, label)
    gens[label.decode()] = {
        "scalars": [hx(s) for s, _ in sc], "kinds": [k for _, k in sc],
        "points": [pt(pm.mul(s, pm.G)) for s, _ in sc[:6]],
    }
stats = {}
for label in (b"gens_r1cs_eval",):
    sc = pm.gen_scalars(1024, label)
    kinds = [k for _, k in sc]
    stats[label.decode()] = {k: kinds.count(k) for k in ("primary", "fallback", "one")}
    stats[label.decode()]["distinct"] = len(set(s for s, _ in sc))
out["generators"] = gens
out["generator_stats_1025"] = stats

# a small Hyrax commit: ell = 5 -> L = 4 rows x R = 8, reference generators, row 3 all-zero with
# zero blind (identity), row 1 blind non-zero
rng = pm.SplitMix64(7)
ell = 5
(gn, h), (g1, _) = pm.dotproduct_gens(8, b"gens_r1cs_eval")
Z = [rng.scalar() for _ in range(32)]
for j in range(8):
    Z[24 + j] = 0
blinds = [rng.scalar(), rng.scalar(), 0, 0]
C = pm.hyrax_commit(Z, blinds, (gn, h))
out["hyrax_commit_4x8"] = {
    "label": "gens_r1cs_eval", "Z": [hx(z) for z in Z], "blinds": [hx(b) for b in blinds],
    "C": [pt(c) for c in C], "C_compressed": [pm.compress(c).hex() for c in C],
}
r = [rng.scalar() for _ in range(ell)]
Lv, Rv = pm.factored_evals(r)
out["bound_4x8"] = {"r": [hx(x) for x in r], "L": [hx(x) for x in Lv], "R": [hx(x) for x in Rv],
                    "LZ": [hx(x) for x in pm.bound(Z, Lv, ell)]}
# bullet reduction n = 8 (mirrors nizk/bullet.rs:215-255 test shape) with fixed challenges
a = [rng.scalar() for _ in range(8)]
b = [rng.scalar() for _ in range(8)]
u = [rng.scalar() for _ in range(3)]
bl = [(rng.scalar(), rng.scalar()) for _ in range(3)]
blind = rng.scalar()
Q = pm.G
Ls, Rs, Gamma, ah, bh, gh, blh = pm.bullet_prove(Q, gn, h, a, b, blind, bl, u)
out["bullet_n8"] = {
    "a": [hx(x) for x in a], "b": [hx(x) for x in b], "u": [hx(x) for x in u],
    "blinds": [[hx(x), hx(y)] for x, y in bl], "blind": hx(blind),
    "L": [pt(p) for p in Ls], "R": [pt(p) for p in Rs], "Gamma": pt(Gamma),
    "a_hat": hx(ah), "b_hat": hx(bh), "g_hat": pt(gh), "blind_hat": hx(blh),
}
json.dump(out, open(os.path.join(HERE, "hyrax_golden.json"), "w"), indent=1)
print("wrote", os.path.join(HERE, "hyrax_golden.json"))
