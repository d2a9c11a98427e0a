"""CPU tests: the C oracle against the committed golden vectors (made by the independent Python
big-int model, tests/golden/make_golden.py) and against public constants."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "hyrax_golden.json")))


def h2i(s):
    return int(s, 16)


def pts(lst):
    return [None if p is None else (h2i(p[0]), h2i(p[1])) for p in lst]


def test_public_constants(orc):
    c = GOLD["constants"]
    # 2G as published in EIP-196 test data
    assert c["two_G"] == ["0x030644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd3",
                          "0x15ed738c0e0a7c92e7845f96b2ae9c0a68a6a449e3538fc7ff3ebf7a5a18a2c4"]
    g = orc.generator()
    one = orc.to_mont([1])[0]
    two = orc.to_mont([2])[0]
    p2, inf = orc.scalar_mul(g, 0, two)
    assert orc.points_to_ints(p2.reshape(1, 8), [inf]) == pts([c["two_G"]])
    rm1 = orc.to_mont([h2i(c["r"]) - 1])[0]
    pm1, inf = orc.scalar_mul(g, 0, rm1)
    assert orc.points_to_ints(pm1.reshape(1, 8), [inf]) == pts([c["minus_G"]])
    zero = orc.to_mont([0])[0]
    _, inf = orc.scalar_mul(g, 0, zero)
    assert inf == 1
    # group.rs:313-321 test_msm: MSM([2,3],[G,G]) == 5G
    G2 = np.stack([g, g])
    for algo in (0, 1):
        o, inf = orc.msm(G2, None, orc.to_mont([2, 3]), algo)
        assert orc.points_to_ints(o.reshape(1, 8), [inf]) == pts([c["five_G"]])
    assert orc.compress(g, 0).hex() == c["compress_G"]
    assert orc.compress(p2, 0).hex() == c["compress_2G"]
    assert orc.compress(pm1, 0).hex() == c["compress_minus_G"]
    assert orc.compress(np.zeros(8, dtype=np.uint64), 1).hex() == c["compress_identity"]
    assert one is not None


def test_merlin_vector(orc):
    t = orc.Transcript(b"test protocol")
    t.append_message(b"some label", b"some data")
    assert t.challenge_bytes(b"challenge", 32).hex() == GOLD["constants"]["merlin_test_vector"]


@pytest.mark.parametrize("label", list(GOLD["generators"].keys()))
def test_generator_derivation(orc, label):
    g = GOLD["generators"][label]
    sc, kinds = orc.gen_scalars(label.encode(), 16)
    assert orc.from_mont(sc) == [h2i(s) for s in g["scalars"]]
    assert [["primary", "fallback", "one"][k] for k in kinds] == g["kinds"]
    G, h = orc.multi_commit_gens(label.encode(), 16)
    assert orc.points_to_ints(G[:6], [0] * 6) == pts(g["points"])


def test_generator_degeneracy_stats(orc):
    """About two thirds of the reference's generators equal 1*G."""
    st = GOLD["generator_stats_1025"]["gens_r1cs_eval"]
    sc, kinds = orc.gen_scalars(b"gens_r1cs_eval", 1024)
    assert int((kinds == 2).sum()) == st["one"] and int((kinds == 0).sum()) == st["primary"]
    assert len(set(orc.from_mont(sc))) == st["distinct"]
    assert st["one"] > 600


def test_hyrax_commit_golden(orc):
    g = GOLD["hyrax_commit_4x8"]
    Gn, h, _ = orc.dotproduct_gens(g["label"].encode(), 8)
    Z = orc.to_mont([h2i(z) for z in g["Z"]])
    bl = orc.to_mont([h2i(b) for b in g["blinds"]])
    C, inf = orc.hyrax_commit(Gn, h, Z, 4, 8, bl)
    assert orc.points_to_ints(C, inf) == pts(g["C"])
    assert inf[3] == 1 and g["C"][3] is None          # all-zero row with zero blind -> identity
    assert [orc.compress(C[i], int(inf[i])).hex() for i in range(4)] == g["C_compressed"]


def test_bound_and_eq_golden(orc):
    g = GOLD["bound_4x8"]
    r = orc.to_mont([h2i(x) for x in g["r"]])
    L = orc.eq_evals(r[:2])
    R = orc.eq_evals(r[2:])
    assert orc.from_mont(L) == [h2i(x) for x in g["L"]]
    assert orc.from_mont(R) == [h2i(x) for x in g["R"]]
    Z = orc.to_mont([h2i(z) for z in GOLD["hyrax_commit_4x8"]["Z"]])
    assert orc.from_mont(orc.bound(Z, L, 4, 8)) == [h2i(x) for x in g["LZ"]]


def test_bullet_golden(orc):
    g = GOLD["bullet_n8"]
    Gn, h, _ = orc.dotproduct_gens(b"gens_r1cs_eval", 8)
    m = lambda xs: orc.to_mont([h2i(x) for x in xs])
    res = orc.bullet_prove(orc.generator(), Gn, h, m(g["a"]), m(g["b"]), m([g["blind"]])[0],
                           m([x for x, _ in g["blinds"]]), m([y for _, y in g["blinds"]]), m(g["u"]))
    assert orc.points_to_ints(res["L"], res["L_inf"]) == pts(g["L"])
    assert orc.points_to_ints(res["R"], res["R_inf"]) == pts(g["R"])
    assert orc.points_to_ints(res["Gamma"].reshape(1, 8), [res["Gamma_inf"]]) == pts([g["Gamma"]])
    assert orc.from_mont(res["a_hat"]) == [h2i(g["a_hat"])]
    assert orc.from_mont(res["b_hat"]) == [h2i(g["b_hat"])]
    assert orc.from_mont(res["blind_hat"]) == [h2i(g["blind_hat"])]
    assert orc.points_to_ints(res["g_hat"].reshape(1, 8), [res["g_hat_inf"]]) == pts([g["g_hat"]])


def test_pippenger_matches_naive_with_degenerate_inputs(orc):
    """Adversarial inputs the reference never tests: repeated bases, P + (-P), scalars 0, 1, r-1."""
    import pymodel as pm
    rng = pm.SplitMix64(11)
    n = 96
    ptsl = [pm.mul(rng.scalar(), pm.G) for _ in range(8)]
    P = [ptsl[i % 8] for i in range(n)]
    P[5] = pm.neg(P[4])
    P[9] = None
    sc = [rng.scalar() for _ in range(n)]
    sc[0], sc[1], sc[2] = 0, 1, pm.R - 1
    sc[5] = sc[4]
    Pa, inf = orc.points_from_ints(P)
    S = orc.to_mont(sc)
    a, ai = orc.msm(Pa, inf, S, 0)
    b, bi = orc.msm(Pa, inf, S, 1)
    assert ai == bi and np.array_equal(a, b)
    assert orc.points_to_ints(a.reshape(1, 8), [ai]) == [pm.msm(sc, P)]


def test_sumcheck_round_and_bind(orc):
    import pymodel as pm
    rng = pm.SplitMix64(3)
    n = 16
    T = [[rng.scalar() for _ in range(n)] for _ in range(4)]
    e = orc.sumcheck_cubic_eval(*[orc.to_mont(t) for t in T])
    half = n // 2
    exp = [0, 0, 0]
    for i in range(half):
        lo = [t[i] for t in T]
        hi = [t[half + i] for t in T]
        for k, tpt in enumerate((0, 2, 3)):
            v = [(l + tpt * (h - l)) % pm.R for l, h in zip(lo, hi)]
            exp[k] = (exp[k] + v[0] * (v[1] * v[2] - v[3])) % pm.R
    assert [orc.from_mont(x)[0] for x in e] == exp
    r = rng.scalar()
    bound = orc.bind_top(orc.to_mont(T[0]), orc.to_mont([r])[0])
    assert orc.from_mont(bound) == [(T[0][i] + r * (T[0][half + i] - T[0][i])) % pm.R for i in range(half)]


def test_prove_workload_runs_and_reports_every_phase(orc):
    """bench.py's CPU baseline of the end-to-end prove (orc_prove_workload): every phase of SNARK::prove is timed, the derefs
    commitment dominates as in the reference's own breakdown (BENCHMARK_RESULTS.md: 166 s of 209 s), and committing a sample
    of the derefs rows scales that phase instead of skipping it."""
    r = orc.prove_workload(10, threads=2)
    assert set(r["phases"]) >= {"witness_commit", "sumcheck_phase1", "sumcheck_phase2", "witness_opening", "derefs_commit",
                                "product_layer_sumchecks", "hash_layer_evaluations", "hash_layer_openings"}
    assert all(v >= 0 for v in r["phases"].values()) and r["seconds"] > 0 and not r["scaled"]
    r2 = orc.prove_workload(10, threads=2, derefs_rows=8)
    assert r2["scaled"] and r2["phases"]["derefs_commit"] > 0
    with pytest.raises(ValueError):
        orc.prove_workload(3)
