"""GPU parity tests of the Hyrax opening path through the C ABI: single MSMs, Pedersen commit, bound,
the bullet reduction (device-resident generators / vectors, host-side challenges) and the sumcheck round."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "hyrax_golden.json")))
R_MOD = int(GOLD["constants"]["r"], 16)


def h2i(s):
    return int(s, 16)


def test_msm_reference_kat(ctx, orc):
    """group.rs:313-321 test_msm: MSM([2,3],[G,G]) == 5G; group.rs:304-311: 2G == G+G."""
    from spartan_bn254_b200.hyrax import GroupElement
    g = GroupElement.generator().xy
    out = GroupElement.msm_affine(orc.to_mont([2, 3]), np.stack([g, g]), ctx=ctx)
    assert out.affine_ints() == tuple(h2i(x) for x in GOLD["constants"]["five_G"])
    two = GroupElement.msm_affine(orc.to_mont([1, 1]), np.stack([g, g]), ctx=ctx)
    assert two.affine_ints() == tuple(h2i(x) for x in GOLD["constants"]["two_G"])
    # length mismatch -> identity (`unwrap_or_default`, group.rs:156,173)
    assert GroupElement.msm_affine(orc.to_mont([2]), np.stack([g, g]), ctx=ctx).inf == 1
    # compress round trip vectors
    assert GroupElement(g).compress().hex() == GOLD["constants"]["compress_G"]
    assert GroupElement.identity().compress().hex() == GOLD["constants"]["compress_identity"]


@pytest.mark.parametrize("n", [1, 2, 33, 64, 300, 1024])
def test_msm_matches_oracle(ctx, orc, n):
    from spartan_bn254_b200 import synth
    G, _ = synth.distinct_generators(ctx, max(n, 2))
    pts = G[:n].copy()
    inf = np.zeros(n, dtype=np.uint8)
    sc = synth.uniform_scalars(17, n)
    if n >= 33:
        pts[5] = pts[4]                       # repeated base
        sc[7] = 0
        inf[9] = 1                            # identity among the inputs
        sc[11] = orc.to_mont([R_MOD - 1])[0]
    out, oinf = ctx.msm(pts, inf, sc)
    exp, einf = orc.msm(pts, inf, sc, 1)
    assert oinf == einf and np.array_equal(out, exp)


def test_pedersen_commit_matches_oracle(ctx, orc):
    """<[Scalar] as Commitments>::commit (commitments.rs:144-154) incl. the n = 1 case of Scalar::commit."""
    from spartan_bn254_b200 import synth
    from spartan_bn254_b200.hyrax import MultiCommitGens
    for n in (1, 3, 64):
        gens = MultiCommitGens.new(n, b"test", ctx)
        sc = synth.uniform_scalars(5, n)
        blind = synth.uniform_scalars(6, 1)[0]
        got = gens.commit(sc, blind)
        exp, einf = orc.msm(np.vstack([gens.G, gens.h[None, :]]), None, np.vstack([sc, blind[None, :]]), 0)
        assert got.inf == einf and np.array_equal(got.xy, exp)
    with pytest.raises(AssertionError):
        gens.commit(sc[:5], blind)            # commitments.rs:146


@pytest.mark.parametrize("ell", [5, 10, 13])
def test_bound_matches_oracle(ctx, orc, ell):
    from spartan_bn254_b200 import synth
    from spartan_bn254_b200.hyrax import DensePolynomial, compute_factored_lens
    l, r = compute_factored_lens(ell)
    Z = synth.uniform_scalars(3, 1 << ell)
    Lv = synth.uniform_scalars(4, 1 << l)
    got = DensePolynomial(Z).bound(Lv, ctx)
    assert np.array_equal(got, orc.bound(Z, Lv, 1 << l, 1 << r))


def test_bound_golden(ctx, orc):
    g = GOLD["bound_4x8"]
    Z = orc.to_mont([h2i(z) for z in GOLD["hyrax_commit_4x8"]["Z"]])
    L = orc.to_mont([h2i(x) for x in g["L"]])
    assert orc.from_mont(ctx.bound(Z, L, 4, 8)) == [h2i(x) for x in g["LZ"]]


def run_bullet(ctx, gens_n, Q, a, b, blind, bl, br, u, orc, bases=None, q_scalar=None):
    if q_scalar is not None:
        st = ctx.bullet_begin(bases, None, a, b, blind, q_scalar=q_scalar)     # table-based rounds, Q = q * g1
    else:
        st = ctx.bullet_begin(gens_n.device_bases(), Q, a, b, blind)           # arbitrary Q, generators folded
    Ls, Rs = [], []
    lg = len(u)
    for i in range(lg):
        (L, Li), (R, Ri) = st.round(bl[i], br[i])
        Ls.append((L, Li)); Rs.append((R, Ri))
        ui = orc.to_mont([pow(orc.from_mont(u[i])[0], -1, R_MOD)])[0]
        st.fold(u[i], ui)
    a_hat, b_hat, g_hat, g_inf = st.end()
    out = dict(L=Ls, R=Rs, Gamma=(st.Gamma, st.Gamma_inf), a_hat=a_hat, b_hat=b_hat, g_hat=(g_hat, g_inf))
    st.close()
    return out


def test_bullet_golden_n8(ctx, orc):
    """nizk/bullet.rs:215-255 test shape (n = 8) with fixed challenges, against the Python-model vectors."""
    from spartan_bn254_b200.hyrax import DotProductProofGens, GroupElement
    g = GOLD["bullet_n8"]
    m = lambda xs: orc.to_mont([h2i(x) for x in xs])
    gens = DotProductProofGens(8, b"gens_r1cs_eval", ctx)
    res = run_bullet(ctx, gens.gens_n, GroupElement.generator().xy, m(g["a"]), m(g["b"]), m([g["blind"]])[0],
                     m([x for x, _ in g["blinds"]]), m([y for _, y in g["blinds"]]), m(g["u"]), orc)
    pt = lambda p: None if p is None else (h2i(p[0]), h2i(p[1]))
    for k in range(3):
        assert orc.points_to_ints(res["L"][k][0].reshape(1, 8), [res["L"][k][1]]) == [pt(g["L"][k])]
        assert orc.points_to_ints(res["R"][k][0].reshape(1, 8), [res["R"][k][1]]) == [pt(g["R"][k])]
    assert orc.points_to_ints(res["Gamma"][0].reshape(1, 8), [res["Gamma"][1]]) == [pt(g["Gamma"])]
    assert orc.from_mont(res["a_hat"]) == [h2i(g["a_hat"])]
    assert orc.from_mont(res["b_hat"]) == [h2i(g["b_hat"])]
    assert orc.points_to_ints(res["g_hat"][0].reshape(1, 8), [res["g_hat"][1]]) == [pt(g["g_hat"])]


@pytest.mark.parametrize("path", ["fold", "tables"])
@pytest.mark.parametrize("n,label", [(2, b"test"), (16, b"gens_r1cs_eval"), (256, b"gens_r1cs_eval"), (1024, b"gens_r1cs_sat")])
def test_bullet_matches_oracle(ctx, orc, n, label, path):
    """Opening sizes of cfg3 (n = 1024) and the reference's n = 16 log-proof test (nizk/mod.rs:575-712)."""
    from spartan_bn254_b200 import synth
    from spartan_bn254_b200.hyrax import DotProductProofGens
    lg = n.bit_length() - 1
    gens = DotProductProofGens(n, label, ctx)
    a = synth.uniform_scalars(31, n)
    b = synth.uniform_scalars(32, n)
    u = synth.uniform_scalars(33, lg)
    bl = synth.uniform_scalars(34, lg)
    br = synth.uniform_scalars(35, lg)
    blind = synth.uniform_scalars(36, 1)[0]
    Q = gens.gens_1.G[0]
    if path == "fold":
        res = run_bullet(ctx, gens.gens_n, Q, a, b, blind, bl, br, u, orc)
    else:                                     # Q = q * g1 with q a full-width scalar (the reference's r * gens_1.G[0])
        q = synth.uniform_scalars(37, 1)[0]
        Q, qinf = orc.scalar_mul(gens.gens_1.G[0], 0, q)
        assert not qinf
        res = run_bullet(ctx, gens.gens_n, None, a, b, blind, bl, br, u, orc, bases=gens.device_bases_ext(), q_scalar=q)
    exp = orc.bullet_prove(Q, gens.gens_n.G, gens.gens_n.h, a, b, blind, bl, br, u)
    for k in range(lg):
        assert res["L"][k][1] == exp["L_inf"][k] and np.array_equal(res["L"][k][0], exp["L"][k]), k
        assert res["R"][k][1] == exp["R_inf"][k] and np.array_equal(res["R"][k][0], exp["R"][k]), k
    assert np.array_equal(res["Gamma"][0], exp["Gamma"]) and res["Gamma"][1] == exp["Gamma_inf"]
    assert np.array_equal(res["a_hat"], exp["a_hat"]) and np.array_equal(res["b_hat"], exp["b_hat"])
    assert np.array_equal(res["g_hat"][0], exp["g_hat"]) and res["g_hat"][1] == exp["g_hat_inf"]


def test_bullet_shape_errors(ctx):
    from spartan_bn254_b200 import SbnError, synth
    from spartan_bn254_b200.hyrax import MultiCommitGens
    gens = MultiCommitGens.new(8, b"test", ctx)
    a = synth.uniform_scalars(1, 4)
    with pytest.raises(SbnError) as e:
        ctx.bullet_begin(gens.device_bases(), gens.G[0], a, a, a[0])      # G.len() != a.len()  (bullet.rs:42)
    assert e.value.status == -2


@pytest.mark.parametrize("lg", [4, 12])
def test_sumcheck_rounds_match_oracle(ctx, orc, lg):
    """sumcheck.rs:501-530 (evaluate at 0, 2, 3) and :551-554 (bind) for every round down to length 1."""
    from spartan_bn254_b200 import synth
    n = 1 << lg
    T = [synth.uniform_scalars(40 + k, n) for k in range(4)]
    rs = synth.uniform_scalars(50, lg)
    st = ctx.sumcheck_begin(*T)
    cur = [t.copy() for t in T]
    for j in range(lg):
        e = st.round_eval()
        exp = orc.sumcheck_cubic_eval(*cur)
        for k in range(3):
            assert np.array_equal(e[k], exp[k]), (j, k)
        st.bind(rs[j])
        cur = [orc.bind_top(t, rs[j]) for t in cur]
    fin = st.end()
    for k in range(4):
        assert np.array_equal(fin[k], cur[k][0])
    st.close()


def test_cpp_host_mirror(orc, tmp_path):
    """The C++ host mirror of the reference interface (compiled-language host side above the C ABI)."""
    import subprocess
    root = os.path.dirname(HERE)
    exe = str(tmp_path / "host_mirror_test")
    libdir = os.path.join(root, "spartan_bn254_b200", "_lib")
    odir = os.path.join(root, "oracle", "_build")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(root, "tests", "host", "host_mirror_test.cpp"),
                           "-L" + libdir, "-lsbn254", "-L" + odir, "-loracle", "-Wl,-rpath," + libdir, "-Wl,-rpath," + odir,
                           "-pthread"])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "host mirror: ok" in out.stdout


@pytest.mark.parametrize("ell,label,use_blinds", [(4, b"gens_r1cs_sat", True), (8, b"gens_r1cs_eval", False),
                                                   (12, b"gens_r1cs_sat", True)])
def test_poly_eval_proof_roundtrip_and_byte_parity(ctx, orc, ell, label, use_blinds):
    """PolyEvalProof::prove (hyrax.rs:65-116) driven through the GPU -- bound, the (n+1)-point commitment Cx, the
    device-resident bullet reduction, the 2-point commitments -- with Merlin challenges on the host:
      * the proof verifies under the independent CPU verifier (hyrax.rs:118-137 restated in the oracle), like the
        reference's own prove->verify tests (nizk/mod.rs:575-712);
      * with the same injected random-tape seed it is byte-identical to the CPU prover's proof."""
    from spartan_bn254_b200 import synth
    from spartan_bn254_b200.hyrax import (DensePolynomial, PolyCommitmentGens, PolyEvalProof, fr_vec_to_ints, fr_to_int,
                                          compute_factored_lens)
    from spartan_bn254_b200.transcript import Transcript, RandomTape
    l, r_ = compute_factored_lens(ell)
    L_size, n = 1 << l, 1 << r_
    gens = PolyCommitmentGens(ell, label, ctx)
    Z = synth.uniform_scalars(61, 1 << ell)
    blinds = synth.uniform_scalars(62, L_size) if use_blinds else None
    rpt = synth.uniform_scalars(63, ell)
    poly = DensePolynomial(Z)
    comm, _ = poly.commit(gens, blinds)
    Zr_m = orc.evaluate(Z, rpt)
    blind_Zr = 424242 if use_blinds else 0
    seed = 987654321
    tr = Transcript(b"example")
    tape = RandomTape(b"proof", seed)
    proof, C_Zr = PolyEvalProof.prove(poly, blinds, fr_vec_to_ints(rpt), fr_to_int(Zr_m), blind_Zr, gens, tr, tape)

    gn, g1 = gens.gens.gens_n, gens.gens.gens_1
    # CPU prover with the same tape
    otr = orc.Transcript(b"example")
    otape = orc.Transcript(b"proof")
    otape.append_message(b"init_randomness", seed.to_bytes(32, "little"))
    oproof, oC, oCi = orc.poly_eval_prove(Z, ell, blinds, rpt, Zr_m, orc.to_mont([blind_Zr])[0] if use_blinds else None,
                                          gn.G, gn.h, g1.G[0], otr, otape)
    od = oproof.to_dict()
    p = proof.proof
    assert len(p.L_vec) == len(od["L"]) == r_
    for k in range(r_):
        assert p.L_vec[k].inf == od["L"][k][1] and np.array_equal(p.L_vec[k].xy, od["L"][k][0]), k
        assert p.R_vec[k].inf == od["R"][k][1] and np.array_equal(p.R_vec[k].xy, od["R"][k][0]), k
    assert np.array_equal(p.delta.xy, od["delta"][0]) and np.array_equal(p.beta.xy, od["beta"][0])
    assert orc.from_mont(od["z1"]) == [p.z1] and orc.from_mont(od["z2"]) == [p.z2]
    assert np.array_equal(C_Zr.xy, oC) and C_Zr.inf == oCi
    # both transcripts end in the same state
    assert tr.challenge_bytes(b"final", 32) == otr.challenge_bytes(b"final", 32)

    # independent verification of the GPU-made proof against the GPU-made commitment
    gd = dict(L=[(e.xy, e.inf) for e in p.L_vec], R=[(e.xy, e.inf) for e in p.R_vec], delta=(p.delta.xy, p.delta.inf),
              beta=(p.beta.xy, p.beta.inf), z1=orc.to_mont([p.z1])[0], z2=orc.to_mont([p.z2])[0])
    assert orc.poly_eval_verify(orc.EvalProof.from_dict(gd), ell, rpt, C_Zr.xy, C_Zr.inf, comm.C, comm.inf, gn.G, gn.h,
                                g1.G[0], orc.Transcript(b"example"))
    gd["z2"] = gd["z2"].copy()
    gd["z2"][0] ^= np.uint64(1)
    assert not orc.poly_eval_verify(orc.EvalProof.from_dict(gd), ell, rpt, C_Zr.xy, C_Zr.inf, comm.C, comm.inf, gn.G, gn.h,
                                    g1.G[0], orc.Transcript(b"example"))


def test_resident_polynomial_commit_and_bound(ctx, orc):
    """sbn_poly_*: Z uploaded once, then committed and bound from HBM -- same results as the host-pointer calls."""
    from spartan_bn254_b200 import synth
    from spartan_bn254_b200.hyrax import DensePolynomial, PolyCommitmentGens, compute_factored_lens
    ell = 11
    l, r_ = compute_factored_lens(ell)
    gens = PolyCommitmentGens(ell, b"gens_r1cs_eval", ctx)
    Z = synth.derefs_scalars(ell)
    blinds = synth.uniform_scalars(4, 1 << l)
    poly = DensePolynomial(Z)
    poly.resident(ctx)
    comm, _ = poly.commit(gens, blinds)
    gn = gens.gens.gens_n
    C, inf = orc.hyrax_commit(gn.G, gn.h, Z, 1 << l, 1 << r_, blinds)
    assert np.array_equal(comm.C, C) and np.array_equal(comm.inf, inf)
    Lv = synth.uniform_scalars(5, 1 << l)
    assert np.array_equal(poly.bound(Lv, ctx), orc.bound(Z, Lv, 1 << l, 1 << r_))
    from spartan_bn254_b200 import SbnError
    with pytest.raises(SbnError):
        poly.resident(ctx).bound(Lv, 1 << l, 1 << (r_ + 1))          # L * R != len


def test_sumcheck_quadratic_rounds_match_oracle(ctx, orc):
    """Phase 2 of the R1CS-sat proof (sumcheck.rs:690-699): two tables, evaluations at 0 and 2, then bind."""
    from spartan_bn254_b200 import synth
    lg = 10
    T = [synth.uniform_scalars(70 + k, 1 << lg) for k in range(2)]
    rs = synth.uniform_scalars(72, lg)
    st = ctx.sumcheck_begin_quad(*T)
    cur = [t.copy() for t in T]
    for j in range(lg):
        e = st.round_eval()
        exp = orc.sumcheck_quad_eval(*cur)
        assert len(e) == 2 and np.array_equal(e[0], exp[0]) and np.array_equal(e[1], exp[1]), j
        st.bind(rs[j])
        cur = [orc.bind_top(t, rs[j]) for t in cur]
    fin = st.end()
    assert np.array_equal(fin[0], cur[0][0]) and np.array_equal(fin[1], cur[1][0]) and not fin[2:].any()
    st.close()
