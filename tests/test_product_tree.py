"""Product-layer argument: product circuits + batched cubic sumcheck.

CPU part: the oracle's prover (oracle/product_model.py over the C Merlin restatement) must be accepted by the package's
verifier (pure-Python Merlin) -- two independent transcripts and two independent restatements of product_tree.rs.
GPU part: the GPU-backed prover must reproduce the oracle's proof integer for integer (bit-exact field elements) and be
accepted by the verifier; at full size (2^20 entries) acceptance is the size-independent property."""
import random

import numpy as np
import pytest

R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001


def _instance(seed, n, P, S):
    rnd = random.Random(seed)
    prod = [[rnd.randrange(R) for _ in range(n)] for _ in range(P)]
    dotp = [tuple([rnd.randrange(R) for _ in range(n // 2)] for _ in range(3)) for _ in range(S)]
    return prod, dotp


def _verify(proof_layers, claims_dotp, claims_prod, dotp_claims, n, label=b"prodtest"):
    from spartan_bn254_b200.product_tree import (ProductCircuitEvalProofBatched, LayerProofBatched, SumcheckInstanceProof,
                                                 CompressedUniPoly)
    from spartan_bn254_b200.transcript import Transcript
    layers = [LayerProofBatched(SumcheckInstanceProof([CompressedUniPoly(c) for c in polys]), left, right)
              for polys, left, right in proof_layers]
    proof = ProductCircuitEvalProofBatched(layers, claims_dotp)
    return proof.verify(claims_prod, dotp_claims, n, Transcript(label))


@pytest.mark.parametrize("n,P,S", [(2, 1, 0), (16, 2, 0), (64, 3, 2), (256, 1, 4)])
def test_oracle_prover_accepted_by_host_verifier(n, P, S):
    import oracle as orc
    import product_model as pm
    prod, dotp = _instance(7 + n, n, P, S)
    out = pm.prove_batched(prod, dotp, orc.Transcript(b"prodtest"))
    assert len(out["layers"]) == n.bit_length() - 1
    dotp_claims = [pm.dotp_evaluate(*d) for d in dotp]
    claims, claims_dotp, rand = _verify(out["layers"], out["claims_dotp"], out["claims_prod"], dotp_claims, n)
    assert rand == out["rand"]
    # the final claims are the circuits' input polynomials evaluated at rand (product_tree.rs doc of verify)
    eq = pm.eq_evals(rand)
    for p, c in zip(prod, claims):
        assert sum(a * b for a, b in zip(p, eq)) % R == c
    # a tampered claim is rejected
    bad = list(out["claims_prod"])
    bad[0] = (bad[0] + 1) % R
    with pytest.raises(ValueError):
        _verify(out["layers"], out["claims_dotp"], bad, dotp_claims, n)


def test_unipoly_matches_reference_kats():
    """unipoly.rs:130-184 tests: interpolation of 2x^2+3x+1 and x^3+2x^2+3x+1 from evaluations."""
    from spartan_bn254_b200.product_tree import UniPoly
    q = UniPoly.from_evals([1, 6, 15])
    assert q.coeffs == [1, 3, 2]
    c = UniPoly.from_evals([1, 7, 23, 55])
    assert c.coeffs == [1, 3, 2, 1]
    assert c.compress().decompress((c.eval_at_zero() + c.eval_at_one()) % R).coeffs == c.coeffs
    assert c.evaluate(4) == 109


@pytest.mark.gpu
@pytest.mark.parametrize("native", [True, False])
@pytest.mark.parametrize("n,P,S", [(2, 1, 0), (8, 2, 1), (64, 3, 2), (1024, 2, 0), (4096, 1, 2)])
def test_gpu_prover_matches_oracle(ctx, orc, n, P, S, native):
    """native: the layer's round loop runs inside the library against the native Merlin state (sbn_bsumcheck_prove);
    otherwise round by round from Python over the pure-Python transcript (sbn_bsumcheck_round_eval / _bind / _end).  Both
    must reproduce the oracle's proof integer for integer."""
    import product_model as pm
    from spartan_bn254_b200.hyrax import fr_vec_from_ints
    from spartan_bn254_b200.product_tree import ProductCircuit, DotProductCircuit, ProductCircuitEvalProofBatched
    from spartan_bn254_b200.transcript import Transcript
    prod, dotp = _instance(100 + n, n, P, S)
    want = pm.prove_batched(prod, dotp, orc.Transcript(b"prodtest"))
    circuits = [ProductCircuit(ctx, fr_vec_from_ints(p)) for p in prod]
    assert [c.evaluate() for c in circuits] == want["claims_prod"]
    dcs = [DotProductCircuit(*[fr_vec_from_ints(t) for t in d]) for d in dotp]
    proof, rand = ProductCircuitEvalProofBatched.prove(ctx, circuits, dcs, Transcript(b"prodtest", native=native))
    assert rand == want["rand"]
    assert len(proof.proof) == len(want["layers"])
    for got, (polys, left, right) in zip(proof.proof, want["layers"]):
        assert [cp.coeffs_except_linear_term for cp in got.proof.compressed_polys] == polys
        assert got.claims_prod_left == left and got.claims_prod_right == right
    assert tuple(proof.claims_dotp) == tuple(want["claims_dotp"])
    dotp_claims = [pm.dotp_evaluate(*d) for d in dotp]
    proof.verify(want["claims_prod"], dotp_claims, n, Transcript(b"prodtest"))
    for c in circuits:
        c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("n,P,S", [(2, 1, 0), (64, 3, 2), (1024, 2, 0), (4096, 1, 2), (1 << 14, 12, 0)])
def test_device_side_transcript_matches_host_loop(ctx, orc, n, P, S, mode):
    """sbn_bsumcheck_prove with the Merlin transcript on the device (csrc/transcript_kernels.cuh; sbn_ctx_set "bsc_device":
    1 = the short last rounds of a layer in one block, 2 = every round) must leave the same proof, the same challenges AND
    the same transcript state as the host round loop (0, the default): the next layer continues from that state."""
    import product_model as pm
    from spartan_bn254_b200.hyrax import fr_vec_from_ints
    from spartan_bn254_b200.product_tree import ProductCircuit, DotProductCircuit, ProductCircuitEvalProofBatched
    from spartan_bn254_b200.transcript import Transcript
    prod, dotp = _instance(300 + n, n, P, S)
    out = {}
    try:
        for m in (0, mode):
            ctx.set("bsc_device", m)
            circuits = [ProductCircuit(ctx, fr_vec_from_ints(p)) for p in prod]
            dcs = [DotProductCircuit(*[fr_vec_from_ints(t) for t in d]) for d in dotp]
            tr = Transcript(b"prodtest")
            proof, rand = ProductCircuitEvalProofBatched.prove(ctx, circuits, dcs, tr)
            out[m] = (rand, [[cp.coeffs_except_linear_term for cp in l.proof.compressed_polys] for l in proof.proof],
                      [(l.claims_prod_left, l.claims_prod_right) for l in proof.proof], tuple(proof.claims_dotp),
                      tr.challenge_scalar(b"after"))
            for c in circuits:
                c.close()
    finally:
        ctx.set("bsc_device", 0)
    assert out[0] == out[mode]
    if n <= 4096:
        want = pm.prove_batched(prod, dotp, orc.Transcript(b"prodtest"))
        assert out[mode][0] == want["rand"]


@pytest.mark.gpu
def test_gpu_prover_full_size_is_accepted(ctx):
    """2^20-entry circuits (the hash-layer size of a 2^18-constraint instance): acceptance by the verifier, whose work is
    logarithmic, plus the final claims against a GPU-independent evaluation of one input polynomial at rand."""
    from spartan_bn254_b200 import synth
    from spartan_bn254_b200.hyrax import fr_vec_to_ints, EqPolynomial
    from spartan_bn254_b200.product_tree import ProductCircuit, ProductCircuitEvalProofBatched
    from spartan_bn254_b200.transcript import Transcript
    n, P = 1 << 20, 3
    polys = [synth.uniform_scalars(40 + i, n) for i in range(P)]
    circuits = [ProductCircuit(ctx, p) for p in polys]
    claims = [c.evaluate() for c in circuits]
    proof, rand = ProductCircuitEvalProofBatched.prove(ctx, circuits, [], Transcript(b"full"))
    final, _, rand_v = proof.verify(claims, [], n, Transcript(b"full"))
    assert rand_v == rand
    # final claim of circuit 0 = its input polynomial at rand: check through the GPU `bound` of the Hyrax path
    # (a different kernel) with the factored eq tables, then the short dot product on the host
    from spartan_bn254_b200.hyrax import DensePolynomial, fr_vec_from_ints
    Lv, Rv = EqPolynomial(rand).compute_factored_evals()
    LZ = DensePolynomial(polys[0]).bound(fr_vec_from_ints(Lv), ctx)
    val = sum(a * b for a, b in zip(fr_vec_to_ints(LZ), Rv)) % R
    assert val == final[0]
    for c in circuits:
        c.close()


@pytest.mark.gpu
def test_prover_entry_points_reject_bad_shapes(ctx):
    """The C ABI returns SBN_ERR_SHAPE / SBN_ERR_ARG instead of unwinding where the reference would assert
    (product_tree.rs:272,298-300; sumcheck rounds past the last one; non-power-of-two tables)."""
    from spartan_bn254_b200 import SbnError, synth
    from spartan_bn254_b200.lib import ProdCircuit, BatchedSumcheckState, SpMat, Addrs
    with pytest.raises(SbnError) as e:
        ProdCircuit(ctx, synth.uniform_scalars(1, 24))                       # not a power of two
    assert e.value.status == -2
    with pytest.raises(SbnError):
        ProdCircuit(ctx, synth.uniform_scalars(1, 1))                        # a product circuit needs two entries
    pc = ProdCircuit(ctx, synth.uniform_scalars(1, 16))
    with pytest.raises(SbnError) as e:                                       # layer 0 has 8 + 8 entries: |eq(rand)| must be 8
        BatchedSumcheckState(ctx, [pc], 0, synth.uniform_scalars(2, 2))
    assert e.value.status == -2
    with pytest.raises(SbnError):
        BatchedSumcheckState(ctx, [pc], 9, synth.uniform_scalars(2, 3))       # no such layer
    with pytest.raises(SbnError):                                            # a sequential instance of the wrong length
        BatchedSumcheckState(ctx, [pc], 0, synth.uniform_scalars(2, 3), [tuple(synth.uniform_scalars(3 + i, 4) for i in range(3))])
    st = BatchedSumcheckState(ctx, [pc], 0, synth.uniform_scalars(2, 3))
    with pytest.raises(SbnError):
        st.end()                                                             # tables not yet of length 1
    for j in range(3):
        st.round_eval()
        st.bind(synth.uniform_scalars(9 + j, 1)[0])
    with pytest.raises(SbnError):
        st.round_eval()                                                      # no round left
    st.end()
    st.close()
    pc.close()
    with pytest.raises(SbnError):                                            # column index outside the vector
        SpMat(ctx, 4, 4, [0, 1], [0, 7], synth.uniform_scalars(5, 2))
    m = SpMat(ctx, 4, 4, [0, 1], [0, 3], synth.uniform_scalars(5, 2))
    with pytest.raises(SbnError):
        SpMat.mulvec([m], synth.uniform_scalars(6, 2))                        # vector shorter than the matrix is wide
    m.close()
    a = Addrs(ctx, np.zeros((1, 8), dtype=np.uint32), np.zeros((1, 8), dtype=np.uint32))
    with pytest.raises(SbnError):                                            # hash layer before the timestamps were uploaded
        a.num_cells = 8
        a.hashlayer(0, synth.uniform_scalars(7, 3), synth.uniform_scalars(8, 1)[0], synth.uniform_scalars(9, 1)[0])
    a.close()


@pytest.mark.gpu
def test_round1h_entry_points_reject_bad_shapes(ctx):
    """sbn_bsumcheck_prove with the wrong number of rounds, sbn_sumcheck_begin_r1cs / _begin_quad_r1cs with matrices that do
    not fit the vectors, sbn_bullet_end_delta on the explicit-folding path: status codes, never a crash."""
    import ctypes as C
    from spartan_bn254_b200 import SbnError, synth
    from spartan_bn254_b200.hyrax import DotProductProofGens
    from spartan_bn254_b200.lib import ProdCircuit, BatchedSumcheckState, SpMat
    from spartan_bn254_b200.transcript import Transcript
    pc = ProdCircuit(ctx, synth.uniform_scalars(1, 16))
    st = BatchedSumcheckState(ctx, [pc], 0, synth.uniform_scalars(2, 3))
    tr = Transcript(b"x")
    with pytest.raises(SbnError) as e:
        st.prove(tr._st, synth.uniform_scalars(3, 1)[0], synth.uniform_scalars(4, 1), 2)     # the tables hold 2^3 entries
    assert e.value.status == -2
    st.len = 8
    polys, r, claim, a, b, c = st.prove(tr._st, synth.uniform_scalars(3, 1)[0], synth.uniform_scalars(4, 1), 3)
    assert polys.shape == (3, 4, 4) and r.shape == (3, 4)
    with pytest.raises(SbnError):
        st.round_eval()                                                                      # no round left
    st.close()
    pc.close()
    n = 8
    mats = [SpMat(ctx, n, 16, [0, 1], [0, 3], synth.uniform_scalars(5 + i, 2)) for i in range(3)]
    z = synth.uniform_scalars(9, 16)
    with pytest.raises(SbnError) as e:
        ctx.sumcheck_begin_r1cs(mats, z, synth.uniform_scalars(10, 2))        # 2^2 != 8 rows
    assert e.value.status == -2
    with pytest.raises(SbnError):
        ctx.sumcheck_begin_r1cs(mats, z[:8], synth.uniform_scalars(10, 3))    # z shorter than the matrices are wide
    with pytest.raises(SbnError):
        ctx.sumcheck_begin_quad_r1cs(mats, synth.uniform_scalars(11, 3), synth.uniform_scalars(12, 3), z)   # rows != len(z)
    ctx.sumcheck_begin_r1cs(mats, z, synth.uniform_scalars(10, 3)).close()
    for m in mats:
        m.close()
    # explicit-folding bullet path (Q given as a point): end_delta is unsupported there, end still works
    d = DotProductProofGens(2, b"test", ctx)
    bases = d.device_bases_ext()
    av, bv = synth.uniform_scalars(13, 2), synth.uniform_scalars(14, 2)
    bl = synth.uniform_scalars(15, 3)
    st = ctx.bullet_begin(bases, d.gens_1.G[0], av, bv, bl[0])
    st.round(bl[1], bl[2])
    u = synth.uniform_scalars(16, 1)[0]
    from spartan_bn254_b200.hyrax import fr_to_int, fr_from_int
    R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
    st.fold(u, fr_from_int(pow(fr_to_int(u), -1, R)))
    with pytest.raises(SbnError) as e:
        st.end_delta(bl[1], bl[2])
    assert e.value.status == -5
    st.end()
    st.close()
