"""Spark evaluation argument end to end (R1CSEvalProof = SparseMatPolyEvalProof, sparse_mlpoly_full.rs:1694-1845): the
GPU-backed prover's proof must be accepted by the oracle's independent restatement of the reference's verifier, and a
tampered proof / claim must be rejected."""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001


def _matrices(seed, nvx, nvy, nnz, batch):
    rnd = random.Random(seed)
    from spartan_bn254_b200.spark import SparseMatPolynomial
    polys = []
    for b in range(batch):
        k = nnz - b          # ragged: the instances have different numbers of non-zeros
        entries = [(rnd.randrange(1 << nvx), rnd.randrange(1 << nvy), rnd.randrange(1, R)) for _ in range(k)]
        polys.append(SparseMatPolynomial(nvx, nvy, entries))
    return polys


def _gens_arrays(g):
    return (g.gens.gens_n.G, g.gens.gens_n.h, g.gens.gens_1.G[0])


def _prove(ctx, polys, nvx, nvy, seed):
    from spartan_bn254_b200.spark import SparseMatPolynomial, SparseMatPolyCommitmentGens, SparseMatPolyEvalProof, multi_commit
    from spartan_bn254_b200.transcript import Transcript, RandomTape
    rnd = random.Random(seed)
    nnz = max(len(p.M) for p in polys)
    gens = SparseMatPolyCommitmentGens(b"gens_r1cs_eval", nvx, nvy, nnz, len(polys), ctx)
    comm, dense = multi_commit(ctx, polys, gens)
    rx = [rnd.randrange(R) for _ in range(nvx)]
    ry = [rnd.randrange(R) for _ in range(nvy)]
    evals = SparseMatPolynomial.multi_evaluate(polys, rx, ry)
    from spartan_bn254_b200.spark import equalize
    assert dense.multi_evaluate(*equalize(rx, ry)) == evals          # the device-side multi_evaluate agrees with the host one
    proof = SparseMatPolyEvalProof.prove(dense, rx, ry, evals, gens, Transcript(b"spark"), RandomTape(b"proof", 777))
    dense.close()
    return proof, comm, gens, rx, ry, evals


def _verify(orc, proof, comm, gens, rx, ry, evals):
    import spark_model as sm
    cd = dict(batch_size=comm.batch_size, num_ops=comm.num_ops, num_mem_cells=comm.num_mem_cells,
              comb_ops=(comm.comm_comb_ops.C, comm.comm_comb_ops.inf), comb_mem=(comm.comm_comb_mem.C, comm.comm_comb_mem.inf))
    g = dict(ops=_gens_arrays(gens.gens_ops), mem=_gens_arrays(gens.gens_mem), derefs=_gens_arrays(gens.gens_derefs))
    return sm.sparse_mat_poly_eval_verify(proof, cd, rx, ry, evals, g, orc.Transcript(b"spark"))


@pytest.mark.parametrize("nvx,nvy,nnz,batch", [(3, 3, 7, 1), (4, 4, 30, 3), (5, 6, 100, 3), (6, 4, 64, 2)])
def test_spark_eval_proof_is_accepted(ctx, orc, nvx, nvy, nnz, batch):
    polys = _matrices(1000 + nnz, nvx, nvy, nnz, batch)
    proof, comm, gens, rx, ry, evals = _prove(ctx, polys, nvx, nvy, 5)
    assert _verify(orc, proof, comm, gens, rx, ry, evals)


def test_spark_eval_proof_rejections(ctx, orc):
    import spark_model as sm
    polys = _matrices(4242, 4, 4, 25, 3)
    proof, comm, gens, rx, ry, evals = _prove(ctx, polys, 4, 4, 6)
    assert _verify(orc, proof, comm, gens, rx, ry, evals)
    # a different claimed evaluation
    bad = list(evals)
    bad[1] = (bad[1] + 1) % R
    with pytest.raises(sm.VerifyError):
        _verify(orc, proof, comm, gens, rx, ry, bad)
    # a tampered hash-layer evaluation
    hl = proof.poly_eval_network_proof.proof_hash_layer
    keep = hl.eval_val[0]
    hl.eval_val[0] = (keep + 1) % R
    with pytest.raises(sm.VerifyError):
        _verify(orc, proof, comm, gens, rx, ry, evals)
    hl.eval_val[0] = keep
    # a tampered product-layer claim
    pl = proof.poly_eval_network_proof.proof_prod_layer
    keep = pl.proof_ops.proof[0].claims_prod_left[0]
    pl.proof_ops.proof[0].claims_prod_left[0] = (keep + 1) % R
    with pytest.raises(sm.VerifyError):
        _verify(orc, proof, comm, gens, rx, ry, evals)
    pl.proof_ops.proof[0].claims_prod_left[0] = keep
    # a different evaluation point
    with pytest.raises(sm.VerifyError):
        _verify(orc, proof, comm, gens, rx[::-1], ry, evals)
    assert _verify(orc, proof, comm, gens, rx, ry, evals)
