"""TEST INFRASTRUCTURE ONLY -- CPU verifier of the reference's R1CS-satisfiability proof and of SNARK::verify.

Only tests/ and benchmark checks may import this file; the product (spartan_bn254_b200/) never does.  Restated on canonical
Python integers, with every group operation done by the oracle's C code (oracle/bn254_oracle.c: MSM, compression, the
Hyrax opening verifier) and the oracle's own Merlin transcript:

  reference nizk/mod.rs:58-82, 126-150, 244-284, 366-401   Knowledge / Equality / Product / DotProduct proof verifiers
  reference sumcheck.rs:366-457                            ZKSumcheckInstanceProof::verify
  reference r1csproof.rs:463-619                           R1CSProof::verify
  reference r1cs.rs:355-363, sparse_mlpoly_full.rs:700-708 commitment transcript form
  reference snark.rs:487-526                               SNARK::verify

PARITY STATUS: unpinned by reference fixtures (the reference has no known-answer vectors and cannot be built here).
"""
import numpy as np

import oracle as orc
import spark_model as sm
from spark_model import R, VerifyError, _append_scalar, _append_scalars, _challenge, _challenges, _protocol


class Gens:
    """MultiCommitGens as plain arrays: G (n x 8), h (8)."""

    def __init__(self, G, h):
        self.G = np.ascontiguousarray(G, dtype=np.uint64).reshape(-1, 8)
        self.h = np.ascontiguousarray(h, dtype=np.uint64).reshape(8)
        self.n = self.G.shape[0]


def _pt(g):
    """(xy, inf) of a package GroupElement or of a tuple."""
    return (g.xy, g.inf) if hasattr(g, "xy") else g


def _msm(points, scalars):
    """sum s_i P_i; points: list of (xy, inf); scalars canonical ints."""
    pts = np.stack([np.asarray(p[0], dtype=np.uint64).reshape(8) for p in points])
    inf = np.array([p[1] for p in points], dtype=np.uint8)
    return orc.msm(pts, inf, orc.to_mont([s % R for s in scalars]))


def _commit(gens, scalars, blind):
    assert gens.n == len(scalars)
    pts = [(gens.G[i], 0) for i in range(gens.n)] + [(gens.h, 0)]
    return _msm(pts, list(scalars) + [blind])


def _eq(a, b):
    return a[1] == b[1] and (a[1] == 1 or np.array_equal(np.asarray(a[0]).reshape(8), np.asarray(b[0]).reshape(8)))


def _append_point(t, label, p):
    t.append_message(label, orc.compress(np.asarray(p[0], dtype=np.uint64).reshape(8), int(p[1])))


def knowledge_verify(p, gens_n, t, C):
    _protocol(t, b"knowledge proof")
    _append_point(t, b"C", C)
    _append_point(t, b"alpha", _pt(p.alpha))
    c = _challenge(t, b"c")
    if not _eq(_commit(gens_n, [p.z1], p.z2), _msm([C, _pt(p.alpha)], [c, 1])):
        raise VerifyError("knowledge proof")


def equality_verify(p, gens_n, t, C1, C2):
    _protocol(t, b"equality proof")
    _append_point(t, b"C1", C1)
    _append_point(t, b"C2", C2)
    _append_point(t, b"alpha", _pt(p.alpha))
    c = _challenge(t, b"c")
    rhs = _msm([C1, C2, _pt(p.alpha)], [c, -c, 1])                      # c * (C1 - C2) + alpha
    if not _eq(_msm([(gens_n.h, 0)], [p.z]), rhs):
        raise VerifyError("equality proof")


def product_verify(p, gens_n, t, X, Y, Z):
    _protocol(t, b"product proof")
    for label, g in ((b"X", X), (b"Y", Y), (b"Z", Z), (b"alpha", _pt(p.alpha)), (b"beta", _pt(p.beta)), (b"delta", _pt(p.delta))):
        _append_point(t, label, g)
    z1, z2, z3, z4, z5 = p.z
    c = _challenge(t, b"c")

    def check(P, Xc, gens, a, b):
        return _eq(_msm([P, Xc], [1, c]), _commit(gens, [a], b))

    gens_X = Gens(np.asarray(X[0]).reshape(1, 8), gens_n.h)
    if X[1]:
        raise VerifyError("product proof: X is the identity")
    if not (check(_pt(p.alpha), X, gens_n, z1, z2) and check(_pt(p.beta), Y, gens_n, z3, z4) and
            check(_pt(p.delta), Z, gens_X, z3, z5)):
        raise VerifyError("product proof")


def dotproduct_verify(p, gens_1, gens_n, t, a, Cx, Cy):
    assert gens_n.n == len(a) and gens_1.n == 1
    _protocol(t, b"dot product proof")
    _append_point(t, b"Cx", Cx)
    _append_point(t, b"Cy", Cy)
    _append_scalars(t, b"a", a)
    _append_point(t, b"delta", _pt(p.delta))
    _append_point(t, b"beta", _pt(p.beta))
    c = _challenge(t, b"c")
    ok = _eq(_msm([Cx, _pt(p.delta)], [c, 1]), _commit(gens_n, p.z, p.z_delta))
    dot = sum(x * y for x, y in zip(p.z, a)) % R
    ok = ok and _eq(_msm([Cy, _pt(p.beta)], [c, 1]), _commit(gens_1, [dot], p.z_beta))
    if not ok:
        raise VerifyError("dot product proof")


def zk_sumcheck_verify(p, comm_claim, num_rounds, degree_bound, gens_1, gens_n, t):
    """sumcheck.rs:366-457."""
    if len(p.comm_polys) != num_rounds or len(p.proofs) != num_rounds:
        raise VerifyError("wrong number of rounds")
    comm_claim_per_round = comm_claim
    r = []
    for i in range(num_rounds):
        comm_poly, comm_eval = _pt(p.comm_polys[i]), _pt(p.comm_evals[i])
        _append_point(t, b"comm_poly", comm_poly)
        r_i = _challenge(t, b"challenge_nextround")
        _append_point(t, b"comm_claim_per_round", comm_claim_per_round)
        _append_point(t, b"comm_eval", comm_eval)
        w = _challenges(t, b"combine_two_claims_to_one", 2)
        comm_target = _msm([comm_claim_per_round, comm_eval], w)
        a_sc = [2] + [1] * degree_bound
        a_eval = [1]
        for _ in range(degree_bound):
            a_eval.append(a_eval[-1] * r_i % R)
        a = [(w[0] * s + w[1] * e) % R for s, e in zip(a_sc, a_eval)]
        dotproduct_verify(p.proofs[i], gens_1, gens_n, t, a, comm_poly, comm_target)
        comm_claim_per_round = comm_eval
        r.append(r_i)
    return _pt(p.comm_evals[-1]), r


def r1cs_proof_verify(p, num_vars, num_cons, inputs, evals, t, gens):
    """r1csproof.rs:463-619.  gens: dict(gens_1, gens_3, gens_4: Gens; pc: (G, h, G1) of gens_pc; pc_1: Gens of gens_pc.gens_1).
    Returns (rx, ry)."""
    _protocol(t, b"R1CS proof")
    _append_scalars(t, b"input", inputs)
    comm_vars = (p.comm_vars.C, p.comm_vars.inf)
    sm._append_poly_commitment(t, b"poly_commitment", comm_vars)
    num_rounds_x, num_rounds_y = sm._log2(num_cons), sm._log2(2 * num_vars)
    tau = _challenges(t, b"challenge_tau", num_rounds_x)
    g1, g3, g4 = gens["gens_1"], gens["gens_3"], gens["gens_4"]
    claim_phase1 = _commit(g1, [0], 0)
    comm_claim_post_phase1, rx = zk_sumcheck_verify(p.sc_proof_phase1, claim_phase1, num_rounds_x, 3, g1, g4, t)
    comm_Az, comm_Bz, comm_Cz, comm_prod = [_pt(x) for x in p.claims_phase2]
    pok_Cz, proof_prod = p.pok_claims_phase2
    knowledge_verify(pok_Cz, g1, t, comm_Cz)
    product_verify(proof_prod, g1, t, comm_Az, comm_Bz, comm_prod)
    _append_point(t, b"comm_Az_claim", comm_Az)
    _append_point(t, b"comm_Bz_claim", comm_Bz)
    _append_point(t, b"comm_Cz_claim", comm_Cz)
    _append_point(t, b"comm_prod_Az_Bz_claims", comm_prod)
    taus_bound_rx = 1
    for a, b in zip(rx, tau):
        taus_bound_rx = taus_bound_rx * ((a * b + (1 - a) * (1 - b)) % R) % R
    expected1 = _msm([comm_prod, comm_Cz], [taus_bound_rx, -taus_bound_rx])
    equality_verify(p.proof_eq_sc_phase1, g1, t, expected1, comm_claim_post_phase1)
    r_A, r_B, r_C = _challenge(t, b"challenge_Az"), _challenge(t, b"challenge_Bz"), _challenge(t, b"challenge_Cz")
    comm_claim_phase2 = _msm([comm_Az, comm_Bz, comm_Cz], [r_A, r_B, r_C])
    comm_claim_post_phase2, ry = zk_sumcheck_verify(p.sc_proof_phase2, comm_claim_phase2, num_rounds_y, 2, g1, g3, t)
    # PolyEvalProof::verify (hyrax.rs:118-137)
    G, h, G1 = gens["pc"]
    pe = p.proof_eval_vars_at_ry.proof
    gd = dict(L=[(e.xy, e.inf) for e in pe.L_vec], R=[(e.xy, e.inf) for e in pe.R_vec], delta=(pe.delta.xy, pe.delta.inf),
              beta=(pe.beta.xy, pe.beta.inf), z1=orc.to_mont([pe.z1])[0], z2=orc.to_mont([pe.z2])[0])
    cv = _pt(p.comm_vars_at_ry)
    if not orc.poly_eval_verify(orc.EvalProof.from_dict(gd), len(ry) - 1, orc.to_mont([x % R for x in ry[1:]]), cv[0], cv[1],
                                comm_vars[0], comm_vars[1], G, h, G1, t):
        raise VerifyError("witness opening rejected")
    # poly_input_eval: (1, inputs...) as a polynomial over the first variables' half (r1csproof.rs:577-588)
    ry_evals = _eq_evals(ry[1:])
    poly_input_eval = ry_evals[0]
    for i, inp in enumerate(inputs):
        poly_input_eval = (poly_input_eval + inp * ry_evals[i + 1]) % R
    pc1 = gens["pc_1"]
    comm_eval_Z_at_ry = _msm([cv, _commit(pc1, [poly_input_eval], 0)], [(1 - ry[0]) % R, ry[0]])
    eval_A, eval_B, eval_C = evals
    coeff = (r_A * eval_A + r_B * eval_B + r_C * eval_C) % R
    expected2 = _msm([comm_eval_Z_at_ry], [coeff])
    equality_verify(p.proof_eq_sc_phase2, g1, t, expected2, comm_claim_post_phase2)
    return rx, ry


def _eq_evals(r):
    ell = len(r)
    evals = [1] * (1 << ell)
    size = 1
    for j in range(ell):
        size *= 2
        for i in range(size - 1, -1, -2):
            s = evals[i // 2]
            evals[i] = s * r[j] % R
            evals[i - 1] = (s - evals[i]) % R
    return evals


def _append_u64(t, label, v):
    t.append_message(label, int(v).to_bytes(8, "little"))


def snark_verify(proof, comm, inputs, gens_sat, gens_eval, t):
    """snark.rs:487-526.  comm: dict(num_cons, num_vars, num_inputs, batch_size, num_ops, num_mem_cells, comb_ops, comb_mem)."""
    _protocol(t, b"Spartan SNARK proof")
    for k in ("num_cons", "num_vars", "num_inputs", "batch_size", "num_ops", "num_mem_cells"):
        _append_u64(t, k.encode(), comm[k])
    sm._append_poly_commitment(t, b"comm_comb_ops", comm["comb_ops"])
    sm._append_poly_commitment(t, b"comm_comb_mem", comm["comb_mem"])
    if len(inputs) != comm["num_inputs"]:
        raise VerifyError("inputs length")
    rx, ry = r1cs_proof_verify(proof.r1cs_sat_proof, comm["num_vars"], comm["num_cons"], inputs, proof.inst_evals, t, gens_sat)
    sm.sparse_mat_poly_eval_verify(proof.r1cs_eval_proof, comm, rx, ry, list(proof.inst_evals), gens_eval, t)
    return True
