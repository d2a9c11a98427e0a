"""TEST INFRASTRUCTURE ONLY -- CPU verifier of the reference's Spark evaluation argument (SparseMatPolyEvalProof::verify).

Only tests/ and benchmark checks may import this file; the product (spartan_bn254_b200/) never does.  It restates the
reference's VERIFIER on canonical Python integers over the oracle's own Merlin transcript and Hyrax opening verifier
(oracle/bn254_oracle.c), so a proof produced through the GPU path is checked by code that shares nothing with the prover:

  reference sumcheck.rs:35-86                 SumcheckInstanceProof::verify
  reference product_tree.rs:394-537           ProductCircuitEvalProofBatched::verify
  reference sparse_mlpoly_full.rs:434-482     DerefsEvalProof::{verify_single, verify}
  reference sparse_mlpoly_full.rs:1046-1266   HashLayerProof::{verify_helper, verify}
  reference sparse_mlpoly_full.rs:1436-1521   ProductLayerProof::verify
  reference sparse_mlpoly_full.rs:1580-1651   PolyEvalNetworkProof::verify
  reference sparse_mlpoly_full.rs:1817-1845   SparseMatPolyEvalProof::verify
  reference hyrax.rs:44-52, 118-151           PolyCommitment transcript form, PolyEvalProof::{verify, verify_plain}

PARITY STATUS: unpinned by reference fixtures (the reference has no known-answer vectors and cannot be built here).
"""
import numpy as np

import oracle as orc

R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001


class VerifyError(Exception):
    pass


def _append_scalar(t, label, s):
    t.append_message(label, int(s % R).to_bytes(32, "little"))


def _append_scalars(t, label, v):
    for s in v:
        _append_scalar(t, label, s)


def _challenge(t, label):
    return int.from_bytes(t.challenge_bytes(label, 64), "little") % R


def _challenges(t, label, n):
    return [_challenge(t, label) for _ in range(n)]


def _protocol(t, name):
    t.append_message(b"protocol-name", name)


def _log2(n):
    assert n >= 1 and n & (n - 1) == 0
    return n.bit_length() - 1


def _next_pow2(n):
    return 1 if n <= 1 else 1 << (n - 1).bit_length()


def _unipoly_decompress(c, hint):
    linear = (hint - 2 * c[0] - sum(c[1:])) % R
    return [c[0], linear] + list(c[1:])


def _unipoly_eval(coeffs, r):
    acc, power = coeffs[0], r
    for c in coeffs[1:]:
        acc = (acc + power * c) % R
        power = power * r % R
    return acc


def sumcheck_verify(compressed_polys, claim, num_rounds, degree_bound, t):
    e, r = claim, []
    if len(compressed_polys) != num_rounds:
        raise VerifyError("wrong number of rounds")
    for cp in compressed_polys:
        poly = _unipoly_decompress(cp, e)
        if len(poly) - 1 != degree_bound:
            raise VerifyError("degree mismatch")
        if (poly[0] + sum(poly)) % R != e:
            raise VerifyError("sum check failed")
        t.append_message(b"poly", b"UniPoly_begin")
        for c in poly:
            _append_scalar(t, b"coeff", c)
        t.append_message(b"poly", b"UniPoly_end")
        r_i = _challenge(t, b"challenge_nextround")
        r.append(r_i)
        e = _unipoly_eval(poly, r_i)
    return e, r


def batched_verify(proof, claims_prod_vec, claims_dotp_vec, length, t):
    """product_tree.rs:394-537.  proof: object with .proof (layers with .proof.compressed_polys[*].coeffs_except_linear_term,
    .claims_prod_left, .claims_prod_right) and .claims_dotp."""
    num_layers = _log2(length)
    if len(proof.proof) != num_layers:
        raise VerifyError("wrong number of layers")
    rand = []
    claims_to_verify = list(claims_prod_vec)
    claims_to_verify_dotp = []
    nprod = len(claims_prod_vec)
    for num_rounds, i in enumerate(range(num_layers)):
        if i == num_layers - 1:
            claims_to_verify = claims_to_verify + list(claims_dotp_vec)
        coeff = _challenges(t, b"rand_coeffs_next_layer", len(claims_to_verify))
        claim = sum(a * b for a, b in zip(claims_to_verify, coeff)) % R
        layer = proof.proof[i]
        polys = [cp.coeffs_except_linear_term for cp in layer.proof.compressed_polys]
        claim_last, rand_prod = sumcheck_verify(polys, claim, num_rounds, 3, t)
        left, right = layer.claims_prod_left, layer.claims_prod_right
        if len(left) != nprod or len(right) != nprod:
            raise VerifyError("claims length")
        for j in range(nprod):
            _append_scalar(t, b"claim_prod_left", left[j])
            _append_scalar(t, b"claim_prod_right", right[j])
        if len(rand) != len(rand_prod):
            raise VerifyError("rand length")
        eq = 1
        for a, b in zip(rand, rand_prod):
            eq = eq * ((a * b + (1 - a) * (1 - b)) % R) % R
        expected = sum(coeff[j] * (left[j] * right[j] % R * eq % R) for j in range(nprod)) % R
        if i == num_layers - 1:
            dl, dr, dw = proof.claims_dotp
            for k in range(len(dl)):
                _append_scalar(t, b"claim_dotp_left", dl[k])
                _append_scalar(t, b"claim_dotp_right", dr[k])
                _append_scalar(t, b"claim_dotp_weight", dw[k])
                expected = (expected + coeff[k + nprod] * dl[k] % R * dr[k] % R * dw[k]) % R
        if expected != claim_last:
            raise VerifyError("product layer %d rejected" % i)
        r_layer = _challenge(t, b"challenge_r_layer")
        claims_to_verify = [(l + r_layer * (rr - l)) % R for l, rr in zip(left, right)]
        if i == num_layers - 1:
            dl, dr, dw = proof.claims_dotp
            for k in range(len(claims_dotp_vec) // 2):
                for v in (dl, dr, dw):
                    claims_to_verify_dotp.append((v[2 * k] + r_layer * (v[2 * k + 1] - v[2 * k])) % R)
        rand = [r_layer] + rand_prod
    return claims_to_verify, claims_to_verify_dotp, rand


def product_layer_verify(p, num_ops, num_mem_cells, evals, t):
    """sparse_mlpoly_full.rs:1436-1521."""
    _protocol(t, b"Sparse polynomial product layer proof")
    n = len(evals)
    for name, (e_init, e_read, e_write, e_audit) in ((b"row", p.eval_row), (b"col", p.eval_col)):
        if len(e_read) != n or len(e_write) != n:
            raise VerifyError("instances")
        ws = rs = 1
        for w, r in zip(e_write, e_read):
            ws, rs = ws * w % R, rs * r % R
        if e_init * ws % R != rs * e_audit % R:
            raise VerifyError("subset check " + name.decode())
        _append_scalar(t, b"claim_" + name + b"_eval_init", e_init)
        _append_scalars(t, b"claim_" + name + b"_eval_read", e_read)
        _append_scalars(t, b"claim_" + name + b"_eval_write", e_write)
        _append_scalar(t, b"claim_" + name + b"_eval_audit", e_audit)
    left_vec, right_vec = p.eval_val
    claims_dotp_circuit = []
    for i in range(n):
        if (left_vec[i] + right_vec[i]) % R != evals[i] % R:
            raise VerifyError("dotp split check")
        _append_scalar(t, b"claim_eval_dotp_left", left_vec[i])
        _append_scalar(t, b"claim_eval_dotp_right", right_vec[i])
        claims_dotp_circuit += [left_vec[i], right_vec[i]]
    claims_prod_circuit = list(p.eval_row[1]) + list(p.eval_row[2]) + list(p.eval_col[1]) + list(p.eval_col[2])
    claims_ops, claims_dotp, rand_ops = batched_verify(p.proof_ops, claims_prod_circuit, claims_dotp_circuit, num_ops, t)
    claims_prod_mem = [p.eval_row[0], p.eval_row[3], p.eval_col[0], p.eval_col[3]]
    claims_mem, _, rand_mem = batched_verify(p.proof_mem, claims_prod_mem, [], num_mem_cells, t)
    return claims_mem, rand_mem, claims_ops, claims_dotp, rand_ops


def _bound_bot_all(vals, challenges):
    vals = list(vals)
    for ch in reversed(challenges):
        vals = [(vals[2 * i] + ch * (vals[2 * i + 1] - vals[2 * i])) % R for i in range(len(vals) // 2)]
    return vals[0]


def _opening_verify_plain(proof, gens, r, Zr, comm, t):
    """PolyEvalProof::verify_plain (hyrax.rs:139-151) with the oracle's verifier.  gens = (G, h, G1) arrays; comm = (C, inf);
    proof: the package's PolyEvalProof (DotProductProofLog with GroupElements and canonical z1, z2)."""
    G, h, G1 = gens
    p = proof.proof
    gd = dict(L=[(e.xy, e.inf) for e in p.L_vec], R=[(e.xy, e.inf) for e in p.R_vec], delta=(p.delta.xy, p.delta.inf),
              beta=(p.beta.xy, p.beta.inf), z1=orc.to_mont([p.z1])[0], z2=orc.to_mont([p.z2])[0])
    C_Zr, C_inf = orc.scalar_mul(G1, 0, orc.to_mont([Zr % R])[0])          # Zr * G1 + 0 * h  (commitments.rs:122-130)
    ok = orc.poly_eval_verify(orc.EvalProof.from_dict(gd), len(r), orc.to_mont([x % R for x in r]), C_Zr, C_inf, comm[0], comm[1],
                              G, h, G1, t)
    if not ok:
        raise VerifyError("Hyrax opening rejected")


def _joint_verify(proof, evals, r, gens, comm, labels, t):
    evals = list(evals) + [0] * (_next_pow2(len(evals)) - len(evals))
    _append_scalars(t, labels[0], evals)
    challenges = _challenges(t, labels[1], _log2(len(evals)))
    joint = _bound_bot_all(evals, challenges)
    r_joint = challenges + list(r)
    _append_scalar(t, labels[2], joint)
    _opening_verify_plain(proof, gens, r_joint, joint, comm, t)


def _hash_verify_helper(rand_mem, claims, eval_ops_val, eval_ops_addr, eval_read_ts, eval_audit_ts, r, r_hash, r_ms):
    """sparse_mlpoly_full.rs:1046-1111."""
    rh2 = r_hash * r_hash % R

    def h(addr, val, ts):
        return (ts * rh2 + val * r_hash + addr) % R

    claim_init, claim_read, claim_write, claim_audit = claims
    addr = 0
    for x in rand_mem:                                   # IdentityPolynomial::evaluate (:1268-1285)
        addr = (addr * 2 + x) % R
    val = 1
    for a, b in zip(r, rand_mem):                        # EqPolynomial::evaluate (hyrax.rs:346-352)
        val = val * ((a * b + (1 - a) * (1 - b)) % R) % R
    if claim_init != (h(addr, val, 0) - r_ms) % R:
        raise VerifyError("init claim")
    if claim_audit != (h(addr, val, eval_audit_ts) - r_ms) % R:
        raise VerifyError("audit claim")
    for i in range(len(eval_ops_val)):
        if claim_read[i] != (h(eval_ops_addr[i], eval_ops_val[i], eval_read_ts[i]) - r_ms) % R:
            raise VerifyError("read claim %d" % i)
        if claim_write[i] != (h(eval_ops_addr[i], eval_ops_val[i], eval_read_ts[i] + 1) - r_ms) % R:
            raise VerifyError("write claim %d" % i)


def hash_layer_verify(p, rand, claims_row, claims_col, claims_dotp, comm, comm_derefs, gens, rx, ry, r_hash, r_ms, t):
    """sparse_mlpoly_full.rs:1113-1266.  comm: dict(comb_ops=(C, inf), comb_mem=(C, inf)); gens: dict(ops, mem, derefs) of
    (G, h, G1) arrays."""
    _protocol(t, b"Sparse polynomial hash layer proof")
    rand_mem, rand_ops = rand
    eval_row_ops_val, eval_col_ops_val = p.eval_derefs
    _protocol(t, b"Derefs evaluation proof")                                   # DerefsEvalProof::verify (:462-482)
    _joint_verify(p.proof_derefs, list(eval_row_ops_val) + list(eval_col_ops_val), rand_ops, gens["derefs"], comm_derefs,
                  (b"evals_ops_val", b"challenge_combine_n_to_one", b"joint_claim_eval"), t)
    row_addr, row_read, row_audit = p.eval_row
    col_addr, col_read, col_audit = p.eval_col
    _hash_verify_helper(rand_mem, claims_row, eval_row_ops_val, row_addr, row_read, row_audit, rx, r_hash, r_ms)
    _hash_verify_helper(rand_mem, claims_col, eval_col_ops_val, col_addr, col_read, col_audit, ry, r_hash, r_ms)
    n = len(eval_row_ops_val)
    if len(claims_dotp) != 3 * n:
        raise VerifyError("dotp claims length")
    for i in range(n):
        if claims_dotp[3 * i] != eval_row_ops_val[i] or claims_dotp[3 * i + 1] != eval_col_ops_val[i] or \
                claims_dotp[3 * i + 2] != p.eval_val[i]:
            raise VerifyError("dotp claim %d" % i)
    _joint_verify(p.proof_ops, list(row_addr) + list(row_read) + list(col_addr) + list(col_read) + list(p.eval_val), rand_ops,
                  gens["ops"], comm["comb_ops"], (b"claim_evals_ops", b"challenge_combine_n_to_one", b"joint_claim_eval_ops"), t)
    _joint_verify(p.proof_mem, [row_audit, col_audit], rand_mem, gens["mem"], comm["comb_mem"],
                  (b"claim_evals_mem", b"challenge_combine_two_to_one", b"joint_claim_eval_mem"), t)


def _append_poly_commitment(t, label, comm):
    C, inf = comm
    t.append_message(label, b"poly_commitment_begin")
    for pt, i in zip(C, inf):
        t.append_message(b"poly_commitment_share", orc.compress(pt, int(i)))
    t.append_message(label, b"poly_commitment_end")


def sparse_mat_poly_eval_verify(proof, comm, rx, ry, evals, gens, t):
    """SparseMatPolyEvalProof::verify (:1817-1845) + PolyEvalNetworkProof::verify (:1580-1651).
    comm: dict(batch_size, num_ops, num_mem_cells, comb_ops=(C, inf), comb_mem=(C, inf))."""
    _protocol(t, b"Sparse polynomial evaluation proof")
    rx, ry = list(rx), list(ry)
    if len(rx) < len(ry):
        rx = [0] * (len(ry) - len(rx)) + rx
    elif len(ry) < len(rx):
        ry = [0] * (len(rx) - len(ry)) + ry
    nz, num_mem_cells = comm["num_ops"], comm["num_mem_cells"]
    if 1 << len(rx) != num_mem_cells:
        raise VerifyError("memory size")
    comm_derefs = (proof.comm_derefs.C, proof.comm_derefs.inf)
    t.append_message(b"derefs_commitment", b"begin_derefs_commitment")
    _append_poly_commitment(t, b"comm_poly_row_col_ops_val", comm_derefs)
    t.append_message(b"derefs_commitment", b"end_derefs_commitment")
    r_hash, r_ms = _challenges(t, b"challenge_r_hash", 2)
    net = proof.poly_eval_network_proof
    _protocol(t, b"Sparse polynomial evaluation proof")
    n = len(evals)
    claims_mem, rand_mem, claims_ops, claims_dotp, rand_ops = product_layer_verify(net.proof_prod_layer, _next_pow2(nz),
                                                                                  num_mem_cells, evals, t)
    if len(claims_mem) != 4 or len(claims_ops) != 4 * n:
        raise VerifyError("claims shape")
    row, col = claims_ops[: 2 * n], claims_ops[2 * n:]
    claims_row = (claims_mem[0], row[:n], row[n:], claims_mem[1])
    claims_col = (claims_mem[2], col[:n], col[n:], claims_mem[3])
    hash_layer_verify(net.proof_hash_layer, (rand_mem, rand_ops), claims_row, claims_col, claims_dotp, comm, comm_derefs, gens,
                      rx, ry, r_hash, r_ms, t)
    return True
