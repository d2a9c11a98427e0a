"""TEST INFRASTRUCTURE ONLY -- CPU restatement (Python integers) of the reference's product-layer argument.

Only tests/ and benchmark baselines may import this file; the product (spartan_bn254_b200/) never does.  It follows the
reference line by line on plain lists of canonical integers mod r and drives the oracle's own Merlin transcript
(oracle/bn254_oracle.c STROBE-128 restatement, checked against Merlin's published test vector):

  reference product_tree.rs:21-65     ProductCircuit::{compute_layer, new, evaluate}   -> ProductCircuit
  reference product_tree.rs:74-86     DotProductCircuit::evaluate                      -> dotp_evaluate()
  reference hyrax.rs:355-369          EqPolynomial::evals                              -> eq_evals()
  reference hyrax.rs:195-203          DensePolynomial::bound_poly_var_top              -> bind_top()
  reference unipoly.rs:28-59,81-97    UniPoly::{from_evals, evaluate, compress}        -> unipoly_*()
  reference unipoly.rs:119-127        UniPoly::append_to_transcript                    -> unipoly_append()
  reference sumcheck.rs:165-330       SumcheckInstanceProof::prove_cubic_batched       -> prove_cubic_batched()
  reference product_tree.rs:251-392   ProductCircuitEvalProofBatched::prove            -> prove_batched()

  reference sparse_mlpoly_full.rs:212-243 AddrTimestamps::new                          -> addr_timestamps()
  reference sparse_mlpoly_full.rs:745-798 Layers::build_hash_layer                     -> build_hash_layer()

PARITY STATUS: unpinned by reference fixtures (the reference has no known-answer tests for this argument and cannot be
built here); pinned by exact field arithmetic (every quantity is a field element, so any correct implementation produces
the same integers) and by the verifier restated in the package accepting the proofs.
"""
R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001


class ProductCircuit:
    def __init__(self, poly):
        n = len(poly)
        assert n >= 2 and n & (n - 1) == 0
        self.left_vec = [list(poly[: n // 2])]
        self.right_vec = [list(poly[n // 2:])]
        num_layers = n.bit_length() - 1
        for i in range(num_layers - 1):
            left, right = self.left_vec[i], self.right_vec[i]
            length = len(left) + len(right)
            self.left_vec.append([left[j] * right[j] % R for j in range(length // 4)])
            self.right_vec.append([left[j] * right[j] % R for j in range(length // 4, length // 2)])

    def evaluate(self):
        assert len(self.left_vec[-1]) == 1 and len(self.right_vec[-1]) == 1
        return self.left_vec[-1][0] * self.right_vec[-1][0] % R


def dotp_evaluate(left, right, weight):
    return sum(a * b % R * c for a, b, c in zip(left, right, weight)) % R


def eq_evals(r):
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


def bind_top(Z, r):
    n = len(Z) // 2
    return [(Z[i] + r * (Z[i + n] - Z[i])) % R for i in range(n)]


def unipoly_from_evals(e):
    two_inv, six_inv = pow(2, -1, R), pow(6, -1, R)
    assert len(e) == 4
    d = e[0]
    a = six_inv * (e[3] - e[2] - e[2] - e[2] + e[1] + e[1] + e[1] - e[0]) % R
    b = two_inv * (e[0] + e[0] - 5 * e[1] + 4 * e[2] - e[3]) % R
    c = (e[1] - d - a - b) % R
    return [d, c, b, a]


def unipoly_evaluate(coeffs, r):
    acc, power = coeffs[0], r
    for c in coeffs[1:]:
        acc = (acc + power * c) % R
        power = power * r % R
    return acc


def _append_scalar(t, label, s):
    t.append_message(label, int(s).to_bytes(32, "little"))          # transcript.rs:46-48, scalar.rs:75-84


def _challenge_scalar(t, label):
    return int.from_bytes(t.challenge_bytes(label, 64), "little") % R   # transcript.rs:56-67


def unipoly_append(t, coeffs):
    t.append_message(b"poly", b"UniPoly_begin")
    for c in coeffs:
        _append_scalar(t, b"coeff", c)
    t.append_message(b"poly", b"UniPoly_end")


def _cubic_evals(A, B, C):
    n = len(A) // 2
    e0 = e2 = e3 = 0
    for i in range(n):
        e0 += A[i] * B[i] % R * C[i]
        a2, b2, c2 = 2 * A[n + i] - A[i], 2 * B[n + i] - B[i], 2 * C[n + i] - C[i]
        e2 += a2 * b2 % R * c2
        a3, b3, c3 = a2 + A[n + i] - A[i], b2 + B[n + i] - B[i], c2 + C[n + i] - C[i]
        e3 += a3 * b3 % R * c3
    return e0 % R, e2 % R, e3 % R


def prove_cubic_batched(claim, num_rounds, par, seq, coeffs, t):
    """par = (list of A, list of B, C); seq = (list of A, list of B, list of C); tables are replaced as they are bound."""
    A_par, B_par, C_par = par
    A_seq, B_seq, C_seq = seq
    e, r, polys = claim, [], []
    for _ in range(num_rounds):
        evals = [_cubic_evals(a, b, C_par) for a, b in zip(A_par, B_par)]
        evals += [_cubic_evals(a, b, c) for a, b, c in zip(A_seq, B_seq, C_seq)]
        c0 = sum(ev[0] * co for ev, co in zip(evals, coeffs)) % R
        c2 = sum(ev[1] * co for ev, co in zip(evals, coeffs)) % R
        c3 = sum(ev[2] * co for ev, co in zip(evals, coeffs)) % R
        poly = unipoly_from_evals([c0, (e - c0) % R, c2, c3])
        unipoly_append(t, poly)
        r_j = _challenge_scalar(t, b"challenge_nextround")
        r.append(r_j)
        A_par = [bind_top(a, r_j) for a in A_par]
        B_par = [bind_top(b, r_j) for b in B_par]
        C_par = bind_top(C_par, r_j)
        A_seq = [bind_top(a, r_j) for a in A_seq]
        B_seq = [bind_top(b, r_j) for b in B_seq]
        C_seq = [bind_top(c, r_j) for c in C_seq]
        e = unipoly_evaluate(poly, r_j)
        polys.append([poly[0]] + poly[2:])                         # compress: drop the linear term
    claims_prod = ([a[0] for a in A_par], [b[0] for b in B_par], C_par[0])
    claims_dotp = ([a[0] for a in A_seq], [b[0] for b in B_seq], [c[0] for c in C_seq])
    return polys, r, claims_prod, claims_dotp


def prove_batched(prod_polys, dotp, t):
    """prod_polys: list of polynomials (lists of ints); dotp: list of (left, right, weight).  Returns a dict of plain lists:
    layers[k] = (compressed round polynomials, claims_prod_left, claims_prod_right), claims_dotp, rand."""
    circuits = [ProductCircuit(p) for p in prod_polys]
    num_layers = len(circuits[0].left_vec)
    claims_to_verify = [c.evaluate() for c in circuits]
    claims_dotp_final = ([], [], [])
    layers, rand = [], []
    for layer_id in reversed(range(num_layers)):
        length = len(circuits[0].left_vec[layer_id]) + len(circuits[0].right_vec[layer_id])
        C_par = eq_evals(rand)
        assert len(C_par) == length // 2
        num_rounds = (length // 2).bit_length() - 1
        seqA, seqB, seqC = [], [], []
        if layer_id == 0 and dotp:
            for left, right, weight in dotp:
                claims_to_verify.append(dotp_evaluate(left, right, weight))
                assert len(left) == len(right) == len(weight) == length // 2
                seqA.append(list(left)); seqB.append(list(right)); seqC.append(list(weight))
        coeff_vec = [_challenge_scalar(t, b"rand_coeffs_next_layer") for _ in claims_to_verify]
        claim = sum(a * b for a, b in zip(claims_to_verify, coeff_vec)) % R
        polys, rand_prod, claims_prod, claims_dotp = prove_cubic_batched(
            claim, num_rounds, ([c.left_vec[layer_id] for c in circuits], [c.right_vec[layer_id] for c in circuits], C_par),
            (seqA, seqB, seqC), coeff_vec, t)
        left, right, _ = claims_prod
        for i in range(len(circuits)):
            _append_scalar(t, b"claim_prod_left", left[i])
            _append_scalar(t, b"claim_prod_right", right[i])
        if layer_id == 0 and dotp:
            dl, dr, dw = claims_dotp
            for i in range(len(dotp)):
                _append_scalar(t, b"claim_dotp_left", dl[i])
                _append_scalar(t, b"claim_dotp_right", dr[i])
                _append_scalar(t, b"claim_dotp_weight", dw[i])
            claims_dotp_final = claims_dotp
        r_layer = _challenge_scalar(t, b"challenge_r_layer")
        claims_to_verify = [(l + r_layer * (rr - l)) % R for l, rr in zip(left, right)]
        rand = [r_layer] + rand_prod
        layers.append((polys, left, right))
    return {"layers": layers, "claims_dotp": claims_dotp_final, "rand": rand,
            "claims_prod": [c.evaluate() for c in circuits]}


def addr_timestamps(num_cells, ops_addr):
    """ops_addr: list of address lists (one per instance).  Returns (read_ts per instance, audit_ts)."""
    audit_ts = [0] * num_cells
    read_ts_vec = []
    for inst in ops_addr:
        read_ts = [0] * len(inst)
        for i, addr in enumerate(inst):
            assert addr < num_cells
            r_ts = audit_ts[addr]
            read_ts[i] = r_ts
            audit_ts[addr] = r_ts + 1
        read_ts_vec.append(read_ts)
    return read_ts_vec, audit_ts


def build_hash_layer(eval_table, addrs_vec, derefs_vec, read_ts_vec, audit_ts, r_hash, r_multiset_check):
    r_hash_sqr = r_hash * r_hash % R

    def h(addr, val, ts):
        return (ts * r_hash_sqr + val * r_hash + addr) % R

    init = [(h(i, eval_table[i], 0) - r_multiset_check) % R for i in range(len(eval_table))]
    audit = [(h(i, eval_table[i], audit_ts[i]) - r_multiset_check) % R for i in range(len(eval_table))]
    reads, writes = [], []
    for addrs, derefs, read_ts in zip(addrs_vec, derefs_vec, read_ts_vec):
        reads.append([(h(a, d, t) - r_multiset_check) % R for a, d, t in zip(addrs, derefs, read_ts)])
        writes.append([(h(a, d, t + 1) - r_multiset_check) % R for a, d, t in zip(addrs, derefs, read_ts)])
    return init, reads, writes, audit
