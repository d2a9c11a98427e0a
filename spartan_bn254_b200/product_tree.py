"""Host mirror of the reference's product-layer argument over the GPU tables.

  reference unipoly.rs:26-127          UniPoly / CompressedUniPoly
  reference sumcheck.rs:24-86          SumcheckInstanceProof::verify
  reference sumcheck.rs:165-330        SumcheckInstanceProof::prove_cubic_batched
  reference product_tree.rs:14-65      ProductCircuit
  reference product_tree.rs:67-106     DotProductCircuit
  reference product_tree.rs:251-537    ProductCircuitEvalProofBatched::{prove, verify}

The tables (every layer of every product circuit, the eq table, the dot-product circuits) live in HBM behind
sbn_prodcircuit / sbn_bsumcheck; per round the GPU returns three field elements per instance, and everything that
touches the Fiat-Shamir transcript (combination with the random coefficients, the cubic round polynomial, the
challenges) runs here on canonical Python integers, in the reference's order, so the transcript is the reference's.
"""
from .hyrax import R_MOD, fr_to_int, fr_from_int, fr_vec_from_ints, fr_vec_to_ints
from .lib import ProdCircuit, BatchedSumcheckState

_TWO_INV = pow(2, -1, R_MOD)
_SIX_INV = pow(6, -1, R_MOD)


class UniPoly:
    """unipoly.rs:14-105: coefficients, lowest degree first."""

    def __init__(self, coeffs):
        self.coeffs = [c % R_MOD for c in coeffs]

    @staticmethod
    def from_evals(evals):
        e = evals
        assert len(e) in (3, 4)
        if len(e) == 3:
            c = e[0]
            a = _TWO_INV * (e[2] - e[1] - e[1] + c) % R_MOD
            b = (e[1] - c - a) % R_MOD
            return UniPoly([c, b, a])
        d = e[0]
        a = _SIX_INV * (e[3] - 3 * e[2] + 3 * e[1] - e[0]) % R_MOD
        b = _TWO_INV * (2 * e[0] - 5 * e[1] + 4 * e[2] - e[3]) % R_MOD
        c = (e[1] - d - a - b) % R_MOD
        return UniPoly([d, c, b, a])

    def degree(self):
        return len(self.coeffs) - 1

    def eval_at_zero(self):
        return self.coeffs[0]

    def eval_at_one(self):
        return sum(self.coeffs) % R_MOD

    def evaluate(self, r):
        acc, power = self.coeffs[0], r
        for c in self.coeffs[1:]:
            acc = (acc + power * c) % R_MOD
            power = power * r % R_MOD
        return acc

    def compress(self):
        return CompressedUniPoly([self.coeffs[0]] + self.coeffs[2:])

    def append_to_transcript(self, label, transcript):           # unipoly.rs:119-127
        transcript.append_message(label, b"UniPoly_begin")
        for c in self.coeffs:
            transcript.append_scalar(b"coeff", c)
        transcript.append_message(label, b"UniPoly_end")


class CompressedUniPoly:
    def __init__(self, coeffs_except_linear_term):
        self.coeffs_except_linear_term = list(coeffs_except_linear_term)

    def decompress(self, hint):                                    # unipoly.rs:103-116
        c = self.coeffs_except_linear_term
        linear = (hint - 2 * c[0] - sum(c[1:])) % R_MOD
        return UniPoly([c[0], linear] + c[1:])


class SumcheckInstanceProof:
    def __init__(self, compressed_polys):
        self.compressed_polys = compressed_polys

    def verify(self, claim, num_rounds, degree_bound, transcript):   # sumcheck.rs:35-86
        e, r = claim, []
        if len(self.compressed_polys) != num_rounds:
            raise ValueError("wrong number of rounds")
        for cp in self.compressed_polys:
            poly = cp.decompress(e)
            if poly.degree() != degree_bound:
                raise ValueError("degree mismatch")
            if (poly.eval_at_zero() + poly.eval_at_one()) % R_MOD != e:
                raise ValueError("sum check failed")
            poly.append_to_transcript(b"poly", transcript)
            r_i = transcript.challenge_scalar(b"challenge_nextround")
            r.append(r_i)
            e = poly.evaluate(r_i)
        return e, r

    @staticmethod
    def prove_cubic_batched(claim, num_rounds, state, coeffs, transcript):
        """sumcheck.rs:165-330 with the tables behind `state` (BatchedSumcheckState); returns
        (proof, r, (A_par, B_par, C_par0), (A_seq, B_seq, C_seq)) as canonical ints."""
        P, S = state.P, state.S
        if getattr(transcript, "_st", None) is not None:
            # native Merlin state: the round loop runs inside the library (sbn_bsumcheck_prove), one call per layer
            pm, rm, _e, a, b, c = state.prove(transcript._st, fr_from_int(claim), fr_vec_from_ints(list(coeffs)), num_rounds)
            co = fr_vec_to_ints(pm.reshape(-1, 4))
            polys = [UniPoly(co[4 * j: 4 * j + 4]).compress() for j in range(num_rounds)]
            r = fr_vec_to_ints(rm)
            a, b, c = fr_vec_to_ints(a), fr_vec_to_ints(b), fr_vec_to_ints(c)
            return (SumcheckInstanceProof(polys), r, (a[:P], b[:P], c[0]), (a[P:], b[P:], c[1:]))
        e, r, polys = claim, [], []
        for _ in range(num_rounds):
            ev = fr_vec_to_ints(state.round_eval().reshape(-1, 4))  # (P + S) x 3: e0, e2, e3 per instance (:201-271)
            comb = [sum(ev[3 * i + k] * coeffs[i] for i in range(P + S)) % R_MOD for k in range(3)]
            poly = UniPoly.from_evals([comb[0], (e - comb[0]) % R_MOD, comb[1], comb[2]])   # :273-284
            poly.append_to_transcript(b"poly", transcript)
            r_j = transcript.challenge_scalar(b"challenge_nextround")
            r.append(r_j)
            state.bind(fr_from_int(r_j))                            # :293-306
            e = poly.evaluate(r_j)
            polys.append(poly.compress())
        a, b, c = state.end()
        a, b, c = fr_vec_to_ints(a), fr_vec_to_ints(b), fr_vec_to_ints(c)
        return (SumcheckInstanceProof(polys), r, (a[:P], b[:P], c[0]), (a[P:], b[P:], c[1:]))


class ProductCircuit:
    """product_tree.rs:14-65; the layers are built and kept on the GPU."""

    def __init__(self, ctx, poly_mont):
        self.gpu = ProdCircuit(ctx, poly_mont)
        self.len = self.gpu.len
        self.num_layers = self.gpu.num_layers

    def evaluate(self):
        return fr_to_int(self.gpu.evaluate())

    def close(self):
        self.gpu.close()


class DotProductCircuit:
    """product_tree.rs:67-106 (left, right, weight as Montgomery uint64[n,4] arrays)."""

    def __init__(self, left, right, weight):
        assert len(left) == len(right) == len(weight)
        self.left, self.right, self.weight = left, right, weight

    def evaluate(self):
        l, r, w = fr_vec_to_ints(self.left), fr_vec_to_ints(self.right), fr_vec_to_ints(self.weight)
        return sum(a * b % R_MOD * c for a, b, c in zip(l, r, w)) % R_MOD

    def split(self):
        idx = len(self.left) // 2
        assert idx * 2 == len(self.left)
        return (DotProductCircuit(self.left[:idx], self.right[:idx], self.weight[:idx]),
                DotProductCircuit(self.left[idx:], self.right[idx:], self.weight[idx:]))


class LayerProofBatched:
    def __init__(self, proof, claims_prod_left, claims_prod_right):
        self.proof, self.claims_prod_left, self.claims_prod_right = proof, claims_prod_left, claims_prod_right


def _eq_point(rand, rand_prod):
    eq = 1
    for a, b in zip(rand, rand_prod):
        eq = eq * ((a * b + (1 - a) * (1 - b)) % R_MOD) % R_MOD
    return eq


class ProductCircuitEvalProofBatched:
    def __init__(self, proof, claims_dotp):
        self.proof = proof
        self.claims_dotp = claims_dotp

    @staticmethod
    def prove(ctx, prod_circuits, dotp_circuits, transcript):
        """product_tree.rs:251-392.  Consumes the circuits' layers (they are bound in place), as the reference does."""
        assert prod_circuits
        claims_dotp_final = ([], [], [])
        layers = []
        num_layers = prod_circuits[0].num_layers
        claims_to_verify = [c.evaluate() for c in prod_circuits]
        rand = []
        for layer_id in reversed(range(num_layers)):
            seq, seq_resident = [], []
            if layer_id == 0 and dotp_circuits:
                for d in dotp_circuits:
                    claims_to_verify.append(d.evaluate())
                    if isinstance(d.left, tuple):       # (Poly, offset) segments of resident polynomials
                        assert d.n == 1 << len(rand)
                        seq_resident.append((d.left, d.right, d.weight))
                    else:
                        assert len(d.left) == 1 << len(rand)
                        seq.append((d.left, d.right, d.weight))
            state = BatchedSumcheckState(ctx, [c.gpu for c in prod_circuits], layer_id,
                                         fr_vec_from_ints(rand) if rand else None, seq, seq_resident)
            num_rounds = len(rand)                                  # log2(len / 2), len / 2 = |eq(rand)|
            coeff_vec = transcript.challenge_scalars(b"rand_coeffs_next_layer", len(claims_to_verify))
            claim = sum(a * b for a, b in zip(claims_to_verify, coeff_vec)) % R_MOD
            proof, rand_prod, claims_prod, claims_dotp = SumcheckInstanceProof.prove_cubic_batched(
                claim, num_rounds, state, coeff_vec, transcript)
            state.close()
            left, right, _eq = claims_prod
            for i in range(len(prod_circuits)):
                transcript.append_scalar(b"claim_prod_left", left[i])
                transcript.append_scalar(b"claim_prod_right", right[i])
            if layer_id == 0 and dotp_circuits:
                dl, dr, dw = claims_dotp
                for i in range(len(dotp_circuits)):
                    transcript.append_scalar(b"claim_dotp_left", dl[i])
                    transcript.append_scalar(b"claim_dotp_right", dr[i])
                    transcript.append_scalar(b"claim_dotp_weight", dw[i])
                claims_dotp_final = (dl, dr, dw)
            r_layer = transcript.challenge_scalar(b"challenge_r_layer")
            claims_to_verify = [(l + r_layer * (r - l)) % R_MOD for l, r in zip(left, right)]
            rand = [r_layer] + rand_prod
            layers.append(LayerProofBatched(proof, left, right))
        return ProductCircuitEvalProofBatched(layers, claims_dotp_final), rand

    def verify(self, claims_prod_vec, claims_dotp_vec, length, transcript):
        """product_tree.rs:394-537; raises on rejection, returns (claims_prod, claims_dotp, rand)."""
        num_layers = length.bit_length() - 1
        assert len(self.proof) == num_layers
        rand = []
        claims_to_verify = list(claims_prod_vec)
        claims_to_verify_dotp = []
        nprod = len(claims_prod_vec)
        for num_rounds, i in enumerate(range(num_layers)):
            if i == num_layers - 1:
                claims_to_verify = claims_to_verify + list(claims_dotp_vec)
            coeff_vec = transcript.challenge_scalars(b"rand_coeffs_next_layer", len(claims_to_verify))
            claim = sum(a * b for a, b in zip(claims_to_verify, coeff_vec)) % R_MOD
            claim_last, rand_prod = self.proof[i].proof.verify(claim, num_rounds, 3, transcript)
            left, right = self.proof[i].claims_prod_left, self.proof[i].claims_prod_right
            assert len(left) == nprod and len(right) == nprod
            for j in range(nprod):
                transcript.append_scalar(b"claim_prod_left", left[j])
                transcript.append_scalar(b"claim_prod_right", right[j])
            assert len(rand) == len(rand_prod)
            eq = _eq_point(rand, rand_prod)
            expected = sum(coeff_vec[j] * (left[j] * right[j] % R_MOD * eq % R_MOD) for j in range(nprod)) % R_MOD
            if i == num_layers - 1:
                dl, dr, dw = self.claims_dotp
                for k in range(len(dl)):
                    transcript.append_scalar(b"claim_dotp_left", dl[k])
                    transcript.append_scalar(b"claim_dotp_right", dr[k])
                    transcript.append_scalar(b"claim_dotp_weight", dw[k])
                    expected = (expected + coeff_vec[k + nprod] * dl[k] % R_MOD * dr[k] % R_MOD * dw[k]) % R_MOD
            if expected != claim_last:
                raise ValueError("product layer %d rejected" % i)
            r_layer = transcript.challenge_scalar(b"challenge_r_layer")
            claims_to_verify = [(l + r_layer * (r - l)) % R_MOD for l, r in zip(left, right)]
            if i == num_layers - 1:
                dl, dr, dw = self.claims_dotp
                for k in range(len(claims_dotp_vec) // 2):
                    for v in (dl, dr, dw):
                        claims_to_verify_dotp.append((v[2 * k] + r_layer * (v[2 * k + 1] - v[2 * k])) % R_MOD)
            rand = [r_layer] + rand_prod
        return claims_to_verify, claims_to_verify_dotp, rand
