"""Host mirror of the reference's R1CS-satisfiability proof and of SNARK::prove over the GPU entry points
(second half).

  reference nizk/mod.rs:23-82, 86-150, 154-284, 288-401   Knowledge / Equality / Product / DotProduct proofs (provers)
  reference sumcheck.rs:465-649, 657-811                  ZKSumcheckInstanceProof::{prove_cubic_with_additive_term, prove_quad}
  reference r1csproof.rs:150-182                          R1CSSumcheckGens, R1CSGens
  reference r1csproof.rs:210-459                          R1CSProof::{commit_poly, prove}
  reference r1cs.rs:34-82, 126-171                        R1CSShape::{new, evaluate, multiply_vec, compute_eval_table_sparse}
  reference snark.rs:290-322, 417-484                     SNARKGens::new, SNARK::{encode, prove}

Table-sized work runs on the GPU: the witness commitment and its opening (Hyrax path), the sparse matrix-vector
products (resident compressed-row copies of A, B, C and of their transposes), both sumchecks' round evaluations and
binds, the eq tables, the witness evaluation.  The host keeps the Merlin transcript, the random tape and the
Sigma-protocols, whose 2- to 5-point commitments go through sbn_msm.
"""
import numpy as np

from .hyrax import (R_MOD, DensePolynomial, GroupElement, MultiCommitGens, PolyCommitmentGens, PolyEvalProof, fr_from_int,
                    fr_to_int, fr_vec_from_ints, fr_vec_to_ints, log_2)
from .lib import SpMat
from .product_tree import UniPoly
from .spark import (MultiSparseMatPolynomialAsDense, SparseMatPolyCommitmentGens, SparseMatPolyEvalProof,
                    append_poly_commitment, commit_dense, equalize)


def commit_scalars(gens, scalars, blind):
    """Commitments::commit (commitments.rs:118-154) for the short vectors of the Sigma-protocols: sum s_i G_i + blind h.
    The generator sets of the sumchecks are fixed, so the commitment runs over their resident digit-multiple tables (one
    launch, ~0.1 ms) rather than as a variable-base MSM (a 254-step double-and-add chain, ~2.4 ms)."""
    n = len(scalars)
    assert gens.n == n, "assert_eq!(gens_n.n, self.len())"
    out, inf = gens.ctx.commit(gens.device_bases(), fr_vec_from_ints(list(scalars)), fr_from_int(blind))
    return GroupElement(out, inf)


def commit_rows(gens, rows, blinds):
    """Several commitments over the same short generator set in one call (one CTA per row on the tabulated path)."""
    n = len(rows[0])
    assert gens.n == n and all(len(r) == n for r in rows), "assert_eq!(gens_n.n, self.len())"
    Z = fr_vec_from_ints([v for r in rows for v in r])
    C, inf = gens.ctx.hyrax_commit(gens.device_bases(), Z, len(rows), n, fr_vec_from_ints(list(blinds)))
    return [GroupElement(C[i], inf[i]) for i in range(len(rows))]


def _append(transcript, label, g):
    transcript.append_point(label, g.compress())


class KnowledgeProof:
    def __init__(self, alpha, z1, z2):
        self.alpha, self.z1, self.z2 = alpha, z1, z2

    @staticmethod
    def prove(gens_n, transcript, tape, x, r):
        transcript.append_protocol_name(b"knowledge proof")
        t1, t2 = tape.random_scalar(b"t1"), tape.random_scalar(b"t2")
        C = commit_scalars(gens_n, [x], r)
        _append(transcript, b"C", C)
        alpha = commit_scalars(gens_n, [t1], t2)
        _append(transcript, b"alpha", alpha)
        c = transcript.challenge_scalar(b"c")
        return KnowledgeProof(alpha, (x * c + t1) % R_MOD, (r * c + t2) % R_MOD), C


class EqualityProof:
    def __init__(self, alpha, z):
        self.alpha, self.z = alpha, z

    @staticmethod
    def prove(gens_n, transcript, tape, v1, s1, v2, s2):
        transcript.append_protocol_name(b"equality proof")
        r = tape.random_scalar(b"r")
        C1 = commit_scalars(gens_n, [v1], s1)
        _append(transcript, b"C1", C1)
        C2 = commit_scalars(gens_n, [v2], s2)
        _append(transcript, b"C2", C2)
        # alpha = r * h (mod.rs:117): the commitment to 0 with blind r over the resident tables -- the same group element as
        # the reference's scalar multiplication, without a 254-step double-and-add chain
        alpha = commit_scalars(gens_n, [0], r)
        _append(transcript, b"alpha", alpha)
        c = transcript.challenge_scalar(b"c")
        return EqualityProof(alpha, (c * (s1 - s2) + r) % R_MOD), C1, C2


class ProductProof:
    def __init__(self, alpha, beta, delta, z):
        self.alpha, self.beta, self.delta, self.z = alpha, beta, delta, z

    @staticmethod
    def prove(gens_n, transcript, tape, x, rX, y, rY, z, rZ):
        transcript.append_protocol_name(b"product proof")
        b1, b2, b3, b4, b5 = (tape.random_scalar(l) for l in (b"b1", b"b2", b"b3", b"b4", b"b5"))
        X = commit_scalars(gens_n, [x], rX)
        _append(transcript, b"X", X)
        Y = commit_scalars(gens_n, [y], rY)
        _append(transcript, b"Y", Y)
        Z = commit_scalars(gens_n, [z], rZ)
        _append(transcript, b"Z", Z)
        alpha = commit_scalars(gens_n, [b1], b2)
        _append(transcript, b"alpha", alpha)
        beta = commit_scalars(gens_n, [b3], b4)
        _append(transcript, b"beta", beta)
        # delta = b3 * X + b5 * h over the one-off generator X (mod.rs:203-206).  X = x * G + rX * h was committed above, so
        # delta = (b3 x) * G + (b3 rX + b5) * h: the same group element, as a commitment over the resident tables
        delta = commit_scalars(gens_n, [b3 * x % R_MOD], (b3 * rX + b5) % R_MOD)
        _append(transcript, b"delta", delta)
        c = transcript.challenge_scalar(b"c")
        zs = [(b1 + c * x) % R_MOD, (b2 + c * rX) % R_MOD, (b3 + c * y) % R_MOD, (b4 + c * rY) % R_MOD,
              (b5 + c * (rZ - rX * y)) % R_MOD]
        return ProductProof(alpha, beta, delta, zs), X, Y, Z


class DotProductProof:
    def __init__(self, delta, beta, z, z_delta, z_beta):
        self.delta, self.beta, self.z, self.z_delta, self.z_beta = delta, beta, z, z_delta, z_beta

    @staticmethod
    def draw(tape, n):
        """The prover's randomness in the order the reference draws it (nizk/mod.rs:252-258)."""
        d_vec = tape.random_vector(b"d_vec", n)
        return d_vec, tape.random_scalar(b"r_delta"), tape.random_scalar(b"r_beta")

    @staticmethod
    def prove(gens_1, gens_n, transcript, tape, x_vec, blind_x, a_vec, y, blind_y, Cx=None, pre=None):
        """nizk/mod.rs:238-296.  Cx: the commitment to (x_vec, blind_x) when the caller already holds it (the sumcheck round
        has just sent it as comm_poly); the reference recomputes the same point.  pre = (d_vec, r_delta, r_beta, delta) when
        the caller drew the randomness ahead (in tape order) and committed delta with the other rounds' in one call.  Cy and
        beta have no challenge between them and are committed as two rows of one call."""
        transcript.append_protocol_name(b"dot product proof")
        n = len(x_vec)
        assert len(a_vec) == n and gens_n.n == n and gens_1.n == 1
        if pre is None:
            d_vec, r_delta, r_beta = DotProductProof.draw(tape, n)
            delta = commit_scalars(gens_n, d_vec, r_delta)
        else:
            d_vec, r_delta, r_beta, delta = pre
        if Cx is None:
            Cx = commit_scalars(gens_n, x_vec, blind_x)
        _append(transcript, b"Cx", Cx)
        dot = sum(a * d for a, d in zip(a_vec, d_vec)) % R_MOD
        Cy, beta = commit_rows(gens_1, [[y], [dot]], [blind_y, r_beta])
        _append(transcript, b"Cy", Cy)
        transcript.append_scalars(b"a", a_vec)
        _append(transcript, b"delta", delta)
        _append(transcript, b"beta", beta)
        c = transcript.challenge_scalar(b"c")
        z = [(c * x + d) % R_MOD for x, d in zip(x_vec, d_vec)]
        return DotProductProof(delta, beta, z, (c * blind_x + r_delta) % R_MOD, (c * blind_y + r_beta) % R_MOD), Cx, Cy


class ZKSumcheckInstanceProof:
    def __init__(self, comm_polys, comm_evals, proofs):
        self.comm_polys, self.comm_evals, self.proofs = comm_polys, comm_evals, proofs

    @staticmethod
    def _prove(state, nevals, claim, blind_claim, num_rounds, gens_1, gens_n, transcript, tape):
        """The round loop shared by prove_cubic_with_additive_term (sumcheck.rs:465-649, nevals = 3) and prove_quad
        (:657-811, nevals = 2); the tables are behind `state` (SumcheckState on the GPU)."""
        blinds_poly = tape.random_vector(b"blinds_poly", num_rounds)
        blinds_evals = tape.random_vector(b"blinds_evals", num_rounds)
        claim_per_round = claim
        comm_claim_per_round = commit_scalars(gens_1, [claim_per_round], blind_claim)
        # the dot-product proofs' randomness of every round, drawn in the reference's tape order (nothing else draws from the
        # tape inside the loop), so that the rounds' delta commitments are one call of num_rounds rows
        drawn = [DotProductProof.draw(tape, nevals + 1) for _ in range(num_rounds)]
        deltas = commit_rows(gens_n, [d[0] for d in drawn], [d[1] for d in drawn]) if num_rounds else []
        r, comm_polys, comm_evals, proofs = [], [], [], []
        for j in range(num_rounds):
            ev = [fr_to_int(e) for e in state.round_eval()]                   # e0, e2(, e3)
            evals = [ev[0], (claim_per_round - ev[0]) % R_MOD] + ev[1:]
            poly = UniPoly.from_evals(evals)
            comm_poly = commit_scalars(gens_n, poly.coeffs, blinds_poly[j])
            _append(transcript, b"comm_poly", comm_poly)
            comm_polys.append(comm_poly)
            r_j = transcript.challenge_scalar(b"challenge_nextround")
            state.bind(fr_from_int(r_j))
            ev_r = poly.evaluate(r_j)
            comm_eval = commit_scalars(gens_1, [ev_r], blinds_evals[j])
            _append(transcript, b"comm_claim_per_round", comm_claim_per_round)
            _append(transcript, b"comm_eval", comm_eval)
            w = transcript.challenge_scalars(b"combine_two_claims_to_one", 2)
            target = (w[0] * claim_per_round + w[1] * ev_r) % R_MOD
            blind_sc = blind_claim if j == 0 else blinds_evals[j - 1]
            blind = (w[0] * blind_sc + w[1] * blinds_evals[j]) % R_MOD
            deg = poly.degree()
            a_sc = [2] + [1] * deg
            a_eval = [1]
            for _ in range(deg):
                a_eval.append(a_eval[-1] * r_j % R_MOD)
            a = [(w[0] * s + w[1] * e) % R_MOD for s, e in zip(a_sc, a_eval)]
            proof, _, _ = DotProductProof.prove(gens_1, gens_n, transcript, tape, poly.coeffs, blinds_poly[j], a, target, blind,
                                                Cx=comm_poly, pre=drawn[j] + (deltas[j],))
            proofs.append(proof)
            claim_per_round, comm_claim_per_round = ev_r, comm_eval
            r.append(r_j)
            comm_evals.append(comm_eval)
        finals = [fr_to_int(x) for x in state.end()]
        return ZKSumcheckInstanceProof(comm_polys, comm_evals, proofs), r, finals, blinds_evals[num_rounds - 1]


class R1CSSumcheckGens:
    def __init__(self, label, gens_1_ref, ctx):
        self.gens_1 = gens_1_ref
        self.gens_3 = MultiCommitGens.new(3, label, ctx)
        self.gens_4 = MultiCommitGens.new(4, label, ctx)


class R1CSGens:
    def __init__(self, label, num_cons, num_vars, ctx):
        self.gens_pc = PolyCommitmentGens(log_2(num_vars), label, ctx)
        self.gens_sc = R1CSSumcheckGens(label, self.gens_pc.gens.gens_1, ctx)


class R1CSShape:
    """r1cs.rs:22-171 with the matrices resident on the device.  A, B, C: (rows, cols, Montgomery values) over
    num_cons x (2 * num_vars) (columns: vars, 1, inputs, zero padding -- z in Spartan's layout); sizes already padded."""

    def __init__(self, ctx, num_cons, num_vars, num_inputs, A, B, C):
        assert num_cons & (num_cons - 1) == 0 and num_vars & (num_vars - 1) == 0 and num_inputs < num_vars
        self.ctx, self.num_cons, self.num_vars, self.num_inputs = ctx, num_cons, num_vars, num_inputs
        self.mats = [tuple(np.asarray(x) for x in m) for m in (A, B, C)]
        ncols = 2 * num_vars
        self.by_row = [SpMat(ctx, num_cons, ncols, r, c, v) for r, c, v in self.mats]
        self.by_col = [SpMat(ctx, ncols, num_cons, c, r, v) for r, c, v in self.mats]

    def multiply_vec(self, z):
        return [SpMat.mulvec([m], z) for m in self.by_row]

    def compute_eval_table_combined(self, evals_rx, coeffs):
        """r_A evals_A + r_B evals_B + r_C evals_C of compute_eval_table_sparse (r1csproof.rs:378-389) in one pass."""
        return SpMat.mulvec(self.by_col, evals_rx, fr_vec_from_ints(coeffs))

    def max_nnz(self):
        return max(len(m[0]) for m in self.mats)


class R1CSProof:
    @staticmethod
    def prove(inst, vars_m, input_m, gens, transcript, tape, timings=None, shard=None):
        """r1csproof.rs:241-459.  vars_m / input_m: Montgomery uint64[n, 4]; returns (proof, rx, ry) with rx, ry canonical."""
        import time
        ctx = inst.ctx
        t = [time.perf_counter()]

        def lap(name):
            ctx.synchronize()
            t.append(time.perf_counter())
            if timings is not None:
                timings[name] = timings.get(name, 0.0) + 1e3 * (t[-1] - t[-2])

        transcript.append_protocol_name(b"R1CS proof")
        assert input_m.shape[0] < vars_m.shape[0]
        transcript.append_scalars(b"input", fr_vec_to_ints(input_m))
        # commit_poly (:210-237): blinds from the tape, one batched GPU commit
        poly_vars = DensePolynomial(vars_m)
        ell = poly_vars.get_num_vars()
        L_size = 1 << (ell // 2)
        blinds_m = tape.random_vector_mont(b"poly_blinds", L_size)
        poly_vars.resident(ctx)
        comm_vars = poly_vars.commit_inner(blinds_m, gens.gens_pc.gens.gens_n, shard=shard)
        append_poly_commitment(transcript, b"poly_commitment", comm_vars)
        lap("witness_commit_ms")
        num_vars = vars_m.shape[0]
        # z = [vars, 1, inputs, 0...] (r1csproof.rs:255-265) is assembled on the device from the witness polynomial that
        # commit_poly left resident; only the tail (1 and the public inputs) crosses the bus
        z_len = 2 * num_vars
        tail = np.concatenate([fr_from_int(1).reshape(1, 4), input_m.reshape(-1, 4)])
        num_rounds_x, num_rounds_y = log_2(inst.num_cons), log_2(2 * num_vars)
        tau = transcript.challenge_scalars(b"challenge_tau", num_rounds_x)
        # eq(tau), A z, B z, C z (r1csproof.rs:268-290) are built in HBM from the resident matrices
        st1 = ctx.sumcheck_begin_r1cs_resident(inst.by_row, poly_vars.resident(ctx), tail, z_len, fr_vec_from_ints(tau))
        lap("sumcheck1_setup(eq(tau), Az, Bz, Cz)_ms")
        sc1, rx, claims1, blind_claim_postsc1 = ZKSumcheckInstanceProof._prove(
            st1, 3, 0, 0, num_rounds_x, gens.gens_sc.gens_1, gens.gens_sc.gens_4, transcript, tape)
        st1.close()
        lap("sumcheck1_rounds_ms")
        tau_claim, Az_claim, Bz_claim, Cz_claim = claims1
        Az_blind, Bz_blind, Cz_blind, prod_blind = (tape.random_scalar(l) for l in
                                                    (b"Az_blind", b"Bz_blind", b"Cz_blind", b"prod_Az_Bz_blind"))
        g1 = gens.gens_sc.gens_1
        pok_Cz, comm_Cz = KnowledgeProof.prove(g1, transcript, tape, Cz_claim, Cz_blind)
        prod = Az_claim * Bz_claim % R_MOD
        proof_prod, comm_Az, comm_Bz, comm_prod = ProductProof.prove(g1, transcript, tape, Az_claim, Az_blind, Bz_claim, Bz_blind,
                                                                     prod, prod_blind)
        _append(transcript, b"comm_Az_claim", comm_Az)
        _append(transcript, b"comm_Bz_claim", comm_Bz)
        _append(transcript, b"comm_Cz_claim", comm_Cz)
        _append(transcript, b"comm_prod_Az_Bz_claims", comm_prod)
        blind_expected1 = tau_claim * (prod_blind - Cz_blind) % R_MOD
        claim_post1 = (Az_claim * Bz_claim - Cz_claim) * tau_claim % R_MOD
        proof_eq1, _, _ = EqualityProof.prove(g1, transcript, tape, claim_post1, blind_expected1, claim_post1, blind_claim_postsc1)
        r_A, r_B, r_C = (transcript.challenge_scalar(l) for l in (b"challenge_Az", b"challenge_Bz", b"challenge_Cz"))
        claim_phase2 = (r_A * Az_claim + r_B * Bz_claim + r_C * Cz_claim) % R_MOD
        blind_claim_phase2 = (r_A * Az_blind + r_B * Bz_blind + r_C * Cz_blind) % R_MOD
        lap("sigma_protocols_phase1_ms")
        # eq(rx) and r_A A^T eq(rx) + r_B B^T eq(rx) + r_C C^T eq(rx) (r1csproof.rs:378-410) likewise
        st2 = ctx.sumcheck_begin_quad_r1cs(inst.by_col, fr_vec_from_ints([r_A, r_B, r_C]), fr_vec_from_ints(rx), None,
                                           z_len=z_len)            # z is still resident from phase 1
        lap("sumcheck2_setup(eq(rx), eval tables)_ms")
        sc2, ry, claims2, blind_claim_postsc2 = ZKSumcheckInstanceProof._prove(
            st2, 2, claim_phase2, blind_claim_phase2, num_rounds_y, gens.gens_sc.gens_1, gens.gens_sc.gens_3, transcript, tape)
        st2.close()
        lap("sumcheck2_rounds_ms")
        eval_vars_at_ry = fr_to_int(poly_vars._poly.evaluate(fr_vec_from_ints(ry[1:])))
        blind_eval = tape.random_scalar(b"blind_eval")
        proof_eval, comm_vars_at_ry = PolyEvalProof.prove(poly_vars, blinds_m, ry[1:], eval_vars_at_ry, blind_eval, gens.gens_pc,
                                                          transcript, tape)
        lap("witness_opening_ms")
        blind_eval_Z = (1 - ry[0]) * blind_eval % R_MOD
        blind_expected2 = claims2[1] * blind_eval_Z % R_MOD
        claim_post2 = claims2[0] * claims2[1] % R_MOD
        proof_eq2, _, _ = EqualityProof.prove(gens.gens_pc.gens.gens_1, transcript, tape, claim_post2, blind_expected2, claim_post2,
                                              blind_claim_postsc2)
        lap("sigma_protocols_phase2_ms")
        proof = R1CSProof()
        proof.comm_vars, proof.sc_proof_phase1 = comm_vars, sc1
        proof.claims_phase2 = (comm_Az, comm_Bz, comm_Cz, comm_prod)
        proof.pok_claims_phase2 = (pok_Cz, proof_prod)
        proof.proof_eq_sc_phase1, proof.sc_proof_phase2 = proof_eq1, sc2
        proof.comm_vars_at_ry, proof.proof_eval_vars_at_ry, proof.proof_eq_sc_phase2 = comm_vars_at_ry, proof_eval, proof_eq2
        return proof, rx, ry


class SNARKGens:
    """snark.rs:296-322 (sizes already padded by the caller)."""

    def __init__(self, ctx, num_cons, num_vars, num_inputs, num_nz_entries):
        self.gens_r1cs_sat = R1CSGens(b"gens_r1cs_sat", num_cons, num_vars, ctx)
        # R1CSCommitmentGens::new (r1cs.rs:275-288)
        self.gens_r1cs_eval = SparseMatPolyCommitmentGens(b"gens_r1cs_eval", log_2(num_cons), log_2(2 * num_vars),
                                                          num_nz_entries, 3, ctx)
        # the resident copies the prover commits and opens over (window tables, digit-multiple tables), built here once
        # rather than inside the first proof
        ev = self.gens_r1cs_eval
        for pc in (self.gens_r1cs_sat.gens_pc, ev.gens_ops, ev.gens_mem, ev.gens_derefs):
            pc.gens.gens_n.device_bases()
            pc.gens.device_bases_ext()
            pc.gens.gens_1.device_bases()
        sc = self.gens_r1cs_sat.gens_sc
        for g in (sc.gens_1, sc.gens_3, sc.gens_4):
            g.device_bases()


class R1CSCommitment:
    def __init__(self, num_cons, num_vars, num_inputs, comm):
        self.num_cons, self.num_vars, self.num_inputs, self.comm = num_cons, num_vars, num_inputs, comm

    def append_to_transcript(self, transcript):
        """r1cs.rs:355-363 + sparse_mlpoly_full.rs:700-708 (append_u64 = 8 little-endian bytes, merlin)."""
        for label, v in ((b"num_cons", self.num_cons), (b"num_vars", self.num_vars), (b"num_inputs", self.num_inputs),
                         (b"batch_size", self.comm.batch_size), (b"num_ops", self.comm.num_ops),
                         (b"num_mem_cells", self.comm.num_mem_cells)):
            transcript.append_message(label, int(v).to_bytes(8, "little"))
        append_poly_commitment(transcript, b"comm_comb_ops", self.comm.comm_comb_ops)
        append_poly_commitment(transcript, b"comm_comb_mem", self.comm.comm_comb_mem)


class SNARK:
    def __init__(self, r1cs_sat_proof, inst_evals, r1cs_eval_proof):
        self.r1cs_sat_proof, self.inst_evals, self.r1cs_eval_proof = r1cs_sat_proof, inst_evals, r1cs_eval_proof

    @staticmethod
    def encode(inst, gens):
        """SNARK::encode (snark.rs:417-426) -> R1CSShape::commit (r1cs.rs:375-400)."""
        row = np.zeros((3, 0), dtype=np.uint32)
        N = 1
        while N < inst.max_nnz():
            N <<= 1
        row = np.zeros((3, N), dtype=np.uint32)
        col = np.zeros((3, N), dtype=np.uint32)
        val = np.zeros((3 * N, 4), dtype=np.uint64)
        for i, (r, c, v) in enumerate(inst.mats):
            row[i, :len(r)] = r
            col[i, :len(c)] = c
            val[i * N: i * N + len(r)] = v
        dense = MultiSparseMatPolynomialAsDense.from_arrays(inst.ctx, max(inst.num_cons, 2 * inst.num_vars), row, col, val)
        comm = commit_dense(dense, gens.gens_r1cs_eval)
        return R1CSCommitment(inst.num_cons, inst.num_vars, inst.num_inputs, comm), dense

    @staticmethod
    def prove(inst, comm, decomm, vars_m, input_m, gens, transcript, tape_seed, timings=None, shard=None):
        """snark.rs:428-484.  The reference seeds its random tape from the OS; the seed is injected here."""
        import time
        from .transcript import RandomTape
        tape = RandomTape(b"snark_proof", tape_seed)
        transcript.append_protocol_name(b"Spartan SNARK proof")
        comm.append_to_transcript(transcript)
        sat_t, eval_t = {}, {}
        t0 = time.perf_counter()
        sat_proof, rx, ry = R1CSProof.prove(inst, vars_m, input_m, gens.gens_r1cs_sat, transcript, tape, timings=sat_t, shard=shard)
        t1 = time.perf_counter()
        rx_e, ry_e = equalize(rx, ry)
        inst_evals = decomm.multi_evaluate(rx_e, ry_e)                     # inst.evaluate(rx, ry), r1cs.rs:126-129
        inst.ctx.synchronize()
        t2 = time.perf_counter()
        eval_proof = SparseMatPolyEvalProof.prove(decomm, rx, ry, inst_evals, gens.gens_r1cs_eval, transcript, tape, timings=eval_t,
                                                  shard=shard)
        inst.ctx.synchronize()
        t3 = time.perf_counter()
        if timings is not None:
            timings["r1cs_sat_proof_ms"] = 1e3 * (t1 - t0)
            timings["instance_evaluations_ms"] = 1e3 * (t2 - t1)
            timings["r1cs_eval_proof_ms"] = 1e3 * (t3 - t2)
            timings["sat"] = sat_t
            timings["eval"] = eval_t
        return SNARK(sat_proof, tuple(inst_evals), eval_proof)
