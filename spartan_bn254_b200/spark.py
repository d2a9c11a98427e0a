"""Host mirror of the reference's Spark evaluation argument (R1CSEvalProof = SparseMatPolyEvalProof, the part of
SNARK::prove that holds ~97 % of the reference's prove time) over the GPU entry points (first half).

  reference sparse_mlpoly_full.rs:60-196    SparseMatPolynomial, multi_sparse_to_dense_rep, multi_commit
  reference sparse_mlpoly_full.rs:362-482   DerefsEvalProof
  reference sparse_mlpoly_full.rs:604-631   SparseMatPolyCommitmentGens::new
  reference sparse_mlpoly_full.rs:880-1046  HashLayerProof::prove
  reference sparse_mlpoly_full.rs:1292-1428 ProductLayerProof::prove
  reference sparse_mlpoly_full.rs:1541-1578 PolyEvalNetworkProof::prove
  reference sparse_mlpoly_full.rs:1674-1755 SparseMatPolyEvalProof::{equalize, prove}

Every table-sized object lives in HBM: the address / timestamp vectors and comb_ops / comb_mem from encode time, the eq
tables, the derefs polynomial, the hash layers and product circuits at prove time.  The host keeps the Merlin transcript
and the few dozen field elements per round; proofs are plain Python objects holding canonical integers and GroupElements.
"""
import numpy as np

from .hyrax import (R_MOD, PolyCommitment, PolyCommitmentGens, PolyEvalProof, fr_from_int, fr_to_int,
                    fr_vec_from_ints, fr_vec_to_ints, log_2)
from .lib import Poly
from .product_tree import ProductCircuitEvalProofBatched
from .sparse_mlpoly import PolyEvalNetwork, SparkAddresses


def _next_pow2(n):
    return 1 if n <= 1 else 1 << (n - 1).bit_length()


class SparseMatPolynomial:
    """sparse_mlpoly_full.rs:60-108: entries are (row, col, canonical value)."""

    def __init__(self, num_vars_x, num_vars_y, entries):
        self.num_vars_x, self.num_vars_y, self.M = num_vars_x, num_vars_y, list(entries)

    def get_num_nz_entries(self):
        return _next_pow2(len(self.M))

    @staticmethod
    def multi_evaluate(polys, rx, ry):
        """:110-118 on the host (small instances / tests): sum val * eq(rx)[row] * eq(ry)[col]."""
        from .hyrax import EqPolynomial
        ex, ey = EqPolynomial(rx).evals(), EqPolynomial(ry).evals()
        return [sum(ex[r] * ey[c] % R_MOD * v for r, c, v in p.M) % R_MOD for p in polys]


class SparseMatPolyCommitmentGens:
    def __init__(self, label, num_vars_x, num_vars_y, num_nz_entries, batch_size, ctx):
        nz = log_2(_next_pow2(num_nz_entries))
        self.gens_ops = PolyCommitmentGens(nz + log_2(_next_pow2(batch_size * 5)), label, ctx)
        self.gens_mem = PolyCommitmentGens(max(num_vars_x, num_vars_y) + 1, label, ctx)
        self.gens_derefs = PolyCommitmentGens(nz + log_2(_next_pow2(batch_size * 2)), label, ctx)


class ResidentDense:
    """A DensePolynomial whose evaluations live in HBM (what PolyEvalProof::prove needs of one)."""

    def __init__(self, poly):
        self.poly = poly
        self.len = poly.len
        self.num_vars = log_2(poly.len)

    def get_num_vars(self):
        return self.num_vars

    def bound(self, L, ctx=None):
        left = self.num_vars // 2
        return self.poly.bound(L, 1 << left, 1 << (self.num_vars - left))

    def commit(self, gens):
        """DensePolynomial::commit(gens, None) (hyrax.rs:283-308): zero blinds."""
        left = self.num_vars // 2
        C, inf = self.poly.commit(gens.gens.gens_n.device_bases(), 1 << left, 1 << (self.num_vars - left), None)
        return PolyCommitment(C, inf)


class MultiSparseMatPolynomialAsDense:
    """sparse_mlpoly_full.rs:120-172 (multi_sparse_to_dense_rep), dense representation resident on the GPU."""

    def __init__(self, ctx, polys):
        assert polys and all(p.num_vars_x == polys[0].num_vars_x and p.num_vars_y == polys[0].num_vars_y for p in polys)
        self.ctx = ctx
        self.batch_size = len(polys)
        self.N = N = max(p.get_num_nz_entries() for p in polys)
        row = np.zeros((self.batch_size, N), dtype=np.uint32)
        col = np.zeros((self.batch_size, N), dtype=np.uint32)
        val = np.zeros((self.batch_size * N, 4), dtype=np.uint64)
        for i, p in enumerate(polys):
            if p.M:
                row[i, :len(p.M)] = [e[0] for e in p.M]
                col[i, :len(p.M)] = [e[1] for e in p.M]
                val[i * N: i * N + len(p.M)] = fr_vec_from_ints([e[2] for e in p.M])
        any_poly = polys[0]
        self.num_mem_cells = 1 << max(any_poly.num_vars_x, any_poly.num_vars_y)
        self.spark = SparkAddresses(ctx, self.num_mem_cells, row, col)
        self.comb_ops, self.comb_mem = self.spark.gpu.comb_polys(val)

    @classmethod
    def from_arrays(cls, ctx, num_mem_cells, row, col, val_mont):
        """The same from ready-made address arrays (uint32[batch, N]) and Montgomery values (benchmarks)."""
        self = cls.__new__(cls)
        self.ctx = ctx
        self.batch_size, self.N = row.shape
        self.num_mem_cells = num_mem_cells
        self.spark = SparkAddresses(ctx, num_mem_cells, row, col)
        self.comb_ops, self.comb_mem = self.spark.gpu.comb_polys(val_mont)
        return self

    def multi_evaluate(self, rx, ry):
        """SparseMatPolynomial::multi_evaluate (:110-118) on the device (rx, ry canonical ints, already equalised)."""
        from .hyrax import fr_vec_to_ints
        return fr_vec_to_ints(self.spark.gpu.evaluate(self.comb_ops, fr_vec_from_ints(rx), fr_vec_from_ints(ry)))

    def close(self):
        self.comb_ops.close()
        self.comb_mem.close()
        self.spark.close()


class SparseMatPolyCommitment:
    def __init__(self, batch_size, num_ops, num_mem_cells, comm_comb_ops, comm_comb_mem):
        self.batch_size, self.num_ops, self.num_mem_cells = batch_size, num_ops, num_mem_cells
        self.comm_comb_ops, self.comm_comb_mem = comm_comb_ops, comm_comb_mem


def multi_commit(ctx, polys, gens):
    """SparseMatPolynomial::multi_commit (:176-196): the encode-time commitments of comb_ops and comb_mem."""
    dense = MultiSparseMatPolynomialAsDense(ctx, polys)
    return commit_dense(dense, gens), dense


def commit_dense(dense, gens):
    c_ops = ResidentDense(dense.comb_ops).commit(gens.gens_ops)
    c_mem = ResidentDense(dense.comb_mem).commit(gens.gens_mem)
    return SparseMatPolyCommitment(dense.batch_size, dense.N, dense.num_mem_cells, c_ops, c_mem)


def append_poly_commitment(transcript, label, comm):
    """hyrax.rs:44-52."""
    transcript.append_message(label, b"poly_commitment_begin")
    transcript.append_points(b"poly_commitment_share", comm.C, comm.inf)
    transcript.append_message(label, b"poly_commitment_end")


def _bound_bot_all(vals, challenges):
    """bound_poly_var_bot for i in (0..len).rev() (hyrax.rs:205-214) on a short host vector."""
    vals = list(vals)
    for ch in reversed(challenges):
        vals = [(vals[2 * i] + ch * (vals[2 * i + 1] - vals[2 * i])) % R_MOD for i in range(len(vals) // 2)]
    assert len(vals) == 1
    return vals[0]


def _joint_opening(poly, evals, r, gens, labels, transcript, random_tape):
    """The n-to-1 reduction + Hyrax opening shared by DerefsEvalProof::prove_single (:375-410) and the comb_ops / comb_mem
    openings of HashLayerProof::prove (:986-1034).  labels = (claims, challenge, joint claim)."""
    evals = list(evals) + [0] * (_next_pow2(len(evals)) - len(evals))
    transcript.append_scalars(labels[0], evals)
    challenges = transcript.challenge_scalars(labels[1], log_2(len(evals)))
    joint = _bound_bot_all(evals, challenges)
    r_joint = challenges + list(r)
    transcript.append_scalar(labels[2], joint)
    proof, _ = PolyEvalProof.prove(ResidentDense(poly), None, r_joint, joint, None, gens, transcript, random_tape)
    return proof


class _SegmentCircuit:
    """DotProductCircuit over segments of resident polynomials (left, right, weight: (Poly, offset))."""

    def __init__(self, left, right, weight, n):
        self.left, self.right, self.weight, self.n = left, right, weight, n

    def evaluate(self):
        (a, oa), (b, ob), (c, oc) = self.left, self.right, self.weight
        return fr_to_int(Poly.triple_dot(a, oa, b, ob, c, oc, self.n))

    def split(self):
        idx = self.n // 2
        assert idx * 2 == self.n
        lo = _SegmentCircuit(self.left, self.right, self.weight, idx)
        hi = _SegmentCircuit(*[(p, o + idx) for p, o in (self.left, self.right, self.weight)], idx)
        return lo, hi


class ProductLayerProof:
    def __init__(self, eval_row, eval_col, eval_val, proof_mem, proof_ops):
        self.eval_row, self.eval_col, self.eval_val, self.proof_mem, self.proof_ops = eval_row, eval_col, eval_val, proof_mem, proof_ops

    @staticmethod
    def prove(ctx, row_pl, col_pl, dense, derefs_poly, evals, transcript):
        """sparse_mlpoly_full.rs:1305-1428."""
        transcript.append_protocol_name(b"Sparse polynomial product layer proof")
        sides = []
        for name, pl in ((b"row", row_pl), (b"col", col_pl)):
            e_init, e_audit = pl.init.evaluate(), pl.audit.evaluate()
            e_read = [c.evaluate() for c in pl.read_vec]
            e_write = [c.evaluate() for c in pl.write_vec]
            ws = rs = 1
            for w, r in zip(e_write, e_read):
                ws, rs = ws * w % R_MOD, rs * r % R_MOD
            assert e_init * ws % R_MOD == rs * e_audit % R_MOD, "memory-checking subset check"
            transcript.append_scalar(b"claim_" + name + b"_eval_init", e_init)
            transcript.append_scalars(b"claim_" + name + b"_eval_read", e_read)
            transcript.append_scalars(b"claim_" + name + b"_eval_write", e_write)
            transcript.append_scalar(b"claim_" + name + b"_eval_audit", e_audit)
            sides.append((e_init, e_read, e_write, e_audit))
        b, N = dense.batch_size, dense.N
        assert len(evals) == b
        dotp, left_vec, right_vec = [], [], []
        for i in range(b):
            circ = _SegmentCircuit((derefs_poly, i * N), (derefs_poly, (b + i) * N), (dense.comb_ops, (4 * b + i) * N), N)
            lo, hi = circ.split()
            el, er = lo.evaluate(), hi.evaluate()
            transcript.append_scalar(b"claim_eval_dotp_left", el)
            transcript.append_scalar(b"claim_eval_dotp_right", er)
            assert (el + er) % R_MOD == evals[i] % R_MOD, "dot-product circuit does not evaluate to the claimed value"
            left_vec.append(el)
            right_vec.append(er)
            dotp += [lo, hi]
        ops = row_pl.read_vec + row_pl.write_vec + col_pl.read_vec + col_pl.write_vec
        proof_ops, rand_ops = ProductCircuitEvalProofBatched.prove(ctx, ops, dotp, transcript)
        mem = [row_pl.init, row_pl.audit, col_pl.init, col_pl.audit]
        proof_mem, rand_mem = ProductCircuitEvalProofBatched.prove(ctx, mem, [], transcript)
        return ProductLayerProof(sides[0], sides[1], (left_vec, right_vec), proof_mem, proof_ops), rand_mem, rand_ops


class HashLayerProof:
    def __init__(self, eval_row, eval_col, eval_val, eval_derefs, proof_ops, proof_mem, proof_derefs):
        self.eval_row, self.eval_col, self.eval_val, self.eval_derefs = eval_row, eval_col, eval_val, eval_derefs
        self.proof_ops, self.proof_mem, self.proof_derefs = proof_ops, proof_mem, proof_derefs

    @staticmethod
    def prove(rand, dense, derefs_poly, gens, transcript, random_tape):
        """sparse_mlpoly_full.rs:922-1046."""
        transcript.append_protocol_name(b"Sparse polynomial hash layer proof")
        rand_mem, rand_ops = rand
        b, N, M = dense.batch_size, dense.N, dense.num_mem_cells
        r_ops_m, r_mem_m = fr_vec_from_ints(rand_ops), fr_vec_from_ints(rand_mem)

        def evs(poly, r_m, stride, count):
            """The evaluations at one point of `count` consecutive segments: one eq table and one launch for all of them."""
            return fr_vec_to_ints(poly.evaluate_strided(r_m, 0, stride, count))

        d = evs(derefs_poly, r_ops_m, N, 2 * b)
        eval_row_ops_val, eval_col_ops_val = d[:b], d[b:]
        # DerefsEvalProof::prove (:412-432)
        transcript.append_protocol_name(b"Derefs evaluation proof")
        proof_derefs = _joint_opening(derefs_poly, eval_row_ops_val + eval_col_ops_val, rand_ops, gens.gens_derefs,
                                      (b"evals_ops_val", b"challenge_combine_n_to_one", b"joint_claim_eval"), transcript,
                                      random_tape)
        ops = dense.comb_ops
        o = evs(ops, r_ops_m, N, 5 * b)
        row_addr, row_read, col_addr, col_read, eval_val = (o[k * b: (k + 1) * b] for k in range(5))
        row_audit, col_audit = evs(dense.comb_mem, r_mem_m, M, 2)
        proof_ops = _joint_opening(ops, row_addr + row_read + col_addr + col_read + eval_val, rand_ops, gens.gens_ops,
                                   (b"claim_evals_ops", b"challenge_combine_n_to_one", b"joint_claim_eval_ops"), transcript,
                                   random_tape)
        proof_mem = _joint_opening(dense.comb_mem, [row_audit, col_audit], rand_mem, gens.gens_mem,
                                   (b"claim_evals_mem", b"challenge_combine_two_to_one", b"joint_claim_eval_mem"), transcript,
                                   random_tape)
        return HashLayerProof((row_addr, row_read, row_audit), (col_addr, col_read, col_audit), eval_val,
                              (eval_row_ops_val, eval_col_ops_val), proof_ops, proof_mem, proof_derefs)


class PolyEvalNetworkProof:
    def __init__(self, proof_prod_layer, proof_hash_layer):
        self.proof_prod_layer, self.proof_hash_layer = proof_prod_layer, proof_hash_layer

    @staticmethod
    def prove(ctx, network, dense, derefs_poly, evals, gens, transcript, random_tape, timings=None):
        """sparse_mlpoly_full.rs:1546-1578."""
        import time
        transcript.append_protocol_name(b"Sparse polynomial evaluation proof")
        t0 = time.perf_counter()
        prod, rand_mem, rand_ops = ProductLayerProof.prove(ctx, network.row_layers.prod_layer, network.col_layers.prod_layer,
                                                           dense, derefs_poly, evals, transcript)
        ctx.synchronize()
        t1 = time.perf_counter()
        hashp = HashLayerProof.prove((rand_mem, rand_ops), dense, derefs_poly, gens, transcript, random_tape)
        ctx.synchronize()
        if timings is not None:
            timings["network_proof.product_layers_ms"] = 1e3 * (t1 - t0)
            timings["network_proof.hash_layer(evaluations + 3 openings)_ms"] = 1e3 * (time.perf_counter() - t1)
        return PolyEvalNetworkProof(prod, hashp)


def equalize(rx, ry):
    """sparse_mlpoly_full.rs:1674-1691."""
    rx, ry = list(rx), list(ry)
    if len(rx) < len(ry):
        rx = [0] * (len(ry) - len(rx)) + rx
    elif len(ry) < len(rx):
        ry = [0] * (len(rx) - len(ry)) + ry
    return rx, ry


class SparseMatPolyEvalProof:
    def __init__(self, comm_derefs, poly_eval_network_proof):
        self.comm_derefs, self.poly_eval_network_proof = comm_derefs, poly_eval_network_proof

    @staticmethod
    def prove(dense, rx, ry, evals, gens, transcript, random_tape, timings=None, shard=None):
        """sparse_mlpoly_full.rs:1694-1755 (Hyrax mode).  rx, ry, evals: canonical ints.  `timings` (dict) receives the wall
        time of the phases keyless_benchmark.rs times separately ([a]-[e], examples/keyless_benchmark.rs:190-235)."""
        import time
        ctx = dense.ctx
        t = [time.perf_counter()]

        def lap(name):
            ctx.synchronize()
            t.append(time.perf_counter())
            if timings is not None:
                timings[name] = timings.get(name, 0.0) + 1e3 * (t[-1] - t[-2])

        transcript.append_protocol_name(b"Sparse polynomial evaluation proof")
        assert len(evals) == dense.batch_size
        rx_ext, ry_ext = equalize(rx, ry)
        rx_m, ry_m = fr_vec_from_ints(rx_ext), fr_vec_from_ints(ry_ext)
        # [a] eq tables, [b] derefs, [c] derefs commitment: one call, everything in HBM
        C, inf, derefs_poly = dense.spark.derefs_commit(gens.gens_derefs.gens.gens_n, rx_m, ry_m, shard)
        comm_derefs = PolyCommitment(C, inf)
        lap("eq_tables+derefs+derefs_commitment_ms")
        transcript.append_message(b"derefs_commitment", b"begin_derefs_commitment")
        append_poly_commitment(transcript, b"comm_poly_row_col_ops_val", comm_derefs)
        transcript.append_message(b"derefs_commitment", b"end_derefs_commitment")
        r_mem_check = transcript.challenge_scalars(b"challenge_r_hash", 2)
        lap("transcript_commitment_ms")
        net = PolyEvalNetwork(dense.spark, rx_m, ry_m, (fr_from_int(r_mem_check[0]), fr_from_int(r_mem_check[1])))
        lap("network_construction_ms")
        proof = PolyEvalNetworkProof.prove(ctx, net, dense, derefs_poly, evals, gens, transcript, random_tape, timings)
        lap("network_proof_ms")
        for lay in (net.row_layers, net.col_layers):
            for c in lay.prod_layer.all():
                c.close()
        derefs_poly.close()
        lap("free_device_tables_ms")
        return SparseMatPolyEvalProof(comm_derefs, proof)
