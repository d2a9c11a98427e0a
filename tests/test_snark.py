"""SNARK::prove end to end (snark.rs:428-484): R1CS-sat proof + instance evaluations + Spark evaluation proof through the GPU
path, on a synthetic satisfiable R1CS, accepted by the oracle's independent restatement of SNARK::verify (the reference's own
tests are prove -> verify round trips of this kind, snark.rs:535-616, r1csproof.rs:651-681)."""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001


def synthetic_r1cs(seed, num_cons, num_vars, num_inputs):
    """Row i: (sum of 3 terms) * (sum of 2 terms) = c_i * 1; columns in Spartan's z layout [vars, 1, inputs, padding]."""
    rnd = random.Random(seed)
    vars_ = [rnd.randrange(R) for _ in range(num_vars)]
    inputs = [rnd.randrange(R) for _ in range(num_inputs)]
    z = vars_ + [1] + inputs + [0] * (num_vars - 1 - num_inputs)
    used = list(range(num_vars + 1 + num_inputs))
    A, B, C = [], [], []
    for i in range(num_cons):
        ta = [(rnd.choice(used), rnd.randrange(1, R)) for _ in range(3)]
        tb = [(rnd.choice(used), rnd.randrange(1, R)) for _ in range(2)]
        az = sum(z[c] * v for c, v in ta) % R
        bz = sum(z[c] * v for c, v in tb) % R
        A += [(i, c, v) for c, v in ta]
        B += [(i, c, v) for c, v in tb]
        C.append((i, num_vars, az * bz % R))
    return vars_, inputs, A, B, C


def _arrays(entries):
    from spartan_bn254_b200.hyrax import fr_vec_from_ints
    return (np.array([e[0] for e in entries], dtype=np.uint32), np.array([e[1] for e in entries], dtype=np.uint32),
            fr_vec_from_ints([e[2] for e in entries]))


def _prove_and_verify(ctx, orc, num_cons, num_vars, num_inputs, seed, tamper=None):
    import snark_model as snm
    from spartan_bn254_b200.hyrax import fr_vec_from_ints
    from spartan_bn254_b200.r1csproof import R1CSShape, SNARK, SNARKGens
    from spartan_bn254_b200.transcript import Transcript
    vars_, inputs, A, B, C = synthetic_r1cs(seed, num_cons, num_vars, num_inputs)
    inst = R1CSShape(ctx, num_cons, num_vars, num_inputs, _arrays(A), _arrays(B), _arrays(C))
    gens = SNARKGens(ctx, num_cons, num_vars, num_inputs, inst.max_nnz())
    comm, decomm = SNARK.encode(inst, gens)
    proof = SNARK.prove(inst, comm, decomm, fr_vec_from_ints(vars_), fr_vec_from_ints(inputs), gens, Transcript(b"snark"), 4242)
    decomm.close()
    g = lambda x: (x.gens.gens_n.G, x.gens.gens_n.h, x.gens.gens_1.G[0])
    sat = gens.gens_r1cs_sat
    gens_sat = dict(gens_1=snm.Gens(sat.gens_sc.gens_1.G, sat.gens_sc.gens_1.h), gens_3=snm.Gens(sat.gens_sc.gens_3.G, sat.gens_sc.gens_3.h),
                    gens_4=snm.Gens(sat.gens_sc.gens_4.G, sat.gens_sc.gens_4.h), pc=g(sat.gens_pc),
                    pc_1=snm.Gens(sat.gens_pc.gens.gens_1.G, sat.gens_pc.gens.gens_1.h))
    ev = gens.gens_r1cs_eval
    gens_eval = dict(ops=g(ev.gens_ops), mem=g(ev.gens_mem), derefs=g(ev.gens_derefs))
    c = comm.comm
    cd = dict(num_cons=comm.num_cons, num_vars=comm.num_vars, num_inputs=comm.num_inputs, batch_size=c.batch_size, num_ops=c.num_ops,
              num_mem_cells=c.num_mem_cells, comb_ops=(c.comm_comb_ops.C, c.comm_comb_ops.inf),
              comb_mem=(c.comm_comb_mem.C, c.comm_comb_mem.inf))
    if tamper:
        tamper(proof, inputs)
    return snm.snark_verify(proof, cd, inputs, gens_sat, gens_eval, orc.Transcript(b"snark"))


@pytest.mark.parametrize("num_cons,num_vars,num_inputs", [(4, 4, 1), (16, 16, 2), (64, 32, 3), (32, 64, 0)])
def test_snark_prove_is_accepted(ctx, orc, num_cons, num_vars, num_inputs):
    assert _prove_and_verify(ctx, orc, num_cons, num_vars, num_inputs, 100 + num_cons)


def test_snark_rejections(ctx, orc):
    import snark_model as snm

    def bad_input(proof, inputs):
        inputs[0] = (inputs[0] + 1) % R

    def bad_eval(proof, inputs):
        proof.inst_evals = ((proof.inst_evals[0] + 1) % R,) + tuple(proof.inst_evals[1:])

    def bad_sigma(proof, inputs):
        proof.r1cs_sat_proof.proof_eq_sc_phase1.z = (proof.r1cs_sat_proof.proof_eq_sc_phase1.z + 1) % R

    for tamper in (bad_input, bad_eval, bad_sigma):
        with pytest.raises(snm.VerifyError):
            _prove_and_verify(ctx, orc, 16, 16, 2, 7, tamper)


def test_sparse_matvec_with_heavy_rows(ctx):
    """sbn_spmat_mulvec against host integers, including a row of 5000 entries (handled by a whole block) and the combined
    r_A A + r_B B form."""
    from spartan_bn254_b200.hyrax import fr_vec_from_ints, fr_vec_to_ints
    from spartan_bn254_b200.lib import SpMat
    rnd = random.Random(3)
    n, ncols = 64, 97
    mats = []
    for m in range(2):
        ent = [(rnd.randrange(n), rnd.randrange(ncols), rnd.randrange(R)) for _ in range(300)]
        ent += [(5 + m, rnd.randrange(ncols), rnd.randrange(R)) for _ in range(5000)]
        mats.append(ent)
    vec = [rnd.randrange(R) for _ in range(ncols)]
    coeffs = [rnd.randrange(R), rnd.randrange(R)]
    want = [0] * n
    single = [0] * n
    for m, ent in enumerate(mats):
        for r, c, v in ent:
            want[r] = (want[r] + coeffs[m] * v % R * vec[c]) % R
            if m == 0:
                single[r] = (single[r] + v * vec[c]) % R
    sp = [SpMat(ctx, n, ncols, [e[0] for e in ent], [e[1] for e in ent], fr_vec_from_ints([e[2] for e in ent])) for ent in mats]
    assert fr_vec_to_ints(SpMat.mulvec(sp, fr_vec_from_ints(vec), fr_vec_from_ints(coeffs))) == want
    assert fr_vec_to_ints(SpMat.mulvec(sp[:1], fr_vec_from_ints(vec))) == single
    for s_ in sp:
        s_.close()


@pytest.mark.parametrize("num_cons,num_vars", [(16, 8), (64, 64), (32, 256)])
def test_resident_sumcheck_setups_match_host_tables(ctx, num_cons, num_vars):
    """sbn_sumcheck_begin_r1cs / _begin_quad_r1cs (tables built in HBM from the resident matrices, r1csproof.rs:268-290 and
    :378-410) against host integers: every round's evaluations and the final table values equal those of the host-table
    set-ups fed with eq(tau), A z, B z, C z (resp. z and the combined evaluation table) computed in Python."""
    from spartan_bn254_b200.hyrax import fr_vec_from_ints, fr_vec_to_ints, fr_to_int, fr_from_int
    from spartan_bn254_b200.r1csproof import R1CSShape
    rnd = random.Random(11)
    vars_, inputs, A, B, Cm = synthetic_r1cs(5, num_cons, num_vars, 1)
    inst = R1CSShape(ctx, num_cons, num_vars, 1, _arrays(A), _arrays(B), _arrays(Cm))
    z = vars_ + [1] + inputs + [0] * (num_vars - 2)
    zm = fr_vec_from_ints(z)

    def eq(point):
        t = [1]
        for r in point:
            t = [v for x in t for v in (x * (1 - r) % R, x * r % R)]
        return t

    def drive(st, n):
        out = []
        for _ in range(n):
            out.append([fr_to_int(e) for e in st.round_eval()])
            st.bind(fr_from_int(rnd.randrange(R)))
        out.append(fr_vec_to_ints(st.end()))
        st.close()
        return out

    lx, ly = num_cons.bit_length() - 1, (2 * num_vars).bit_length() - 1
    tau = [rnd.randrange(R) for _ in range(lx)]
    mv = lambda ent: [sum(v * z[c] for r, c, v in ent if r == i) % R for i in range(num_cons)]
    host = ctx.sumcheck_begin(fr_vec_from_ints(eq(tau)), fr_vec_from_ints(mv(A)), fr_vec_from_ints(mv(B)), fr_vec_from_ints(mv(Cm)))
    state = rnd.getstate()
    want = drive(host, lx)
    rnd.setstate(state)
    assert drive(ctx.sumcheck_begin_r1cs(inst.by_row, zm, fr_vec_from_ints(tau)), lx) == want
    rx = [rnd.randrange(R) for _ in range(lx)]
    co = [rnd.randrange(R) for _ in range(3)]
    e = eq(rx)
    table = [0] * (2 * num_vars)
    for k, ent in enumerate((A, B, Cm)):
        for r, c, v in ent:
            table[c] = (table[c] + co[k] * v % R * e[r]) % R
    host = ctx.sumcheck_begin_quad(zm, fr_vec_from_ints(table))
    state = rnd.getstate()
    want = drive(host, ly)
    rnd.setstate(state)
    assert drive(ctx.sumcheck_begin_quad_r1cs(inst.by_col, fr_vec_from_ints(co), fr_vec_from_ints(rx), zm), ly) == want
    # z = None: the z uploaded by the phase-1 set-up above is still resident
    rnd.setstate(state)
    assert drive(ctx.sumcheck_begin_quad_r1cs(inst.by_col, fr_vec_from_ints(co), fr_vec_from_ints(rx), None, z_len=len(z)), ly) == want
    from spartan_bn254_b200 import SbnError
    with pytest.raises(SbnError):
        ctx.sumcheck_begin_quad_r1cs(inst.by_col, fr_vec_from_ints(co), fr_vec_from_ints(rx), None, z_len=len(z) * 2)
    for m_ in inst.by_row + inst.by_col:
        m_.close()


def test_keyless_scale_proof_is_accepted(ctx, orc):
    """BASELINE configs[4] at full size: the proof scripts/bench_snark.py times (2^20 constraints, nnz padded to 2^22) is
    accepted by the oracle's CPU restatement of SNARK::verify."""
    import os
    import sys
    import snark_model as snm
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
    import bench_snark
    keep = {}
    out = bench_snark.run(20, verify=False, quiet=True, ctx_in=ctx, keep=keep)
    assert out["ms"]["prove.SNARK_total"] > 0
    gens, comm, proof = keep["gens"], keep["comm"], keep["proof"]
    g = lambda x: (x.gens.gens_n.G, x.gens.gens_n.h, x.gens.gens_1.G[0])
    sat = gens.gens_r1cs_sat
    gens_sat = dict(gens_1=snm.Gens(sat.gens_sc.gens_1.G, sat.gens_sc.gens_1.h), gens_3=snm.Gens(sat.gens_sc.gens_3.G, sat.gens_sc.gens_3.h),
                    gens_4=snm.Gens(sat.gens_sc.gens_4.G, sat.gens_sc.gens_4.h), pc=g(sat.gens_pc),
                    pc_1=snm.Gens(sat.gens_pc.gens.gens_1.G, sat.gens_pc.gens.gens_1.h))
    ev = gens.gens_r1cs_eval
    c = comm.comm
    cd = dict(num_cons=comm.num_cons, num_vars=comm.num_vars, num_inputs=comm.num_inputs, batch_size=c.batch_size, num_ops=c.num_ops,
              num_mem_cells=c.num_mem_cells, comb_ops=(c.comm_comb_ops.C, c.comm_comb_ops.inf), comb_mem=(c.comm_comb_mem.C, c.comm_comb_mem.inf))
    assert snm.snark_verify(proof, cd, keep["inputs"], gens_sat, dict(ops=g(ev.gens_ops), mem=g(ev.gens_mem), derefs=g(ev.gens_derefs)),
                            orc.Transcript(b"snark"))
