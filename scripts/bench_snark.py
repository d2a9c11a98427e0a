"""SNARK::prove end to end (BASELINE.json configs[0] and configs[4]): a synthetic satisfiable R1CS of keyless shape
(2^k constraints, 2^k variables, one public input, three non-zeros per row in the densest matrix so that nnz pads to
2^(k+2), as the Aptos-keyless circuit does) is encoded, proved through the GPU path with real Merlin transcripts, and the
proof is checked by the oracle's independent CPU restatement of SNARK::verify.  The reference has no .r1cs/.wtns offline.
Under torchrun (one process per GPU) every rank holds the instance and runs the prover; the derefs commitment -- the only
table-sized step that shards without an exchange per round -- is split by rows across the ranks and its row blocks are
all-gathered over NCCL; the reported time is the maximum over ranks.
Usage: bench_snark.py [log2_constraints=20] [--verify]   (--verify: also run the oracle's CPU verifier on the proof)"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from spartan_bn254_b200 import Context, synth
from spartan_bn254_b200.hyrax import fr_from_int, fr_vec_to_ints
from spartan_bn254_b200.lib import SpMat, ProdCircuit
from spartan_bn254_b200.r1csproof import R1CSShape, SNARK, SNARKGens
from spartan_bn254_b200.transcript import Transcript



def run(k=20, verify=False, quiet=False, ctx_in=None, keep=None, keep_instance=False):
    """Builds the instance, encodes and proves (3 warm proofs, best; max over ranks); returns the result dict on rank 0 and None
    on the other ranks.  Under torchrun every rank must call it.  `keep` (dict) receives the proof and what a verifier needs
    (tests/test_snark.py checks the keyless-scale proof with the oracle's verifier that way); verify=True does the same check
    here -- the oracle is imported only then, as the checker."""
    own_group = False
    n = 1 << k
    num_cons = num_vars = n
    num_inputs = 1
    rank, world, local_rank = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    shard = None
    if world > 1:
        import torch
        import torch.distributed as dist
        from spartan_bn254_b200.parallel import make_all_gather
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        own_group = not dist.is_initialized()
        if own_group:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        shard = (rank, world, make_all_gather(torch.device("cuda", local_rank)))
    ctx = ctx_in or Context(local_rank)
    out = {"log2_constraints": k, "num_vars": num_vars, "num_inputs": num_inputs, "n_gpus": world, "ms": {}}


    def timed(name, fn):
        ctx.synchronize(); t0 = time.perf_counter(); r = fn(); ctx.synchronize()
        out["ms"][name] = round(1e3 * (time.perf_counter() - t0), 3)
        if rank == 0 and not quiet:
            print(name, out["ms"][name], "ms", flush=True)
        return r


    # ---- synthetic satisfiable instance: row i is (3 terms) * (2 terms) = c_i * 1, c_i = (A z)_i (B z)_i
    rng = np.random.default_rng(7)
    vars_m = synth.uniform_scalars(1, num_vars)
    input_m = synth.uniform_scalars(2, num_inputs)
    z = np.zeros((2 * num_vars, 4), dtype=np.uint64)
    z[:num_vars] = vars_m; z[num_vars] = fr_from_int(1); z[num_vars + 1: num_vars + 1 + num_inputs] = input_m
    used = num_vars + 1 + num_inputs
    rows = np.arange(n, dtype=np.uint32)
    A = (np.repeat(rows, 3), rng.integers(0, used, size=3 * n, dtype=np.uint32), synth.uniform_scalars(3, 3 * n))
    B = (np.repeat(rows, 2), rng.integers(0, used, size=2 * n, dtype=np.uint32), synth.uniform_scalars(4, 2 * n))
    ta = SpMat(ctx, n, 2 * num_vars, *A); tb = SpMat(ctx, n, 2 * num_vars, *B)
    Az = SpMat.mulvec([ta], z); Bz = SpMat.mulvec([tb], z)
    ta.close(); tb.close()
    pc = ProdCircuit(ctx, np.concatenate([Az, Bz]))           # layer 1 = Az (.) Bz element-wise
    C = (rows, np.full(n, num_vars, dtype=np.uint32), pc.layer(1))
    pc.close()
    inst = timed("setup.instance_upload(A, B, C by row and by column)", lambda: R1CSShape(ctx, num_cons, num_vars, num_inputs, A, B, C))
    out["nnz"] = [len(m[0]) for m in inst.mats]
    gens = timed("setup.generators(one-off)", lambda: SNARKGens(ctx, num_cons, num_vars, num_inputs, inst.max_nnz()))
    comm, decomm = SNARK.encode(inst, gens)
    decomm.close()
    comm, decomm = timed("encode(dense representation + comb_ops/comb_mem commitments)", lambda: SNARK.encode(inst, gens))
    def barrier():
        if world > 1:
            dist.barrier()


    timings = {}
    barrier()
    t0 = time.perf_counter()
    proof = SNARK.prove(inst, comm, decomm, vars_m, input_m, gens, Transcript(b"snark"), 1, timings=timings, shard=shard)
    ctx.synchronize()
    out["ms"]["prove.first_call(cold kernels and workspaces)"] = round(1e3 * (time.perf_counter() - t0), 3)
    best = None
    for rep in range(3):
        timings = {}
        barrier()
        t0 = time.perf_counter()
        proof = SNARK.prove(inst, comm, decomm, vars_m, input_m, gens, Transcript(b"snark"), 1, timings=timings, shard=shard)
        ctx.synchronize()
        dt = 1e3 * (time.perf_counter() - t0)
        if world > 1:
            tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        if best is None or dt < best[0]:
            best = (dt, timings)
    out["ms"]["prove.SNARK_total"] = round(best[0], 3)
    out["prove_total_note"] = "best of 3 warm proofs; with N > 1 the maximum over ranks of each proof"
    timings = best[1]
    if rank != 0:
        decomm.close()
        for m_ in inst.by_row + inst.by_col:
            m_.close()
        barrier()
        if own_group:
            dist.destroy_process_group()
        return None
    out["prove_phases_ms"] = {a: (round(b, 3) if not isinstance(b, dict) else {x: round(y, 3) for x, y in b.items()}) for a, b in timings.items()}
    if not quiet:
        print(json.dumps(out["prove_phases_ms"], indent=1), flush=True)
    if keep is not None:
        keep.update(proof=proof, comm=comm, gens=gens, inputs=fr_vec_to_ints(input_m))
    if verify:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle as orc
        import snark_model as snm
        orc.build()
        g = lambda x: (x.gens.gens_n.G, x.gens.gens_n.h, x.gens.gens_1.G[0])
        sat = gens.gens_r1cs_sat
        gens_sat = dict(gens_1=snm.Gens(sat.gens_sc.gens_1.G, sat.gens_sc.gens_1.h), gens_3=snm.Gens(sat.gens_sc.gens_3.G, sat.gens_sc.gens_3.h),
                        gens_4=snm.Gens(sat.gens_sc.gens_4.G, sat.gens_sc.gens_4.h), pc=g(sat.gens_pc),
                        pc_1=snm.Gens(sat.gens_pc.gens.gens_1.G, sat.gens_pc.gens.gens_1.h))
        ev = gens.gens_r1cs_eval
        c = comm.comm
        cd = dict(num_cons=comm.num_cons, num_vars=comm.num_vars, num_inputs=comm.num_inputs, batch_size=c.batch_size, num_ops=c.num_ops,
                  num_mem_cells=c.num_mem_cells, comb_ops=(c.comm_comb_ops.C, c.comm_comb_ops.inf), comb_mem=(c.comm_comb_mem.C, c.comm_comb_mem.inf))
        t0 = time.perf_counter()
        ok = snm.snark_verify(proof, cd, fr_vec_to_ints(input_m), gens_sat, dict(ops=g(ev.gens_ops), mem=g(ev.gens_mem), derefs=g(ev.gens_derefs)),
                              orc.Transcript(b"snark"))
        out["verified_by_cpu_oracle"] = bool(ok)
        out["ms"]["verify.cpu_oracle(python + C, 1 thread)"] = round(1e3 * (time.perf_counter() - t0), 3)
    out["reference_published_M2Max_1thread_s"] = {"r1cs_sat_proof": 3.45, "instance_evaluations": 0.36, "eq_evals": 0.10, "derefs_computation": 0.14,
                                                 "derefs_commitment": 166.2, "network_construction": 4.07, "network_proof": 34.5,
                                                 "total_prove": 208.8, "encode": 60.7, "verify": 0.39,
                                                 "source": "BENCHMARK_RESULTS.md:22-42 (Aptos keyless: 1 040 083 constraints, nnz padded to 2^22)"}
    if not quiet:
        print(json.dumps(out, indent=1))
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(out, open(os.path.join(ROOT, "gpurun_out", "snark_%d%s.json" % (k, "" if world == 1 else "_n%d" % world)), "w"), indent=1)
    if keep_instance and keep is not None:      # scripts/profile_snark.py proves again on the warm instance
        keep["instance"] = (inst, comm, decomm, vars_m, input_m, gens, ctx)
        return out
    decomm.close()
    for m_ in inst.by_row + inst.by_col:
        m_.close()
    if world > 1:
        barrier()
        if own_group:
            dist.destroy_process_group()
    return out


if __name__ == "__main__":
    pos = [a for a in sys.argv[1:] if not a.startswith("-")]
    run(int(pos[0]) if pos else 20, "--verify" in sys.argv)
