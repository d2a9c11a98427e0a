"""Latency of a short commitment (commitments.rs:118-154 over gens_1 / gens_3 / gens_4: the Sigma-protocol and ZK-sumcheck
commitments, a few hundred per proof) through sbn_commit: the tabulated one-launch path (small_kernels.cuh) against one row
of the general commit pipeline.  Usage: bench_small_commit.py [calls=500]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spartan_bn254_b200 import Context, synth
from spartan_bn254_b200.hyrax import MultiCommitGens

calls = int(sys.argv[1]) if len(sys.argv) > 1 else 500
out = {}
for path in (1, 0):
    ctx = Context(0)
    ctx.set("small_commit_path", path)
    for n in (1, 4):
        gens = MultiCommitGens.new(n, b"gens_r1cs_sat", ctx)
        b = gens.device_bases()
        Z = synth.uniform_scalars(1, n)
        bl = synth.uniform_scalars(2, 1)
        for _ in range(20):
            ctx.commit(b, Z, bl[0])
        t0 = time.perf_counter()
        for _ in range(calls):
            ctx.commit(b, Z, bl[0])
        out["%s_n%d_us" % ("tabulated" if path else "pipeline", n)] = round(1e6 * (time.perf_counter() - t0) / calls, 1)
    # single rows over an opening's generator set (Cx of nizk/mod.rs:470; every bullet round commits two such rows):
    # the tabulated sum-of-table-points path against the bucket pipeline
    from spartan_bn254_b200.hyrax import DotProductProofGens
    import time as _t
    for n in (1024, 8192):
        d = DotProductProofGens(n, b"gens_r1cs_eval", ctx)
        t0 = _t.perf_counter()
        b = d.device_bases_ext()
        ctx.synchronize()
        if path:
            out["table_build_n%d_ms" % n] = round(1e3 * (_t.perf_counter() - t0), 1)
        Z = synth.uniform_scalars(3, n)
        bl = synth.uniform_scalars(4, 1)
        for _ in range(10):
            ctx.commit(b, Z, bl[0])
        t0 = time.perf_counter()
        for _ in range(calls // 5):
            ctx.commit(b, Z, bl[0])
        out["%s_n%d_us" % ("tabulated" if path else "pipeline", n)] = round(1e6 * (time.perf_counter() - t0) / (calls // 5), 1)
        b.close()
    ctx.close()
print(json.dumps(out))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "small_commit.json"), "w"))
