"""Product-layer argument at keyless scale: P product circuits of 2^k entries proved on the GPU (tables resident, Merlin
on the host), next to the CPU restatement's single-thread cost of one first-round evaluation + bind of the same length."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
from spartan_bn254_b200 import Context, synth
from spartan_bn254_b200.product_tree import ProductCircuit, ProductCircuitEvalProofBatched
from spartan_bn254_b200.transcript import Transcript

k = int(sys.argv[1]) if len(sys.argv) > 1 else 22
P = int(sys.argv[2]) if len(sys.argv) > 2 else 12
n = 1 << k
ctx = Context(0)
polys = [synth.uniform_scalars(50 + i, n) for i in range(P)]
out = {"log2_len": k, "circuits": P}
t0 = time.perf_counter()
circuits = [ProductCircuit(ctx, p) for p in polys]
ctx.synchronize()
out["build_ms_incl_h2d"] = 1e3 * (time.perf_counter() - t0)
claims = [c.evaluate() for c in circuits]
ctx.counters(reset=True)
t0 = time.perf_counter()
proof, rand = ProductCircuitEvalProofBatched.prove(ctx, circuits, [], Transcript(b"bench"))
out["prove_ms"] = 1e3 * (time.perf_counter() - t0)
out["kernel_launches"] = ctx.counters()["kernel_launches"]
t0 = time.perf_counter()
proof.verify(claims, [], n, Transcript(b"bench"))
out["verify_ms_host_python"] = 1e3 * (time.perf_counter() - t0)
# algorithmic HBM bytes of the whole proof: per layer of table length T (T = n/2, n/4, ...), round j reads
# (2P + 1) * T / 2^j * 32 B in the evaluation and again in the bind, which writes half of it back
byt = 0
T = n // 2
while T >= 1:
    t = T
    while t >= 2:
        byt += (2 * P + 1) * t * 32 * 2 + (2 * P + 1) * (t // 2) * 32
        t //= 2
    T //= 2
out["algorithmic_hbm_bytes"] = byt
out["achieved_GBps_over_whole_prove"] = byt / (out["prove_ms"] * 1e-3) / 1e9
try:
    import oracle as orc
    orc.build()
    m = 1 << min(k, 20)
    A = synth.uniform_scalars(1, m); B = synth.uniform_scalars(2, m); Cc = synth.uniform_scalars(3, m); D = synth.uniform_scalars(4, m)
    t0 = time.perf_counter(); orc.sumcheck_cubic_eval(A, B, Cc, D); te = time.perf_counter() - t0
    t0 = time.perf_counter(); orc.bind_top(A, synth.uniform_scalars(9, 1)[0]); tb = time.perf_counter() - t0
    out["cpu_1thread_round0_eval_ms_4tables_len_2^%d" % min(k, 20)] = 1e3 * te
    out["cpu_1thread_bind_ms_1table_len_2^%d" % min(k, 20)] = 1e3 * tb
except Exception as ex:
    out["cpu"] = "unavailable: %r" % (ex,)
print(json.dumps(out, indent=1))
