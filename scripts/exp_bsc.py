"""Product-layer argument with the transcript on the device (bsc_device = 1) against the host round loop (0): same proof,
wall time per prove at a few circuit sizes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from spartan_bn254_b200 import Context, synth
from spartan_bn254_b200.product_tree import ProductCircuit, ProductCircuitEvalProofBatched
from spartan_bn254_b200.transcript import Transcript

ctx = Context(0)
SIZES = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]] or [(10, 12), (14, 12), (18, 12), (22, 12), (21, 4)]
for k, P in SIZES:
    n = 1 << k
    polys = [synth.uniform_scalars(50 + i, n) for i in range(P)]
    res = {}
    for mode in (0, 1, 2, 0, 1, 2):
        ctx.set("bsc_device", mode)
        circuits = [ProductCircuit(ctx, p) for p in polys]
        ctx.synchronize()
        ctx.counters(reset=True)
        t0 = time.perf_counter()
        proof, rand = ProductCircuitEvalProofBatched.prove(ctx, circuits, [], Transcript(b"bench"))
        dt = 1e3 * (time.perf_counter() - t0)
        res[mode] = (dt, ctx.counters()["kernel_launches"], [bytes(np.asarray(r)) for r in rand])
        for c in circuits:
            c.close() if hasattr(c, "close") else None
    print(f"2^{k} x {P}: host loop {res[0][0]:.2f} ms ({res[0][1]} launches)  device tail {res[1][0]:.2f} ms ({res[1][1]} launches)  all rounds on device {res[2][0]:.2f} ms ({res[2][1]} launches)", flush=True)
