"""Tunable sweep on one GPU: stage times with a single sequential chunk, and wall time pipelined."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from spartan_bn254_b200 import Context, synth

L = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
R = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
cs = [0] if len(sys.argv) < 4 else [int(x) for x in sys.argv[3].split(",")]
chunks = [L, max(1, L // 2), max(1, L // 4)] if len(sys.argv) < 5 else [int(x) for x in sys.argv[4].split(",")]
ms = [8, 16, 32] if len(sys.argv) < 6 else [int(x) for x in sys.argv[5].split(",")]
caps = [0] if len(sys.argv) < 7 else [int(x) for x in sys.argv[6].split(",")]
ctx = Context(0)
G, h = synth.distinct_generators(ctx, R)
Z = synth.uniform_scalars(1, L * R)
for c in cs:
    ctx.set("window_bits", c)
    bases = ctx.bases(G, h)
    for chunk in chunks:
        for m in ms:
            for cap in caps:
                ctx.set("chunk_rows", chunk); ctx.set("leaf_m", m); ctx.set("task_cap", cap)
                best = 1e9
                for it in range(4):
                    t0 = time.perf_counter()
                    ctx.hyrax_commit(bases, Z, L, R, None)
                    dt = time.perf_counter() - t0
                    if it: best = min(best, dt)
                p = ctx.last_commit_profile()
                print(f"c={bases.window_bits} chunk={chunk} m={m} cap={cap}: wall {best*1e3:.2f} ms | " +
                      " ".join(f"{k}={v['ms']:.2f}" for k, v in p.items()), flush=True)
