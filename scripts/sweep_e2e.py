"""End-to-end (pinned host buffers -> sbn_hyrax_commit) wall time vs chunk_rows."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from spartan_bn254_b200 import Context, synth
L = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
R = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
chunks = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [L, L // 2, L // 4]
firsts = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [0]
ctx = Context(0)
G, h = synth.distinct_generators(ctx, R)
bases = ctx.bases(G, h)
zs = [torch.from_numpy(synth.uniform_scalars(1 + i, L * R).view(np.int64)).pin_memory() for i in range(4)]
hC = torch.empty((L, 8), dtype=torch.int64).pin_memory(); hinf = torch.empty((L,), dtype=torch.uint8).pin_memory()
for chunk, first in [(c, f) for c in chunks for f in firsts]:
    ctx.set("chunk_rows", chunk); ctx.set("first_chunk_rows", first)
    for i in range(3):
        ctx.hyrax_commit_raw(bases, zs[i % 4].data_ptr(), L, R, 0, hC.data_ptr(), hinf.data_ptr())
    n = 20
    t0 = time.perf_counter()
    for i in range(n):
        ctx.hyrax_commit_raw(bases, zs[i % 4].data_ptr(), L, R, 0, hC.data_ptr(), hinf.data_ptr())
    dt = (time.perf_counter() - t0) / n
    print(f"chunk={chunk} first={first}: e2e {dt*1e3:.3f} ms  {L*R/dt/1e6:.1f} Mpts/s", flush=True)
