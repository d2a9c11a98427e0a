"""Stage timings of one device-resident commit vs leaf_m (buckets per leaf thread of the two-level reduction) -- single chunk."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from spartan_bn254_b200 import Context, synth

L = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
R = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
variants = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "0,4,8,16,32").split(",")]
ctx = Context(0)
dev = torch.device("cuda", 0)
G, h = synth.distinct_generators(ctx, R)
bases = ctx.bases(G, h)
z = torch.from_numpy(synth.uniform_scalars(1, L * R).view(np.int64)).to(dev)
dC = torch.empty((L, 8), dtype=torch.int64, device=dev); dinf = torch.empty((L,), dtype=torch.uint8, device=dev)
ctx.set("chunk_rows", L)
ref = None
for leaf in variants:
    ctx.set("leaf_m", leaf)
    for _ in range(3):
        ctx.hyrax_commit_device(bases, z.data_ptr(), L, R, 0, dC.data_ptr(), dinf.data_ptr(), stream=0)
    p = ctx.last_commit_profile()
    out = dC.cpu().numpy().copy()
    if ref is None: ref = out
    same = bool(np.array_equal(ref, out))
    print(f"L={L} R={R} leaf_m={leaf}: " + " ".join(f"{k}={v['ms']:.3f}" for k, v in p.items()) + f" same={same}", flush=True)
