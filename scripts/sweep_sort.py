"""Stage timings of one device-resident commit vs a context tunable: python sweep_sort.py L R key v1,v2,..."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from spartan_bn254_b200 import Context, synth

L = int(sys.argv[1]); R = int(sys.argv[2]); key = sys.argv[3]; vals = [int(x) for x in sys.argv[4].split(",")]
ctx = Context(0)
dev = torch.device("cuda", 0)
G, h = synth.distinct_generators(ctx, R)
bases = ctx.bases(G, h)
z = torch.from_numpy(synth.uniform_scalars(1, L * R).view(np.int64)).to(dev)
dC = torch.empty((L, 8), dtype=torch.int64, device=dev); dinf = torch.empty((L,), dtype=torch.uint8, device=dev)
ctx.set("chunk_rows", L)
ref = None
for v in vals:
    ctx.set(key, v)
    for _ in range(3):
        ctx.hyrax_commit_device(bases, z.data_ptr(), L, R, 0, dC.data_ptr(), dinf.data_ptr(), stream=0)
    p = ctx.last_commit_profile()
    out = dC.cpu().numpy().copy()
    if ref is None: ref = out
    print(f"L={L} R={R} {key}={v}: " + " ".join(f"{k}={v_['ms']:.3f}" for k, v_ in p.items()) + f" same={bool(np.array_equal(ref, out))}", flush=True)
