"""Minimal driver for profilers: N device-resident commits of L x R over distinct generators with the digit-multiple table
(bench.py's configuration).  usage: commit_dev_once.py [L R N table_mb key=value ...]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch  # noqa: E402
from spartan_bn254_b200 import Context, synth  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
R = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
N = int(sys.argv[3]) if len(sys.argv) > 3 else 2
table_mb = int(sys.argv[4]) if len(sys.argv) > 4 else 36000
ctx = Context(0)
ctx.set("mult_max_mb", table_mb)
for kv in sys.argv[5:]:
    k, v = kv.split("=")
    ctx.set(k, int(v))
dev = torch.device("cuda", 0)
G, h = synth.distinct_generators(ctx, R)
bases = ctx.bases(G, h)
Z = torch.from_numpy(synth.uniform_scalars(1, L * R).view(np.int64)).to(dev)
dC = torch.empty((L, 8), dtype=torch.int64, device=dev)
dinf = torch.empty((L,), dtype=torch.uint8, device=dev)
for _ in range(N):
    ctx.hyrax_commit_device(bases, Z.data_ptr(), L, R, 0, dC.data_ptr(), dinf.data_ptr(), stream=0)
torch.cuda.synchronize()
print("table", bases.mult_table(), "profile", ctx.last_commit_profile())
