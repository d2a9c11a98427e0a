"""Minimal driver for profilers: builds distinct generators and runs N commits of L x R (cfg1 default)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spartan_bn254_b200 import Context, synth  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
R = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
N = int(sys.argv[3]) if len(sys.argv) > 3 else 2
ctx = Context(0)
G, h = synth.distinct_generators(ctx, R)
bases = ctx.bases(G, h)
Z = synth.uniform_scalars(1, L * R)
for _ in range(N):
    C, inf = ctx.hyrax_commit(bases, Z, L, R, None)
print("profile", ctx.last_commit_profile())
