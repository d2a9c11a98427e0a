"""Profiler driver for the latency-bound stages: commits of L x R, one per leaf_m value (0 = auto)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spartan_bn254_b200 import Context, synth  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
R = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
variants = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0]
ctx = Context(0)
G, h = synth.distinct_generators(ctx, R)
bases = ctx.bases(G, h)
Z = synth.uniform_scalars(1, L * R)
ctx.set("chunk_rows", L)
for m in variants:
    ctx.set("leaf_m", m)
    C, inf = ctx.hyrax_commit(bases, Z, L, R, None)
    print("leaf_m", m, ctx.last_commit_profile())
