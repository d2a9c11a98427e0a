"""cfg2-scale commit on one GPU: timing per stage, e2e, and sampled parity against the oracle."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import oracle as orc
from spartan_bn254_b200 import Context, synth
from spartan_bn254_b200.hyrax import MultiCommitGens

L = int(sys.argv[1]); R = int(sys.argv[2]); kind = sys.argv[3] if len(sys.argv) > 3 else "uniform"
gens_kind = sys.argv[4] if len(sys.argv) > 4 else "distinct"
ctx = Context(0)
t0 = time.time()
if gens_kind == "ref":
    g = MultiCommitGens.new(R, b"gens_r1cs_eval", ctx); G, h = g.G, g.h
else:
    G, h = synth.distinct_generators(ctx, R)
t1 = time.time()
bases = ctx.bases(G, h)
t2 = time.time()
Z = synth.uniform_scalars(1, L * R) if kind == "uniform" else synth.derefs_scalars((L * R).bit_length() - 1)
t3 = time.time()
print(f"gens {t1-t0:.2f}s tables {t2-t1:.2f}s (c={bases.window_bits}) scalars {t3-t2:.2f}s", flush=True)
for chunk in (512, 1024):
    ctx.set("chunk_rows", chunk)
    for it in range(2):
        t = time.perf_counter()
        C, inf = ctx.hyrax_commit(bases, Z, L, R, None)
        dt = time.perf_counter() - t
    p = ctx.last_commit_profile()
    print(f"L={L} R={R} {kind}/{gens_kind} chunk={chunk}: e2e(pageable numpy) {dt*1e3:.1f} ms = {L*R/dt/1e6:.1f} Mpts/s | " +
          " ".join(f"{k}={v['ms']:.1f}" for k, v in p.items()), flush=True)
rows = np.unique(np.concatenate([np.arange(0, L, max(1, L // 24)), [L - 1, 3 * L // 4, 3 * L // 4 - 1]]))
Co, info = orc.hyrax_commit(G, h, Z.reshape(L, R, 4)[rows].reshape(-1, 4), len(rows), R, None)
ok = bool(np.array_equal(C[rows], Co) and np.array_equal(inf[rows], info))
print(f"sampled parity ({len(rows)} rows): {ok}; identity rows {int(inf.sum())}", flush=True)
assert ok
