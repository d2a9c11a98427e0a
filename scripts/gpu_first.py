"""Exploratory GPU run: integer-pipe microbenchmarks, small parity checks against the oracle, and a
first timing of the cfg1 commit (1024 x 1024).  Writes gpurun_out/first.json."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import oracle as orc  # noqa: E402  (checker only)
from spartan_bn254_b200 import Context, synth  # noqa: E402
from spartan_bn254_b200.hyrax import MultiCommitGens  # noqa: E402

out = {}
ctx = Context(0)

names = {0: "imad_lo", 1: "imad_hi", 2: "imad_wide_x2", 3: "fq_mul"}
for k in range(4):
    v = ctx.microbench(k)
    out["microbench_" + names[k]] = v
    print(names[k], "%.4g /s" % v, flush=True)


def check(label, L, R, gens_kind, blinds, seed=1, zero_row=None, derefs=False):
    if gens_kind == "ref":
        gens = MultiCommitGens.new(R, b"gens_r1cs_eval", ctx)
        G, h = gens.G, gens.h
        Go, ho = orc.multi_commit_gens(b"gens_r1cs_eval", R)
        assert np.array_equal(G, Go) and np.array_equal(h, ho), "generator derivation mismatch"
    else:
        G, h = synth.distinct_generators(ctx, R)
    Z = synth.derefs_scalars((L * R).bit_length() - 1) if derefs else synth.uniform_scalars(seed, L * R)
    if zero_row is not None:
        Z.reshape(L, R, 4)[zero_row] = 0
    bl = synth.uniform_scalars(4, L) if blinds else None
    if bl is not None and zero_row is not None:
        bl[zero_row] = 0
    bases = ctx.bases(G, h)
    t0 = time.time()
    C, inf = ctx.hyrax_commit(bases, Z, L, R, bl)
    t1 = time.time()
    Co, info = orc.hyrax_commit(G, h, Z, L, R, bl)
    ok = bool(np.array_equal(C, Co) and np.array_equal(inf, info))
    nbad = int((C != Co).any(axis=1).sum())
    print(f"{label}: L={L} R={R} c={bases.window_bits} ok={ok} bad_rows={nbad} inf_rows={int(inf.sum())} "
          f"gpu_call={1e3*(t1-t0):.1f} ms", flush=True)
    out[label] = dict(ok=ok, bad_rows=nbad, c=bases.window_bits)
    return ok


check("tiny_ref", 4, 8, "ref", True, zero_row=3)
check("small_ref", 16, 16, "ref", False, zero_row=5)
check("cfg0_witness_ref", 64, 64, "ref", True)
check("cfg0_derefs_ref", 128, 256, "ref", False, derefs=True)
check("mid_distinct", 64, 512, "distinct", True)

# cfg1 timing: 1024 x 1024, distinct generators
L = R = 1024
G, h = synth.distinct_generators(ctx, R)
Z = synth.uniform_scalars(1, L * R)
bases = ctx.bases(G, h)
for it in range(4):
    t0 = time.time()
    C, inf = ctx.hyrax_commit(bases, Z, L, R, None)
    dt = time.time() - t0
    prof = ctx.last_commit_profile()
    print(f"cfg1 it{it}: wall {1e3*dt:.2f} ms  {L*(R+1)/dt/1e6:.1f} Mpts/s  profile " +
          " ".join(f"{k}={v['ms']:.2f}" for k, v in prof.items()), flush=True)
out["cfg1_wall_ms"] = 1e3 * dt
out["cfg1_profile"] = prof
# verify 64 sampled rows against the oracle
rows = np.arange(0, L, 16)
Co, info = orc.hyrax_commit(G, h, Z.reshape(L, R, 4)[rows].reshape(-1, 4), len(rows), R, None)
ok = bool(np.array_equal(C[rows], Co) and np.array_equal(inf[rows], info))
print("cfg1 sampled parity:", ok, flush=True)
out["cfg1_parity"] = ok

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "first.json"), "w"), indent=1)
