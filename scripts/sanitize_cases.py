"""Small instances of every kernel family for compute-sanitizer (memcheck / racecheck / initcheck / synccheck):
  1. tabulated-sum commit 300 x 33 (layouts 1 and 2, table budget 64 MiB so the table build stays small)
  2. small-scalar schedule on the same shape
  3. bucket-pipeline commit 64 x 64 and a 300 x 33 one with the table disabled (batched-affine rounds of the bucket path)
  4. one Hyrax opening (PolyEvalProof::prove: bound, Cx, bullet rounds, delta) and one MSM
  5. one SNARK::prove of a 16-constraint R1CS: both ZK sumchecks, product layers, hash layer, derefs, openings
Every result is compared with the CPU oracle, so a sanitizer run is also a parity run.
usage: compute-sanitizer --tool memcheck python scripts/sanitize_cases.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import oracle as orc  # noqa: E402
from spartan_bn254_b200 import Context, synth  # noqa: E402

orc.build()
ctx = Context(0)
ctx.set("mult_max_mb", 64)
L, R = 300, 33
G, h = synth.distinct_generators(ctx, R)
Z = synth.uniform_scalars(1, L * R)
Z[7 * R:8 * R] = 0
blinds = synth.uniform_scalars(2, L)
C_ref, inf_ref = orc.hyrax_commit(G, h, Z, L, R, blinds)
C0_ref, inf0_ref = orc.hyrax_commit(G, h, Z, L, R, None)
for layout in (1, 2, 0):
    ctx.set("mult_layout", layout)
    bases = ctx.bases(G, h)
    C, inf = ctx.hyrax_commit(bases, Z, L, R, blinds)
    assert np.array_equal(C, C_ref) and np.array_equal(inf, inf_ref), layout
    assert bases.mult_table()[0] > 0
    C, inf = ctx.hyrax_commit(bases, Z, L, R, None)
    assert np.array_equal(C, C0_ref) and np.array_equal(inf, inf0_ref), layout
    bases.close()
    print("tabulated-sum commit 300 x 33, layout", layout, "ok", flush=True)
ctx.set("mult_layout", 1)
Zs = ctx.fr_from_canonical(synth.small_scalars_canonical(3, L * R))
bases = ctx.bases(G, h)
C, inf = ctx.hyrax_commit(bases, Zs, L, R, None)
Cs_ref, infs_ref = orc.hyrax_commit(G, h, Zs, L, R, None)
assert np.array_equal(C, Cs_ref) and np.array_equal(inf, infs_ref)
print("small-scalar schedule ok, commits that took it:", ctx.memory_stats()["small_scalar_commits"], flush=True)
bases.close()
ctx.set("mult_max_mb", 0)
ctx.set("ba_rounds", 1)
bases = ctx.bases(G, h)
C, inf = ctx.hyrax_commit(bases, Z, L, R, blinds)
assert np.array_equal(C, C_ref) and np.array_equal(inf, inf_ref)
bases.close()
ctx.set("ba_rounds", -1)
G64, h64 = synth.distinct_generators(ctx, 64)
Z64 = synth.uniform_scalars(4, 64 * 64)
bases = ctx.bases(G64, h64)
C, inf = ctx.hyrax_commit(bases, Z64, 64, 64, blinds[:64])
Cr, ir = orc.hyrax_commit(G64, h64, Z64, 64, 64, blinds[:64])
assert np.array_equal(C, Cr) and np.array_equal(inf, ir)
bases.close()
print("bucket pipeline 300 x 33 (one batched-affine round) and 64 x 64 ok", flush=True)
ctx.set("mult_max_mb", 64)
out, oinf = ctx.msm(G64, np.zeros(64, dtype=np.uint8), Z64[:64])
exp, einf = orc.msm(G64, np.zeros(64, dtype=np.uint8), Z64[:64], 1)
assert oinf == einf and np.array_equal(out, exp)
print("msm ok", flush=True)
import test_snark  # noqa: E402
assert test_snark._prove_and_verify(ctx, orc, 16, 16, 2, 116)
print("SNARK::prove (16 constraints) accepted by the oracle's verifier", flush=True)
ctx.close()
print("all sanitizer cases ok")
