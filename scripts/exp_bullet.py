"""One opening's bullet reduction (n generators) for a profiler: begin, log2(n) rounds + folds, end."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import oracle as orc
from spartan_bn254_b200 import Context, synth
from spartan_bn254_b200.hyrax import DotProductProofGens
R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
ctx = Context(0)
gens = DotProductProofGens(n, b"gens_r1cs_eval", ctx)
bases = gens.device_bases_ext()
q = synth.uniform_scalars(13, 1)[0]
a_vec, b_vec = synth.uniform_scalars(6, n), synth.uniform_scalars(8, n)
lg = n.bit_length() - 1
us = synth.uniform_scalars(9, lg); bl = synth.uniform_scalars(10, lg); br = synth.uniform_scalars(11, lg)
uinv = orc.to_mont([pow(v, -1, R_MOD) for v in orc.from_mont(us)])
blind = synth.uniform_scalars(12, 1)[0]
for rep in range(6):
    ctx.set("host_normalize", rep & 1)
    t0 = time.perf_counter()
    st = ctx.bullet_begin(bases, None, a_vec, b_vec, blind, q_scalar=q)
    t1 = time.perf_counter()
    tr = tf = 0.0
    per = []
    for i in range(lg):
        a = time.perf_counter(); st.round(bl[i], br[i]); b = time.perf_counter(); st.fold(us[i], uinv[i]); c = time.perf_counter()
        tr += b - a; tf += c - b; per.append(round(1e6 * (b - a)))
    st.end(); st.close()
    print(f"host_normalize={rep & 1} n={n} begin {1e3*(t1-t0):.3f} ms rounds {1e3*tr:.3f} ms folds {1e3*tf:.3f} ms per-round us {per}", flush=True)
