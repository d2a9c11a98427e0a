"""Timing of the opening path (BASELINE.json configs[3]): Hyrax opening pieces at n = R_size in {1024, 2048, 8192}
(bound over 2^ell scalars, the (n+1)-point commitment Cx, the bullet reduction with device-resident generators) and the
R1CS-sat sumcheck round evaluation + bind over four 2^20 tables, next to the CPU oracle on the host cores.
Writes gpurun_out/opening_bench.json."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import oracle as orc
from spartan_bn254_b200 import Context, synth
from spartan_bn254_b200.hyrax import DotProductProofGens

ctx = Context(0)
out = {"host_cores": os.cpu_count()}
R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001


def opening(ell, cpu=True):
    l, r_ = ell // 2, ell - ell // 2
    L, n = 1 << l, 1 << r_
    gens = DotProductProofGens(n, b"gens_r1cs_eval", ctx)
    bases = gens.device_bases_ext()
    q = synth.uniform_scalars(13, 1)[0]
    Z = synth.uniform_scalars(6, 1 << ell)
    Lv = synth.uniform_scalars(7, L)
    b_vec = synth.uniform_scalars(8, n)
    lg = n.bit_length() - 1
    us = synth.uniform_scalars(9, lg); bl = synth.uniform_scalars(10, lg); br = synth.uniform_scalars(11, lg)
    uinv = orc.to_mont([pow(v, -1, R_MOD) for v in orc.from_mont(us)])
    blind = synth.uniform_scalars(12, 1)[0]
    res = {"ell": ell, "L_size": L, "n": n}
    tu = time.perf_counter(); poly = ctx.poly_upload(Z); res["poly_upload_ms"] = 1e3 * (time.perf_counter() - tu)
    for rep in range(2):
        t0 = time.perf_counter(); LZ = poly.bound(Lv, L, n); t1 = time.perf_counter()
        Cx = ctx.commit(bases, LZ, blind); t2 = time.perf_counter()
        st = ctx.bullet_begin(bases, None, LZ, b_vec, blind, q_scalar=q); t3 = time.perf_counter()
        tr = tf = 0.0
        for i in range(lg):
            a = time.perf_counter(); st.round(bl[i], br[i]); b = time.perf_counter(); st.fold(us[i], uinv[i]); c = time.perf_counter()
            tr += b - a; tf += c - b
        st.end(); st.close()
        t4 = time.perf_counter()
    poly.close()
    res["gpu_ms"] = dict(bound_resident=1e3 * (t1 - t0), commit_Cx=1e3 * (t2 - t1), bullet_begin_Gamma=1e3 * (t3 - t2),
                         bullet_rounds_LR=1e3 * tr, bullet_folds=1e3 * tf, total=1e3 * (t4 - t0))
    if cpu:
        t0 = time.perf_counter(); LZo = orc.bound(Z, Lv, L, n); t1 = time.perf_counter()
        assert np.array_equal(LZo, LZ)
        Q, _ = orc.scalar_mul(gens.gens_1.G[0], 0, q)
        o = orc.bullet_prove(Q, gens.gens_n.G, gens.gens_n.h, LZ, b_vec, blind, bl, br, us); t2 = time.perf_counter()
        res["cpu_ms"] = dict(bound=1e3 * (t1 - t0), bullet_prove=1e3 * (t2 - t1), total=1e3 * (t2 - t0))
    print(json.dumps(res), flush=True)
    return res


out["openings"] = [opening(20), opening(22), opening(25, cpu=False)]

# sumcheck: 4 tables x 2^20, all 20 rounds
n = 1 << 20
T = [synth.uniform_scalars(40 + k, n) for k in range(4)]
rs = synth.uniform_scalars(50, 20)
t0 = time.perf_counter(); st = ctx.sumcheck_begin(*T); t1 = time.perf_counter()
te = tb = 0.0
first = None
for j in range(20):
    a = time.perf_counter(); e = st.round_eval(); b = time.perf_counter(); st.bind(rs[j]); c = time.perf_counter()
    te += b - a; tb += c - b
    if j == 0:
        first = (1e3 * (b - a), 1e3 * (c - b), e)
st.end(); st.close()
t2 = time.perf_counter()
c0 = time.perf_counter(); eo = orc.sumcheck_cubic_eval(*T); c1 = time.perf_counter(); [orc.bind_top(t, rs[0]) for t in T]; c2 = time.perf_counter()
assert all(np.array_equal(first[2][k], eo[k]) for k in range(3))
out["sumcheck_2^20"] = dict(gpu_upload_ms=1e3 * (t1 - t0), gpu_20_round_evals_ms=1e3 * te, gpu_20_binds_ms=1e3 * tb,
                            gpu_round0_eval_ms=first[0], gpu_round0_bind_ms=first[1],
                            cpu_round0_eval_ms_1thread=1e3 * (c1 - c0), cpu_round0_bind_ms_1thread=1e3 * (c2 - c1))
print(json.dumps(out["sumcheck_2^20"]), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "opening_bench.json"), "w"), indent=1)
