"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper over oracle/_build/liboracle.so (bn254_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  Arrays use the same layouts as the product C ABI (include/sbn254.h):
  scalars: uint64[n,4]  Montgomery Fr, LE limbs
  points : uint64[n,8]  affine Montgomery (x[4], y[4]); identity = zeros + inf byte 1
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")


def build(force=False):
    src = os.path.join(_HERE, "bn254_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _u64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if shape is not None:
        a = a.reshape(shape)
    return a


FQ, FR = 0, 1


def to_mont(vals, mod=FR):
    """list of python ints (canonical) -> uint64[n,4] Montgomery."""
    n = len(vals)
    canon = np.zeros((n, 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        for k in range(4):
            canon[i, k] = (int(v) >> (64 * k)) & 0xFFFFFFFFFFFFFFFF
    out = np.zeros((n, 4), dtype=np.uint64)
    L = lib()
    for i in range(n):
        L.orc_fp_from_u64x4(mod, _p(canon[i]), _p(out[i]))
    return out


def from_mont(arr, mod=FR):
    arr = _u64(arr).reshape(-1, 4)
    L = lib()
    out = []
    tmp = np.zeros(4, dtype=np.uint64)
    for i in range(arr.shape[0]):
        L.orc_fp_to_u64x4(mod, _p(arr[i]), _p(tmp))
        out.append(sum(int(tmp[k]) << (64 * k) for k in range(4)))
    return out


def points_to_ints(pts, inf):
    """uint64[n,8] + inf -> list of (x,y) ints or None."""
    pts = _u64(pts).reshape(-1, 8)
    xs = from_mont(pts[:, :4], FQ)
    ys = from_mont(pts[:, 4:], FQ)
    return [None if inf[i] else (xs[i], ys[i]) for i in range(len(xs))]


def points_from_ints(pl):
    n = len(pl)
    pts = np.zeros((n, 8), dtype=np.uint64)
    inf = np.zeros(n, dtype=np.uint8)
    xs = to_mont([0 if p is None else p[0] for p in pl], FQ)
    ys = to_mont([0 if p is None else p[1] for p in pl], FQ)
    for i, p in enumerate(pl):
        if p is None:
            inf[i] = 1
        else:
            pts[i, :4] = xs[i]
            pts[i, 4:] = ys[i]
    return pts, inf


def generator():
    g = np.zeros(8, dtype=np.uint64)
    lib().orc_g1_generator(_p(g))
    return g


def msm(pts, inf, scalars, algo=1):
    pts = _u64(pts).reshape(-1, 8)
    scalars = _u64(scalars).reshape(-1, 4)
    n = pts.shape[0]
    assert scalars.shape[0] == n
    out = np.zeros(8, dtype=np.uint64)
    oinf = np.zeros(1, dtype=np.uint8)
    infa = None if inf is None else np.ascontiguousarray(inf, dtype=np.uint8)
    lib().orc_msm(_p(pts), _p(infa), _p(scalars), C.c_size_t(n), algo, _p(out), _p(oinf))
    return out, int(oinf[0])


def scalar_mul(pt, inf, s):
    out = np.zeros(8, dtype=np.uint64)
    oinf = np.zeros(1, dtype=np.uint8)
    lib().orc_g1_scalar_mul(_p(_u64(pt)), C.c_uint8(inf), _p(_u64(s)), _p(out), _p(oinf))
    return out, int(oinf[0])


def compress(pt, inf):
    out = np.zeros(32, dtype=np.uint8)
    lib().orc_g1_compress(_p(_u64(pt)), C.c_uint8(inf), _p(out))
    return bytes(out)


def gen_scalars(label, n):
    sc = np.zeros((n + 1, 4), dtype=np.uint64)
    kinds = np.zeros(n + 1, dtype=np.uint8)
    lib().orc_gen_scalars(label, C.c_size_t(len(label)), C.c_size_t(n), _p(sc), _p(kinds))
    return sc, kinds


def multi_commit_gens(label, n):
    """-> (G uint64[n,8], h uint64[8])  (commitments.rs:31-62)."""
    out = np.zeros((n + 1, 8), dtype=np.uint64)
    lib().orc_multi_commit_gens(label, C.c_size_t(len(label)), C.c_size_t(n), _p(out))
    return out[:n].copy(), out[n].copy()


def dotproduct_gens(label, n):
    """DotProductProofGens::new (nizk/mod.rs:412-415) -> (G_n[n,8], h, G_1[8])."""
    G, h = multi_commit_gens(label, n + 1)
    return G[:n].copy(), h, G[n].copy()


def hyrax_commit(G, h, Z, L_size, R_size, blinds=None, threads=0):
    G = _u64(G).reshape(R_size, 8)
    Z = _u64(Z).reshape(L_size * R_size, 4)
    Cout = np.zeros((L_size, 8), dtype=np.uint64)
    inf = np.zeros(L_size, dtype=np.uint8)
    b = None if blinds is None else _u64(blinds).reshape(L_size, 4)
    lib().orc_hyrax_commit(_p(G), _p(_u64(h)), _p(Z), C.c_size_t(L_size), C.c_size_t(R_size), _p(b),
                           C.c_int(threads), _p(Cout), _p(inf))
    return Cout, inf


def bound(Z, Lvec, L_size, R_size, threads=0):
    Z = _u64(Z).reshape(L_size * R_size, 4)
    Lvec = _u64(Lvec).reshape(L_size, 4)
    out = np.zeros((R_size, 4), dtype=np.uint64)
    lib().orc_bound(_p(Z), _p(Lvec), C.c_size_t(L_size), C.c_size_t(R_size), C.c_int(threads), _p(out))
    return out


def eq_evals(r):
    r = _u64(r).reshape(-1, 4)
    ell = r.shape[0]
    out = np.zeros((1 << ell, 4), dtype=np.uint64)
    lib().orc_eq_evals(_p(r), C.c_size_t(ell), _p(out))
    return out


def bind_top(Z, r):
    Z = _u64(Z).reshape(-1, 4).copy()
    lib().orc_bind_top(_p(Z), C.c_size_t(Z.shape[0]), _p(_u64(r)))
    return Z[: Z.shape[0] // 2].copy()


def sumcheck_cubic_eval(A, B, Cc, D):
    A, B, Cc, D = [_u64(x).reshape(-1, 4) for x in (A, B, Cc, D)]
    e = [np.zeros(4, dtype=np.uint64) for _ in range(3)]
    lib().orc_sumcheck_cubic_eval(_p(A), _p(B), _p(Cc), _p(D), C.c_size_t(A.shape[0]), _p(e[0]), _p(e[1]), _p(e[2]))
    return e


def sumcheck_quad_eval(Zt, ABC):
    Zt, ABC = _u64(Zt).reshape(-1, 4), _u64(ABC).reshape(-1, 4)
    e = [np.zeros(4, dtype=np.uint64) for _ in range(2)]
    lib().orc_sumcheck_quad_eval(_p(Zt), _p(ABC), C.c_size_t(Zt.shape[0]), _p(e[0]), _p(e[1]))
    return e


def bullet_prove(Q, G, H, a, b, blind, blinds_L, blinds_R, u):
    G = _u64(G).reshape(-1, 8)
    n = G.shape[0]
    lg = n.bit_length() - 1
    a = _u64(a).reshape(n, 4)
    b = _u64(b).reshape(n, 4)
    Lo = np.zeros((lg, 8), dtype=np.uint64); Li = np.zeros(lg, dtype=np.uint8)
    Ro = np.zeros((lg, 8), dtype=np.uint64); Ri = np.zeros(lg, dtype=np.uint8)
    Gam = np.zeros(8, dtype=np.uint64); Gi = np.zeros(1, dtype=np.uint8)
    ah = np.zeros(4, dtype=np.uint64); bh = np.zeros(4, dtype=np.uint64)
    gh = np.zeros(8, dtype=np.uint64); ghi = np.zeros(1, dtype=np.uint8)
    blh = np.zeros(4, dtype=np.uint64)
    lib().orc_bullet_prove(_p(_u64(Q)), _p(G), C.c_size_t(n), _p(_u64(H)), _p(a), _p(b), _p(_u64(blind)),
                           _p(_u64(blinds_L).reshape(lg, 4)), _p(_u64(blinds_R).reshape(lg, 4)),
                           _p(_u64(u).reshape(lg, 4)), _p(Lo), _p(Li), _p(Ro), _p(Ri), _p(Gam), _p(Gi),
                           _p(ah), _p(bh), _p(gh), _p(ghi), _p(blh))
    return dict(L=Lo, L_inf=Li, R=Ro, R_inf=Ri, Gamma=Gam, Gamma_inf=int(Gi[0]), a_hat=ah, b_hat=bh,
                g_hat=gh, g_hat_inf=int(ghi[0]), blind_hat=blh)


class Transcript:
    """Merlin transcript (third-party merlin 3.0) as used by transcript.rs."""

    class _T(C.Structure):
        _fields_ = [("st", C.c_uint8 * 200), ("pos", C.c_uint8), ("pos_begin", C.c_uint8), ("cur_flags", C.c_uint8)]

    def __init__(self, label):
        self.t = Transcript._T()
        lib().orc_transcript_new(C.byref(self.t), label, C.c_size_t(len(label)))

    def append_message(self, label, msg):
        lib().orc_transcript_append(C.byref(self.t), label, C.c_size_t(len(label)), msg, C.c_size_t(len(msg)))

    def challenge_bytes(self, label, n):
        buf = (C.c_uint8 * n)()
        lib().orc_transcript_challenge(C.byref(self.t), label, C.c_size_t(len(label)), buf, C.c_size_t(n))
        return bytes(buf)

    def challenge_scalar(self, label):
        out = np.zeros(4, dtype=np.uint64)
        lib().orc_transcript_challenge_scalar(C.byref(self.t), label, C.c_size_t(len(label)), _p(out))
        return out


class EvalProof(C.Structure):
    """orc_eval_proof: PolyEvalProof { DotProductProofLog { bullet L/R, delta, beta, z1, z2 } }."""
    _fields_ = [("lg_n", C.c_size_t), ("L", (C.c_uint64 * 8) * 32), ("R", (C.c_uint64 * 8) * 32),
                ("L_inf", C.c_uint8 * 32), ("R_inf", C.c_uint8 * 32), ("delta", C.c_uint64 * 8), ("beta", C.c_uint64 * 8),
                ("delta_inf", C.c_uint8), ("beta_inf", C.c_uint8), ("z1", C.c_uint64 * 4), ("z2", C.c_uint64 * 4)]

    def to_dict(self):
        lg = self.lg_n
        return dict(L=[(np.array(self.L[i], dtype=np.uint64), int(self.L_inf[i])) for i in range(lg)],
                    R=[(np.array(self.R[i], dtype=np.uint64), int(self.R_inf[i])) for i in range(lg)],
                    delta=(np.array(self.delta, dtype=np.uint64), int(self.delta_inf)),
                    beta=(np.array(self.beta, dtype=np.uint64), int(self.beta_inf)),
                    z1=np.array(self.z1, dtype=np.uint64), z2=np.array(self.z2, dtype=np.uint64))

    @staticmethod
    def from_dict(d):
        p = EvalProof()
        p.lg_n = len(d["L"])
        for i, ((l, li), (r, ri)) in enumerate(zip(d["L"], d["R"])):
            for k in range(8):
                p.L[i][k] = int(l[k]); p.R[i][k] = int(r[k])
            p.L_inf[i] = li; p.R_inf[i] = ri
        for k in range(8):
            p.delta[k] = int(d["delta"][0][k]); p.beta[k] = int(d["beta"][0][k])
        p.delta_inf = d["delta"][1]; p.beta_inf = d["beta"][1]
        for k in range(4):
            p.z1[k] = int(d["z1"][k]); p.z2[k] = int(d["z2"][k])
        return p


def poly_eval_prove(Z, ell, blinds, r, Zr, blind_Zr, G, h, G1, transcript, tape):
    """PolyEvalProof::prove (hyrax.rs:65-116) on the CPU; returns (EvalProof, C_Zr', inf)."""
    Z = _u64(Z).reshape(-1, 4)
    proof = EvalProof()
    Cz = np.zeros(8, dtype=np.uint64); Czi = np.zeros(1, dtype=np.uint8)
    b = None if blinds is None else _u64(blinds).reshape(-1, 4)
    bz = None if blind_Zr is None else _u64(blind_Zr)
    lib().orc_poly_eval_prove(_p(Z), C.c_size_t(ell), _p(b), _p(_u64(r).reshape(-1, 4)), _p(_u64(Zr)), _p(bz),
                              _p(_u64(G).reshape(-1, 8)), _p(_u64(h)), _p(_u64(G1)), C.byref(transcript.t),
                              C.byref(tape.t), C.byref(proof), _p(Cz), _p(Czi))
    return proof, Cz, int(Czi[0])


def poly_eval_verify(proof, ell, r, C_Zr, C_Zr_inf, comm, comm_inf, G, h, G1, transcript):
    """PolyEvalProof::verify (hyrax.rs:118-137) on the CPU; True when the proof verifies."""
    comm = _u64(comm).reshape(-1, 8)
    ci = np.ascontiguousarray(comm_inf, dtype=np.uint8)
    return bool(lib().orc_poly_eval_verify(C.byref(proof), C.c_size_t(ell), _p(_u64(r).reshape(-1, 4)), _p(_u64(C_Zr)),
                                           C.c_uint8(C_Zr_inf), _p(comm), _p(ci), _p(_u64(G).reshape(-1, 8)), _p(_u64(h)),
                                           _p(_u64(G1)), C.byref(transcript.t)))


def evaluate(Z, r):
    """DensePolynomial::evaluate (hyrax.rs:217-222): <Z, eq(r)>."""
    Z = _u64(Z).reshape(-1, 4)
    chis = eq_evals(r)
    acc = 0
    zs, cs = from_mont(Z), from_mont(chis)
    R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
    for a, b in zip(zs, cs):
        acc = (acc + a * b) % R_MOD
    return to_mont([acc])[0]


def prove_workload(log_cons, threads=0, derefs_rows=0):
    """CPU baseline of the end-to-end prove (orc_prove_workload): -> dict(seconds=total, phases={name: s}, scaled=bool,
    gens_seconds=one-off generator derivation)."""
    L = lib()
    L.orc_prove_workload_phase_name.restype = C.c_char_p
    n = L.orc_prove_workload_phases()
    sec = (C.c_double * n)()
    gens = C.c_double()
    rc = L.orc_prove_workload(C.c_int(log_cons), C.c_int(threads), C.c_size_t(derefs_rows), sec, C.byref(gens))
    if rc < 0:
        raise ValueError("orc_prove_workload: bad argument")
    phases = {L.orc_prove_workload_phase_name(i).decode(): float(sec[i]) for i in range(n)}
    return dict(seconds=sum(phases.values()), phases=phases, scaled=bool(rc), gens_seconds=float(gens.value))
