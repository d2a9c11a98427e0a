"""ctypes binding of the libsbn254 C ABI (include/sbn254.h).  numpy arrays carry the ABI layouts:
scalars uint64[n,4] (Montgomery Fr), points uint64[n,8] (affine Montgomery x|y) + uint8[n] inf."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lib", "libsbn254.so")

EXPORTS = [
    "sbn_strerror", "sbn_last_cuda_error", "sbn_version",
    "sbn_ctx_create", "sbn_ctx_destroy", "sbn_ctx_synchronize", "sbn_ctx_set", "sbn_ctx_counters",
    "sbn_ctx_last_commit_profile", "sbn_ctx_memory_stats", "sbn_host_alloc", "sbn_host_free", "sbn_stream_create", "sbn_stream_synchronize", "sbn_stream_destroy",
    "sbn_bases_create", "sbn_bases_create_ext", "sbn_bases_destroy", "sbn_bases_len", "sbn_bases_window_bits", "sbn_bases_mult_table",
    "sbn_hyrax_commit", "sbn_hyrax_commit_async", "sbn_hyrax_commit_device", "sbn_hyrax_commit_multi", "sbn_msm", "sbn_commit",
    "sbn_g1_scalar_mul_batch", "sbn_g1_scale_points", "sbn_bound",
    "sbn_poly_upload", "sbn_poly_destroy", "sbn_poly_commit", "sbn_poly_commit_rows", "sbn_poly_bound",
    "sbn_bullet_begin", "sbn_bullet_round", "sbn_bullet_fold", "sbn_bullet_end", "sbn_bullet_end_delta", "sbn_bullet_destroy",
    "sbn_sumcheck_begin", "sbn_sumcheck_begin_quad", "sbn_sumcheck_round_eval", "sbn_sumcheck_bind", "sbn_sumcheck_end",
    "sbn_sumcheck_destroy", "sbn_fr_from_canonical", "sbn_fr_to_canonical", "sbn_microbench",
    "sbn_spmat_upload", "sbn_spmat_destroy", "sbn_spmat_mulvec", "sbn_eq_evals",
    "sbn_spark_evaluate", "sbn_bsumcheck_begin_resident", "sbn_spark_comb_polys", "sbn_poly_triple_dot",
    "sbn_poly_evaluate", "sbn_poly_evaluate_strided", "sbn_addrs_set_timestamps", "sbn_hashlayer_build", "sbn_prodcircuit_download_layer",
    "sbn_derefs_commit_rows", "sbn_keccak_f1600", "sbn_fr_to_canonical_host", "sbn_fr_from_canonical_host", "sbn_sumcheck_begin_r1cs", "sbn_sumcheck_begin_r1cs_resident", "sbn_sumcheck_begin_quad_r1cs", "sbn_g1_compress", "sbn_merlin_append_points", "sbn_merlin_init", "sbn_merlin_append", "sbn_merlin_append_many", "sbn_merlin_challenge", "sbn_merlin_challenge_scalars", "sbn_addrs_upload", "sbn_addrs_destroy", "sbn_derefs_commit", "sbn_poly_len", "sbn_poly_download",
    "sbn_prodcircuit_create", "sbn_prodcircuit_evaluate", "sbn_prodcircuit_num_layers", "sbn_prodcircuit_destroy",
    "sbn_bsumcheck_begin", "sbn_bsumcheck_round_eval", "sbn_bsumcheck_bind", "sbn_bsumcheck_end", "sbn_bsumcheck_prove", "sbn_bsumcheck_destroy",
]


class SbnError(RuntimeError):
    def __init__(self, status, what, detail=""):
        self.status = status
        super().__init__(f"{what}: status {status}" + (f" ({detail})" if detail else ""))


_lib = None


def load_library():
    """Loads libsbn254.so; raises (never falls back) when the CUDA library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SbnError(-3, "libsbn254.so missing",
                       f"build it with `python -m spartan_bn254_b200.build` (expected at {LIB_PATH})")
    lib = C.CDLL(LIB_PATH)
    lib.sbn_strerror.restype = C.c_char_p
    lib.sbn_last_cuda_error.restype = C.c_char_p
    lib.sbn_last_cuda_error.argtypes = [C.c_void_p]
    lib.sbn_bases_len.restype = C.c_size_t
    lib.sbn_bases_len.argtypes = [C.c_void_p]
    lib.sbn_bases_window_bits.argtypes = [C.c_void_p]
    lib.sbn_poly_len.restype = C.c_size_t
    lib.sbn_poly_len.argtypes = [C.c_void_p]
    lib.sbn_prodcircuit_num_layers.restype = C.c_size_t
    lib.sbn_prodcircuit_num_layers.argtypes = [C.c_void_p]
    _lib = lib
    return lib


def _ptr(a):
    if a is None:
        return None
    return C.c_void_p(a.ctypes.data)        # three times cheaper than data_as(); the prover makes ~1700 of these per proof


def _u64(a, cols):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    return a.reshape(-1, cols)


_live_contexts = []          # most recent last: the host mirrors' bulk Fr conversions borrow one (hyrax._bulk_ctx)


class Context:
    """sbn_ctx: one CUDA device."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = C.c_void_p()
        st = self.lib.sbn_ctx_create(C.c_int(device), C.byref(h))
        if st != 0:
            raise SbnError(st, "sbn_ctx_create", self.lib.sbn_strerror(st).decode() +
                           " -- a CUDA device is required, there is no CPU fallback")
        self.h = h
        self.device = device
        _live_contexts.append(self)

    def _check(self, st, what):
        if st != 0:
            detail = self.lib.sbn_strerror(st).decode()
            cuda = self.lib.sbn_last_cuda_error(self.h).decode()
            raise SbnError(st, what, detail + (": " + cuda if cuda and st in (-3, -4) else ""))

    def close(self):
        if self in _live_contexts:
            _live_contexts.remove(self)
        if self.h:
            self.lib.sbn_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set(self, key, value):
        self._check(self.lib.sbn_ctx_set(self.h, key.encode(), C.c_long(value)), "sbn_ctx_set")

    def synchronize(self):
        self._check(self.lib.sbn_ctx_synchronize(self.h), "sbn_ctx_synchronize")

    def counters(self, reset=False):
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._check(self.lib.sbn_ctx_counters(self.h, C.byref(a), C.byref(b), C.byref(c), C.c_int(int(reset))),
                    "sbn_ctx_counters")
        return dict(kernel_launches=a.value, h2d_bytes=b.value, d2h_bytes=c.value)

    def memory_stats(self):
        out = (C.c_uint64 * 8)()
        self._check(self.lib.sbn_ctx_memory_stats(self.h, out), "sbn_ctx_memory_stats")
        return dict(pool_bytes=int(out[0]), pool_flushes=int(out[1]), mult_table_fallbacks=int(out[2]), pool_buffers=int(out[3]),
                    small_scalar_commits=int(out[4]))

    def last_commit_profile(self):
        ms = (C.c_float * 4)()
        n = (C.c_int * 4)()
        self._check(self.lib.sbn_ctx_last_commit_profile(self.h, ms, n), "sbn_ctx_last_commit_profile")
        names = ["sort", "accumulate", "reduce", "normalize"]
        return {names[i]: dict(ms=float(ms[i]), launches=int(n[i])) for i in range(4)}

    # ---- generators
    def bases(self, G, h, G_inf=None, g1=None):
        return Bases(self, G, h, G_inf, g1)

    # ---- a7: DensePolynomial::commit_inner
    def hyrax_commit(self, bases, Z, L_size, R_size, blinds=None):
        Z = _u64(Z, 4)
        if Z.shape[0] != L_size * R_size:
            raise SbnError(-2, "sbn_hyrax_commit", "len(Z) != L_size * R_size")
        bl = None if blinds is None else _u64(blinds, 4)
        if bl is not None and bl.shape[0] != L_size:
            raise SbnError(-2, "sbn_hyrax_commit", "len(blinds) != L_size")
        out = np.zeros((L_size, 8), dtype=np.uint64)
        inf = np.zeros(L_size, dtype=np.uint8)
        st = self.lib.sbn_hyrax_commit(self.h, bases.h, _ptr(Z), C.c_size_t(L_size), C.c_size_t(R_size), _ptr(bl),
                                       _ptr(out), _ptr(inf))
        self._check(st, "sbn_hyrax_commit")
        return out, inf

    def hyrax_commit_raw(self, bases, Z_ptr, L_size, R_size, blinds_ptr, out_ptr, inf_ptr):
        """Host-pointer call without numpy marshalling (bench e2e: pinned buffers)."""
        st = self.lib.sbn_hyrax_commit(self.h, bases.h, C.c_void_p(Z_ptr), C.c_size_t(L_size), C.c_size_t(R_size),
                                       C.c_void_p(blinds_ptr) if blinds_ptr else None, C.c_void_p(out_ptr),
                                       C.c_void_p(inf_ptr))
        self._check(st, "sbn_hyrax_commit")

    def stream_create(self):
        """A non-blocking CUDA stream of the library's own (sbn_stream_create): the handle the asynchronous calls take."""
        h = C.c_void_p()
        self._check(self.lib.sbn_stream_create(self.h, C.byref(h)), "sbn_stream_create")
        return h.value

    def stream_synchronize(self, stream):
        self._check(self.lib.sbn_stream_synchronize(self.h, C.c_void_p(stream)), "sbn_stream_synchronize")

    def stream_destroy(self, stream):
        self._check(self.lib.sbn_stream_destroy(self.h, C.c_void_p(stream)), "sbn_stream_destroy")

    def hyrax_commit_raw_async(self, bases, Z_ptr, L_size, R_size, blinds_ptr, out_ptr, inf_ptr, stream):
        """sbn_hyrax_commit_async on raw host pointers (pinned for real asynchrony); the caller synchronises `stream`."""
        st = self.lib.sbn_hyrax_commit_async(self.h, bases.h, C.c_void_p(Z_ptr), C.c_size_t(L_size), C.c_size_t(R_size),
                                             C.c_void_p(blinds_ptr) if blinds_ptr else None, C.c_void_p(out_ptr),
                                             C.c_void_p(inf_ptr), C.c_void_p(stream))
        self._check(st, "sbn_hyrax_commit_async")

    def hyrax_commit_device(self, bases, dZ_ptr, L_size, R_size, dblinds_ptr, dC_ptr, dinf_ptr, stream=0):
        st = self.lib.sbn_hyrax_commit_device(self.h, bases.h, C.c_void_p(dZ_ptr), C.c_size_t(L_size),
                                              C.c_size_t(R_size), C.c_void_p(dblinds_ptr) if dblinds_ptr else None,
                                              C.c_void_p(dC_ptr), C.c_void_p(dinf_ptr) if dinf_ptr else None,
                                              C.c_void_p(stream) if stream else None)
        self._check(st, "sbn_hyrax_commit_device")

    # ---- a6 / a5
    def msm(self, points, inf, scalars):
        points = _u64(points, 8)
        scalars = _u64(scalars, 4)
        n = points.shape[0]
        if scalars.shape[0] != n:
            # group.rs:156,173 `.unwrap_or_default()`: a length mismatch yields the identity
            return np.zeros(8, dtype=np.uint64), 1
        infa = None if inf is None else np.ascontiguousarray(inf, dtype=np.uint8)
        out = np.zeros(8, dtype=np.uint64)
        oinf = np.zeros(1, dtype=np.uint8)
        st = self.lib.sbn_msm(self.h, _ptr(points), _ptr(infa), _ptr(scalars), C.c_size_t(n), _ptr(out), _ptr(oinf))
        self._check(st, "sbn_msm")
        return out, int(oinf[0])

    def commit(self, bases, scalars, blind):
        scalars = _u64(scalars, 4)
        out = np.zeros(8, dtype=np.uint64)
        oinf = np.zeros(1, dtype=np.uint8)
        st = self.lib.sbn_commit(self.h, bases.h, _ptr(scalars), C.c_size_t(scalars.shape[0]),
                                 _ptr(_u64(blind, 4)), _ptr(out), _ptr(oinf))
        self._check(st, "sbn_commit")
        return out, int(oinf[0])

    def scalar_mul_batch(self, P, scalars):
        scalars = _u64(scalars, 4)
        n = scalars.shape[0]
        out = np.zeros((n, 8), dtype=np.uint64)
        inf = np.zeros(n, dtype=np.uint8)
        st = self.lib.sbn_g1_scalar_mul_batch(self.h, _ptr(_u64(P, 8)), _ptr(scalars), C.c_size_t(n), _ptr(out), _ptr(inf))
        self._check(st, "sbn_g1_scalar_mul_batch")
        return out, inf

    def scale_points(self, P, inf, s):
        P = _u64(P, 8)
        n = P.shape[0]
        infa = None if inf is None else np.ascontiguousarray(inf, dtype=np.uint8)
        out = np.zeros((n, 8), dtype=np.uint64)
        oinf = np.zeros(n, dtype=np.uint8)
        st = self.lib.sbn_g1_scale_points(self.h, _ptr(P), _ptr(infa), C.c_size_t(n), _ptr(_u64(s, 4)), _ptr(out), _ptr(oinf))
        self._check(st, "sbn_g1_scale_points")
        return out, oinf

    # ---- a11
    def bound(self, Z, Lvec, L_size, R_size):
        Z = _u64(Z, 4)
        Lvec = _u64(Lvec, 4)
        if Z.shape[0] != L_size * R_size or Lvec.shape[0] != L_size:
            raise SbnError(-2, "sbn_bound", "shape mismatch")
        out = np.zeros((R_size, 4), dtype=np.uint64)
        st = self.lib.sbn_bound(self.h, _ptr(Z), _ptr(Lvec), C.c_size_t(L_size), C.c_size_t(R_size), _ptr(out))
        self._check(st, "sbn_bound")
        return out

    # ---- resident polynomial
    def poly_upload(self, Z):
        return Poly(self, Z)

    # ---- a14: bullet reduction (state on device, transcript on host)
    def bullet_begin(self, bases, Q, a, b, blind, q_scalar=None):
        """Q: arbitrary point (generators are folded explicitly) -- or Q=None with q_scalar: Q = q_scalar * g1 of
        bases created with g1 (table-based rounds)."""
        return BulletState(self, bases, Q, a, b, blind, q_scalar)

    # ---- a16: sumcheck rounds
    def sumcheck_begin(self, tau, Az, Bz, Cz):
        return SumcheckState(self, tau, Az, Bz, Cz)

    def sumcheck_begin_quad(self, z, ABC):
        return SumcheckState(self, z, ABC)

    def sumcheck_begin_r1cs(self, mats, z, tau):
        """Phase-1 tables eq(tau), A z, B z, C z built on the device from the resident matrices (r1csproof.rs:268-290)."""
        return SumcheckState._resident(self, 4, mats, z, tau, None)

    def sumcheck_begin_r1cs_resident(self, mats, vars_poly, tail, z_len, tau):
        """sumcheck_begin_r1cs with z = (vars, tail, 0...) assembled on the device from the resident witness polynomial."""
        st = SumcheckState.__new__(SumcheckState)
        st.ctx, st.ntables = self, 4
        tau = _u64(tau, 4)
        tail = _u64(tail, 4)
        st.len = 1 << tau.shape[0]
        hs = (C.c_void_p * 3)(*[m.h for m in mats])
        h = C.c_void_p()
        rc = self.lib.sbn_sumcheck_begin_r1cs_resident(self.h, hs, vars_poly.h, _ptr(tail), C.c_size_t(tail.shape[0]), C.c_size_t(z_len),
                                                       _ptr(tau), C.c_size_t(tau.shape[0]), C.byref(h))
        self._check(rc, "sbn_sumcheck_begin_r1cs_resident")
        st.h = h
        return st

    def sumcheck_begin_quad_r1cs(self, mats_t, coeffs, rx, z, z_len=None):
        """Phase-2 tables z and sum_m coeffs[m] M_m^T eq(rx) built on the device (r1csproof.rs:378-410).  z=None with z_len:
        the z uploaded by the preceding sumcheck_begin_r1cs is still resident."""
        return SumcheckState._resident(self, 2, mats_t, z, rx, coeffs, z_len)

    # ---- utilities
    def fr_from_canonical(self, canon):
        canon = _u64(canon, 4)
        out = np.zeros_like(canon)
        self._check(self.lib.sbn_fr_from_canonical(self.h, _ptr(canon), C.c_size_t(canon.shape[0]), _ptr(out)),
                    "sbn_fr_from_canonical")
        return out

    def fr_to_canonical(self, mont):
        mont = _u64(mont, 4)
        out = np.zeros_like(mont)
        self._check(self.lib.sbn_fr_to_canonical(self.h, _ptr(mont), C.c_size_t(mont.shape[0]), _ptr(out)),
                    "sbn_fr_to_canonical")
        return out

    def microbench(self, kind):
        v = C.c_double()
        self._check(self.lib.sbn_microbench(self.h, C.c_int(kind), C.byref(v)), "sbn_microbench")
        return v.value


class Bases:
    """sbn_bases: a MultiCommitGens (G[0..n) + h) resident in HBM with its window tables."""

    def __init__(self, ctx, G, h, G_inf=None, g1=None):
        self.ctx = ctx
        G = _u64(G, 8)
        h = _u64(h, 8)
        infa = None if G_inf is None else np.ascontiguousarray(G_inf, dtype=np.uint8)
        hd = C.c_void_p()
        if g1 is None:
            st = ctx.lib.sbn_bases_create(ctx.h, _ptr(G), _ptr(infa), C.c_size_t(G.shape[0]), _ptr(h), C.byref(hd))
        else:
            st = ctx.lib.sbn_bases_create_ext(ctx.h, _ptr(G), _ptr(infa), C.c_size_t(G.shape[0]), _ptr(_u64(g1, 8)), _ptr(h),
                                              C.byref(hd))
        ctx._check(st, "sbn_bases_create")
        self.h = hd
        self.n = G.shape[0]

    @property
    def window_bits(self):
        return self.ctx.lib.sbn_bases_window_bits(self.h)

    def mult_table(self):
        """(window bits, bytes) of the digit-multiple table many-row commits sum over; (0, 0) when none has been built."""
        c, nbytes = C.c_int(), C.c_uint64()
        self.ctx._check(self.ctx.lib.sbn_bases_mult_table(self.h, C.byref(c), C.byref(nbytes)), "sbn_bases_mult_table")
        return int(c.value), int(nbytes.value)

    def close(self):
        if self.h and self.ctx.h:
            self.ctx.lib.sbn_bases_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class BulletState:
    """sbn_bullet: G, a, b of BulletReductionProof::prove (nizk/bullet.rs:24-126) resident on the GPU."""

    def __init__(self, ctx, bases, Q, a, b, blind, q_scalar=None):
        self.ctx = ctx
        a = _u64(a, 4)
        b = _u64(b, 4)
        if a.shape[0] != b.shape[0]:
            raise SbnError(-2, "sbn_bullet_begin", "len(a) != len(b)")      # bullet.rs:42-43
        self.n = a.shape[0]
        self.Gamma = np.zeros(8, dtype=np.uint64)
        ginf = np.zeros(1, dtype=np.uint8)
        h = C.c_void_p()
        qp = None if Q is None else _u64(Q, 8)
        qs = None if q_scalar is None else _u64(q_scalar, 4)
        st = ctx.lib.sbn_bullet_begin(ctx.h, bases.h, _ptr(qp), _ptr(qs), _ptr(a), _ptr(b), C.c_size_t(self.n),
                                      _ptr(_u64(blind, 4)), _ptr(self.Gamma), _ptr(ginf), C.byref(h))
        ctx._check(st, "sbn_bullet_begin")
        self.Gamma_inf = int(ginf[0])
        self.h = h

    def round(self, blind_L, blind_R):
        L = np.zeros(8, dtype=np.uint64); Li = np.zeros(1, dtype=np.uint8)
        R = np.zeros(8, dtype=np.uint64); Ri = np.zeros(1, dtype=np.uint8)
        st = self.ctx.lib.sbn_bullet_round(self.h, _ptr(_u64(blind_L, 4)), _ptr(_u64(blind_R, 4)), _ptr(L), _ptr(Li),
                                           _ptr(R), _ptr(Ri))
        self.ctx._check(st, "sbn_bullet_round")
        return (L, int(Li[0])), (R, int(Ri[0]))

    def fold(self, u, u_inv):
        self.ctx._check(self.ctx.lib.sbn_bullet_fold(self.h, _ptr(_u64(u, 4)), _ptr(_u64(u_inv, 4))), "sbn_bullet_fold")
        self.n //= 2

    def end(self):
        a = np.zeros(4, dtype=np.uint64); b = np.zeros(4, dtype=np.uint64)
        g = np.zeros(8, dtype=np.uint64); gi = np.zeros(1, dtype=np.uint8)
        self.ctx._check(self.ctx.lib.sbn_bullet_end(self.h, _ptr(a), _ptr(b), _ptr(g), _ptr(gi)), "sbn_bullet_end")
        return a, b, g, int(gi[0])

    def end_delta(self, d, r_delta):
        """end() plus delta = d * g_hat + r_delta * h (nizk/mod.rs:497-500) as a second row over the resident tables."""
        a = np.zeros(4, dtype=np.uint64); b = np.zeros(4, dtype=np.uint64)
        g = np.zeros(8, dtype=np.uint64); gi = np.zeros(1, dtype=np.uint8)
        dl = np.zeros(8, dtype=np.uint64); di = np.zeros(1, dtype=np.uint8)
        self.ctx._check(self.ctx.lib.sbn_bullet_end_delta(self.h, _ptr(_u64(d, 4)), _ptr(_u64(r_delta, 4)), _ptr(a), _ptr(b), _ptr(g),
                                                          _ptr(gi), _ptr(dl), _ptr(di)), "sbn_bullet_end_delta")
        return a, b, g, int(gi[0]), dl, int(di[0])

    def close(self):
        if self.h and self.ctx.h:
            self.ctx.lib.sbn_bullet_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SumcheckState:
    """sbn_sumcheck: the four tables of the R1CS-sat cubic sumcheck (sumcheck.rs:465-649) on the GPU."""

    @staticmethod
    def _resident(ctx, ntables, mats, z, point, coeffs, z_len=None):
        self = SumcheckState.__new__(SumcheckState)
        self.ctx, self.ntables = ctx, ntables
        point = _u64(point, 4)
        z = None if z is None else _u64(z, 4)
        z_len = z.shape[0] if z is not None else int(z_len)
        hs = (C.c_void_p * 3)(*[m.h for m in mats])
        h = C.c_void_p()
        if ntables == 4:
            self.len = 1 << point.shape[0]
            st = ctx.lib.sbn_sumcheck_begin_r1cs(ctx.h, hs, _ptr(z), C.c_size_t(z.shape[0]), _ptr(point), C.c_size_t(point.shape[0]),
                                                 C.byref(h))
            ctx._check(st, "sbn_sumcheck_begin_r1cs")
        else:
            self.len = z_len
            st = ctx.lib.sbn_sumcheck_begin_quad_r1cs(ctx.h, hs, _ptr(_u64(coeffs, 4)), _ptr(point), C.c_size_t(point.shape[0]),
                                                      _ptr(z), C.c_size_t(z_len), C.byref(h))
            ctx._check(st, "sbn_sumcheck_begin_quad_r1cs")
        self.h = h
        return self

    def __init__(self, ctx, *tables):
        self.ctx = ctx
        t = [_u64(x, 4) for x in tables]
        self.ntables = len(t)
        self.len = t[0].shape[0]
        if any(x.shape[0] != self.len for x in t):
            raise SbnError(-2, "sbn_sumcheck_begin", "table lengths differ")
        h = C.c_void_p()
        if self.ntables == 4:
            st = ctx.lib.sbn_sumcheck_begin(ctx.h, _ptr(t[0]), _ptr(t[1]), _ptr(t[2]), _ptr(t[3]), C.c_size_t(self.len),
                                            C.byref(h))
        else:
            st = ctx.lib.sbn_sumcheck_begin_quad(ctx.h, _ptr(t[0]), _ptr(t[1]), C.c_size_t(self.len), C.byref(h))
        ctx._check(st, "sbn_sumcheck_begin")
        self.h = h

    def round_eval(self):
        n = 3 if self.ntables == 4 else 2
        e = [np.zeros(4, dtype=np.uint64) for _ in range(3)]
        self.ctx._check(self.ctx.lib.sbn_sumcheck_round_eval(self.h, _ptr(e[0]), _ptr(e[1]), _ptr(e[2])),
                        "sbn_sumcheck_round_eval")
        return e[:n]

    def bind(self, r):
        self.ctx._check(self.ctx.lib.sbn_sumcheck_bind(self.h, _ptr(_u64(r, 4))), "sbn_sumcheck_bind")
        self.len //= 2

    def end(self):
        f = np.zeros((4, 4), dtype=np.uint64)
        self.ctx._check(self.ctx.lib.sbn_sumcheck_end(self.h, _ptr(f)), "sbn_sumcheck_end")
        return f

    def close(self):
        if self.h and self.ctx.h:
            self.ctx.lib.sbn_sumcheck_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Poly:
    """sbn_poly: the evaluation vector of a DensePolynomial resident in HBM (commit and bound reuse it)."""

    def __init__(self, ctx, Z):
        self.ctx = ctx
        Z = _u64(Z, 4)
        self.len = Z.shape[0]
        h = C.c_void_p()
        ctx._check(ctx.lib.sbn_poly_upload(ctx.h, _ptr(Z), C.c_size_t(self.len), C.byref(h)), "sbn_poly_upload")
        self.h = h

    def commit(self, bases, L_size, R_size, blinds=None):
        bl = None if blinds is None else _u64(blinds, 4)
        out = np.zeros((L_size, 8), dtype=np.uint64)
        inf = np.zeros(L_size, dtype=np.uint8)
        st = self.ctx.lib.sbn_poly_commit(self.ctx.h, bases.h, self.h, C.c_size_t(L_size), C.c_size_t(R_size), _ptr(bl),
                                          _ptr(out), _ptr(inf))
        self.ctx._check(st, "sbn_poly_commit")
        return out, inf

    def commit_rows(self, bases, first_row, n_rows, R_size, blinds=None):
        """sbn_poly_commit_rows: rows [first_row, first_row + n_rows) of the commitment (blinds: one per committed row)."""
        bl = None if blinds is None else _u64(blinds, 4)
        out = np.zeros((n_rows, 8), dtype=np.uint64)
        inf = np.zeros(n_rows, dtype=np.uint8)
        st = self.ctx.lib.sbn_poly_commit_rows(self.ctx.h, bases.h, self.h, C.c_size_t(first_row), C.c_size_t(n_rows), C.c_size_t(R_size),
                                               _ptr(bl), _ptr(out), _ptr(inf))
        self.ctx._check(st, "sbn_poly_commit_rows")
        return out, inf

    def evaluate_strided(self, r, offset0, stride, count):
        """DensePolynomial::evaluate of `count` segments of 2^len(r) evaluations starting at offset0 + i * stride, at one point:
        uint64[count, 4]."""
        r = _u64(r, 4)
        out = np.zeros((count, 4), dtype=np.uint64)
        st = self.ctx.lib.sbn_poly_evaluate_strided(self.ctx.h, self.h, C.c_size_t(offset0), C.c_size_t(stride), C.c_size_t(count),
                                                    _ptr(r), C.c_size_t(r.shape[0]), _ptr(out))
        self.ctx._check(st, "sbn_poly_evaluate_strided")
        return out

    def evaluate(self, r, offset=0):
        """DensePolynomial::evaluate of the 2^len(r) evaluations starting at `offset`."""
        r = _u64(r, 4)
        out = np.zeros(4, dtype=np.uint64)
        st = self.ctx.lib.sbn_poly_evaluate(self.ctx.h, self.h, C.c_size_t(offset), _ptr(r), C.c_size_t(r.shape[0]), _ptr(out))
        self.ctx._check(st, "sbn_poly_evaluate")
        return out

    @staticmethod
    def triple_dot(A, offA, B, offB, Cp, offC, n):
        """sum_i A[offA + i] * B[offB + i] * C[offC + i] over resident polynomials."""
        out = np.zeros(4, dtype=np.uint64)
        st = A.ctx.lib.sbn_poly_triple_dot(A.ctx.h, A.h, C.c_size_t(offA), B.h, C.c_size_t(offB), Cp.h, C.c_size_t(offC),
                                           C.c_size_t(n), _ptr(out))
        A.ctx._check(st, "sbn_poly_triple_dot")
        return out

    def download(self):
        out = np.zeros((self.len, 4), dtype=np.uint64)
        self.ctx._check(self.ctx.lib.sbn_poly_download(self.ctx.h, self.h, _ptr(out)), "sbn_poly_download")
        return out

    def bound(self, Lvec, L_size, R_size):
        Lvec = _u64(Lvec, 4)
        if Lvec.shape[0] != L_size:
            raise SbnError(-2, "sbn_poly_bound", "len(L) != L_size")
        out = np.zeros((R_size, 4), dtype=np.uint64)
        st = self.ctx.lib.sbn_poly_bound(self.ctx.h, self.h, _ptr(Lvec), C.c_size_t(L_size), C.c_size_t(R_size), _ptr(out))
        self.ctx._check(st, "sbn_poly_bound")
        return out

    def close(self):
        if self.h and self.ctx.h:
            self.ctx.lib.sbn_poly_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ProdCircuit:
    """sbn_prodcircuit: every layer of a ProductCircuit (product_tree.rs:39-57) resident in HBM."""

    def __init__(self, ctx, poly):
        self.ctx = ctx
        Z = _u64(poly, 4)
        self.len = Z.shape[0]
        h = C.c_void_p()
        ctx._check(ctx.lib.sbn_prodcircuit_create(ctx.h, _ptr(Z), C.c_size_t(self.len), C.byref(h)), "sbn_prodcircuit_create")
        self.h = h
        self.num_layers = int(ctx.lib.sbn_prodcircuit_num_layers(h))

    @classmethod
    def _from_handle(cls, ctx, h, length):
        self = cls.__new__(cls)
        self.ctx, self.h, self.len = ctx, h, length
        self.num_layers = int(ctx.lib.sbn_prodcircuit_num_layers(h))
        return self

    def evaluate(self):
        out = np.zeros(4, dtype=np.uint64)
        self.ctx._check(self.ctx.lib.sbn_prodcircuit_evaluate(self.h, _ptr(out)), "sbn_prodcircuit_evaluate")
        return out

    def layer(self, l):
        out = np.zeros((self.len >> l, 4), dtype=np.uint64)
        self.ctx._check(self.ctx.lib.sbn_prodcircuit_download_layer(self.h, C.c_size_t(l), _ptr(out)),
                        "sbn_prodcircuit_download_layer")
        return out

    def close(self):
        if self.h and self.ctx.h:
            self.ctx.lib.sbn_prodcircuit_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class BatchedSumcheckState:
    """sbn_bsumcheck: the tables of one layer's batched cubic sumcheck (sumcheck.rs:165-330) on the GPU."""

    def __init__(self, ctx, circuits, layer_id, rand, seq=(), seq_resident=()):
        """seq: host tables (left, right, weight) per sequential instance; seq_resident: the same as segments of resident
        polynomials, ((Poly, offset), (Poly, offset), (Poly, offset)) per instance (not both)."""
        self.ctx = ctx
        self.P, self.S = len(circuits), len(seq) + len(seq_resident)
        rand = np.zeros((0, 4), dtype=np.uint64) if rand is None or len(rand) == 0 else _u64(rand, 4)
        self.len = 1 << rand.shape[0]
        handles = (C.c_void_p * self.P)(*[c.h for c in circuits])
        if seq_resident:
            if seq:
                raise SbnError(-1, "sbn_bsumcheck_begin", "host and resident sequential instances cannot be mixed")
            flat = [pair for inst in seq_resident for pair in inst]
            polys = (C.c_void_p * len(flat))(*[p.h for p, _ in flat])
            offs = (C.c_size_t * len(flat))(*[int(o) for _, o in flat])
            h = C.c_void_p()
            st = ctx.lib.sbn_bsumcheck_begin_resident(ctx.h, handles, C.c_size_t(self.P), C.c_size_t(layer_id),
                                                      _ptr(rand) if rand.shape[0] else None, C.c_size_t(rand.shape[0]), polys, offs,
                                                      C.c_size_t(self.S), C.byref(h))
            ctx._check(st, "sbn_bsumcheck_begin_resident")
            self.h = h
            return
        keep = [[_u64(t, 4) for t in inst] for inst in seq]
        for inst in keep:
            if len(inst) != 3 or any(t.shape[0] != self.len for t in inst):
                raise SbnError(-2, "sbn_bsumcheck_begin", "a sequential instance needs three tables of 2^|rand| scalars")
        arrs = [(C.c_void_p * max(1, self.S))(*[inst[w].ctypes.data for inst in keep]) for w in range(3)]
        h = C.c_void_p()
        st = ctx.lib.sbn_bsumcheck_begin(ctx.h, handles, C.c_size_t(self.P), C.c_size_t(layer_id), _ptr(rand) if rand.shape[0] else None,
                                         C.c_size_t(rand.shape[0]), arrs[0] if self.S else None, arrs[1] if self.S else None,
                                         arrs[2] if self.S else None, C.c_size_t(self.S), C.byref(h))
        ctx._check(st, "sbn_bsumcheck_begin")
        self.h = h

    def round_eval(self):
        e = np.zeros((self.P + self.S, 3, 4), dtype=np.uint64)
        self.ctx._check(self.ctx.lib.sbn_bsumcheck_round_eval(self.h, _ptr(e)), "sbn_bsumcheck_round_eval")
        return e

    def bind(self, r):
        self.ctx._check(self.ctx.lib.sbn_bsumcheck_bind(self.h, _ptr(_u64(r, 4))), "sbn_bsumcheck_bind")
        self.len //= 2

    def end(self):
        n = self.P + self.S
        a = np.zeros((n, 4), dtype=np.uint64); b = np.zeros((n, 4), dtype=np.uint64)
        c = np.zeros((1 + self.S, 4), dtype=np.uint64)
        self.ctx._check(self.ctx.lib.sbn_bsumcheck_end(self.h, _ptr(a), _ptr(b), _ptr(c)), "sbn_bsumcheck_end")
        return a, b, c

    def prove(self, merlin_state, claim, coeffs, num_rounds):
        """The layer's whole round loop inside the library (sbn_bsumcheck_prove) against the native Merlin state: returns
        (polys uint64[num_rounds, 4, 4], r uint64[num_rounds, 4], final claim, A_final, B_final, C_final), all Montgomery."""
        n = self.P + self.S
        polys = np.zeros((max(1, num_rounds), 4, 4), dtype=np.uint64)
        r = np.zeros((max(1, num_rounds), 4), dtype=np.uint64)
        e = np.zeros(4, dtype=np.uint64)
        a = np.zeros((n, 4), dtype=np.uint64); b = np.zeros((n, 4), dtype=np.uint64)
        c = np.zeros((1 + self.S, 4), dtype=np.uint64)
        st = self.ctx.lib.sbn_bsumcheck_prove(self.h, merlin_state, _ptr(_u64(claim, 4)), _ptr(_u64(coeffs, 4)), C.c_size_t(num_rounds),
                                              _ptr(polys), _ptr(r), _ptr(e), _ptr(a), _ptr(b), _ptr(c))
        self.ctx._check(st, "sbn_bsumcheck_prove")
        self.len = 1
        return polys[:num_rounds], r[:num_rounds], e, a, b, c

    def close(self):
        if self.h and self.ctx.h:
            self.ctx.lib.sbn_bsumcheck_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Addrs:
    """sbn_addrs: the row / column address vectors of a Spark commitment, resident in HBM (uint32[batch, N] each)."""

    def __init__(self, ctx, row_addrs, col_addrs):
        self.ctx = ctx
        row = np.ascontiguousarray(row_addrs, dtype=np.uint32)
        col = np.ascontiguousarray(col_addrs, dtype=np.uint32)
        if row.ndim != 2 or row.shape != col.shape:
            raise SbnError(-2, "sbn_addrs_upload", "row / col address arrays must both be [batch, N]")
        self.batch, self.N = row.shape
        h = C.c_void_p()
        ctx._check(ctx.lib.sbn_addrs_upload(ctx.h, _ptr(row), _ptr(col), C.c_size_t(self.batch), C.c_size_t(self.N), C.byref(h)),
                   "sbn_addrs_upload")
        self.h = h

    def set_timestamps(self, row_read_ts, row_audit_ts, col_read_ts, col_audit_ts):
        arrs = [np.ascontiguousarray(x, dtype=np.uint32) for x in (row_read_ts, row_audit_ts, col_read_ts, col_audit_ts)]
        if arrs[0].shape != (self.batch, self.N) or arrs[2].shape != (self.batch, self.N) or arrs[1].shape != arrs[3].shape:
            raise SbnError(-2, "sbn_addrs_set_timestamps", "read_ts must be [batch, N], audit_ts [num_cells]")
        self.num_cells = arrs[1].shape[0]
        st = self.ctx.lib.sbn_addrs_set_timestamps(self.h, _ptr(arrs[0]), _ptr(arrs[1]), _ptr(arrs[2]), _ptr(arrs[3]),
                                                   C.c_size_t(self.num_cells))
        self.ctx._check(st, "sbn_addrs_set_timestamps")

    def comb_polys(self, val):
        """(comb_ops, comb_mem) as resident polynomials (sparse_mlpoly_full.rs:155-170); val: uint64[batch * N, 4]."""
        val = _u64(val, 4)
        if val.shape[0] != self.batch * self.N:
            raise SbnError(-2, "sbn_spark_comb_polys", "val must hold batch * N scalars")
        po, pm = C.c_void_p(), C.c_void_p()
        self.ctx._check(self.ctx.lib.sbn_spark_comb_polys(self.ctx.h, self.h, _ptr(val), C.byref(po), C.byref(pm)),
                        "sbn_spark_comb_polys")
        out = []
        for h in (po, pm):
            p = Poly.__new__(Poly)
            p.ctx, p.h, p.len = self.ctx, h, int(self.ctx.lib.sbn_poly_len(h))
            out.append(p)
        return out[0], out[1]

    def evaluate(self, comb_ops, rx, ry):
        """multi_evaluate (sparse_mlpoly_full.rs:110-118) on the device: uint64[batch, 4]."""
        rx, ry = _u64(rx, 4), _u64(ry, 4)
        out = np.zeros((self.batch, 4), dtype=np.uint64)
        st = self.ctx.lib.sbn_spark_evaluate(self.ctx.h, self.h, comb_ops.h, _ptr(rx), C.c_size_t(rx.shape[0]), _ptr(ry),
                                             C.c_size_t(ry.shape[0]), _ptr(out))
        self.ctx._check(st, "sbn_spark_evaluate")
        return out

    def hashlayer(self, side, r, r_hash, r_multiset_check):
        """Layers::new for one side: returns [init, read..., write..., audit] as ProdCircuit handles."""
        r = _u64(r, 4)
        n = 2 + 2 * self.batch
        handles = (C.c_void_p * n)()
        st = self.ctx.lib.sbn_hashlayer_build(self.ctx.h, self.h, C.c_int(side), _ptr(r), C.c_size_t(r.shape[0]),
                                              _ptr(_u64(r_hash, 4)), _ptr(_u64(r_multiset_check, 4)), handles)
        self.ctx._check(st, "sbn_hashlayer_build")
        lens = [self.num_cells] + [self.N] * (2 * self.batch) + [self.num_cells]
        return [ProdCircuit._from_handle(self.ctx, C.c_void_p(handles[i]), lens[i]) for i in range(n)]

    def derefs_rows(self):
        """Rows of the Hyrax matrix of the derefs polynomial (2^(ell/2))."""
        used = 2 * self.batch * self.N
        return 1 << (max(0, (used - 1).bit_length()) // 2)

    def derefs_commit(self, bases, rx, ry, keep=True, rows=None):
        """Returns (C, inf, Poly or None): the Hyrax commitment of the derefs polynomial built on the device; rows =
        (row0, nrows) commits only that block of rows (multi-GPU sharding)."""
        rx, ry = _u64(rx, 4), _u64(ry, 4)
        row0, L = rows if rows is not None else (0, self.derefs_rows())
        out = np.zeros((L, 8), dtype=np.uint64)
        inf = np.zeros(L, dtype=np.uint8)
        ph = C.c_void_p()
        st = self.ctx.lib.sbn_derefs_commit_rows(self.ctx.h, bases.h, self.h, _ptr(rx), C.c_size_t(rx.shape[0]), _ptr(ry),
                                                 C.c_size_t(ry.shape[0]), C.c_size_t(row0), C.c_size_t(L), _ptr(out), _ptr(inf),
                                                 C.byref(ph) if keep else None)
        self.ctx._check(st, "sbn_derefs_commit_rows")
        if os.environ.get("SBN_DEBUG_TIMING"):
            print("derefs_commit profile", self.ctx.last_commit_profile(), flush=True)
        poly = None
        if keep:
            poly = Poly.__new__(Poly)
            poly.ctx, poly.h, poly.len = self.ctx, ph, int(self.ctx.lib.sbn_poly_len(ph))
        return out, inf, poly

    def close(self):
        if self.h and self.ctx.h:
            self.ctx.lib.sbn_addrs_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SpMat:
    """sbn_spmat: a sparse matrix in compressed-row form resident on the device (entries: (row, col, Montgomery value))."""

    def __init__(self, ctx, n, ncols, rows, cols, vals):
        self.ctx = ctx
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        order = np.argsort(rows, kind="stable")
        ptr = np.zeros(n + 1, dtype=np.uint32)
        ptr[1:] = np.cumsum(np.bincount(rows, minlength=n)).astype(np.uint32)
        idx = np.ascontiguousarray(np.asarray(cols)[order], dtype=np.uint32)
        val = np.ascontiguousarray(_u64(vals, 4)[order])
        self.n, self.ncols = n, ncols
        h = C.c_void_p()
        ctx._check(ctx.lib.sbn_spmat_upload(ctx.h, _ptr(ptr), _ptr(idx), _ptr(val), C.c_size_t(n), C.c_size_t(len(rows)),
                                            C.c_size_t(ncols), C.byref(h)), "sbn_spmat_upload")
        self.h = h

    @staticmethod
    def mulvec(mats, vec, coeffs=None):
        ctx = mats[0].ctx
        vec = _u64(vec, 4)
        out = np.zeros((mats[0].n, 4), dtype=np.uint64)
        hs = (C.c_void_p * len(mats))(*[m.h for m in mats])
        cf = None if coeffs is None else _u64(coeffs, 4)
        ctx._check(ctx.lib.sbn_spmat_mulvec(ctx.h, hs, _ptr(cf), C.c_size_t(len(mats)), _ptr(vec), C.c_size_t(vec.shape[0]),
                                            _ptr(out)), "sbn_spmat_mulvec")
        return out

    def close(self):
        if self.h and self.ctx.h:
            self.ctx.lib.sbn_spmat_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def hyrax_commit_multi(ctxs, bases, Z, L_size, R_size, blinds=None):
    """sbn_hyrax_commit_multi: one commit over len(ctxs) devices of this process (contiguous row blocks, no exchange)."""
    Z = _u64(Z, 4)
    k = len(ctxs)
    if Z.shape[0] != L_size * R_size:
        raise AssertionError("assert_eq!(L_size * R_size, self.Z.len())")
    bl = _u64(blinds, 4) if blinds is not None else None
    out = np.zeros((L_size, 8), dtype=np.uint64)
    inf = np.zeros(L_size, dtype=np.uint8)
    ca = (C.c_void_p * k)(*[c.h for c in ctxs])
    ba = (C.c_void_p * k)(*[b.h for b in bases])
    st = ctxs[0].lib.sbn_hyrax_commit_multi(ca, ba, C.c_size_t(k), _ptr(Z), C.c_size_t(L_size), C.c_size_t(R_size),
                                            _ptr(bl) if bl is not None else None, _ptr(out), _ptr(inf))
    ctxs[0]._check(st, "sbn_hyrax_commit_multi")
    return out, inf


def eq_evals(ctx, r):
    """EqPolynomial::evals on the device: uint64[2^len(r), 4]."""
    r = np.zeros((0, 4), dtype=np.uint64) if len(r) == 0 else _u64(r, 4)
    out = np.zeros((1 << r.shape[0], 4), dtype=np.uint64)
    ctx._check(ctx.lib.sbn_eq_evals(ctx.h, _ptr(r) if r.shape[0] else None, C.c_size_t(r.shape[0]), _ptr(out)), "sbn_eq_evals")
    return out
