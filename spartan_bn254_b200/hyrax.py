"""Host-side mirror of the reference's Hyrax / Pedersen interface for the GPU hot path.

Same names, argument meaning and error behaviour as the Rust reference so the parity tests read like
the reference's own tests; every heavy operation goes through the libsbn254 C ABI (no CPU fallback).

  MultiCommitGens      reference commitments.rs:17-114
  DotProductProofGens  reference nizk/mod.rs:404-415
  PolyCommitmentGens   reference hyrax.rs:20-31
  DensePolynomial      reference hyrax.rs:155-324  (commit, commit_inner, bound)
  GroupElement         reference group.rs:20,98-175 (compress, msm_affine, vartime_multiscalar_mul)
"""
import ctypes as _C
import hashlib

import numpy as np

from .lib import Context, SbnError

P_MOD = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
_RINV_P = pow(1 << 256, -1, P_MOD)
_MASK64 = 0xFFFFFFFFFFFFFFFF


def _limbs(v):
    return [(v >> (64 * i)) & _MASK64 for i in range(4)]


def _int(limbs):
    return sum(int(x) << (64 * i) for i, x in enumerate(limbs))


def log_2(n):
    """math.rs:11-15 (Math::log_2): floor(log2(n)), n > 0."""
    assert n > 0
    return n.bit_length() - 1


def compute_factored_lens(ell):
    """EqPolynomial::compute_factored_lens (hyrax.rs:371-373)."""
    return ell // 2, ell - ell // 2


_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


class GroupElement:
    """Affine G1 point in ABI layout (group.rs:20 wraps G1Projective; the GPU hands back affine)."""

    def __init__(self, xy, inf=0):
        self.xy = np.ascontiguousarray(xy, dtype=np.uint64).reshape(8)
        self.inf = int(inf)

    @staticmethod
    def generator():
        one = (1 << 256) % P_MOD
        two = (2 << 256) % P_MOD
        return GroupElement(np.array(_limbs(one) + _limbs(two), dtype=np.uint64), 0)

    @staticmethod
    def identity():
        return GroupElement(np.zeros(8, dtype=np.uint64), 1)

    def affine_ints(self):
        if self.inf:
            return None
        return (_int(self.xy[:4]) * _RINV_P % P_MOD, _int(self.xy[4:]) * _RINV_P % P_MOD)

    def compress(self):
        """group.rs:135-140: ark compressed SW encoding (x LE, bit7 = y > -y, bit6 = infinity).  Through the library's host code
        (sbn_g1_compress) when it is built -- a proof compresses ~400 single points on its Fiat-Shamir path; compress_py is the
        same in Python integers (and what the tests hold the library's version against)."""
        lib = _host_lib()
        if lib is None:
            return self.compress_py()
        out = _C.create_string_buffer(32)
        inf = _C.c_uint8(1 if self.inf else 0)
        if lib.sbn_g1_compress(_C.c_void_p(self.xy.ctypes.data), _C.byref(inf), _C.c_size_t(1), out) != 0:
            raise RuntimeError("sbn_g1_compress failed")
        return out.raw

    def compress_py(self):
        b = bytearray(32)
        a = self.affine_ints()
        if a is None:
            b[31] |= 0x40
            return bytes(b)
        x, y = a
        b[:] = x.to_bytes(32, "little")
        if y > (P_MOD - y) % P_MOD:
            b[31] |= 0x80
        return bytes(b)

    def __eq__(self, o):
        return self.inf == o.inf and (self.inf == 1 or bool(np.array_equal(self.xy, o.xy)))

    @staticmethod
    def msm_affine(scalars, points, points_inf=None, ctx=None):
        """group.rs:171-175.  A length mismatch yields the identity (`unwrap_or_default`)."""
        ctx = ctx or default_context()
        out, inf = ctx.msm(points, points_inf, scalars)
        return GroupElement(out, inf)

    vartime_multiscalar_mul = msm_affine


class MultiCommitGens:
    """commitments.rs:17-27.  G: uint64[n,8], h: uint64[8] (affine Montgomery)."""

    def __init__(self, G, h, ctx=None):
        self.G = np.ascontiguousarray(G, dtype=np.uint64).reshape(-1, 8)
        self.h = np.ascontiguousarray(h, dtype=np.uint64).reshape(8)
        self.n = self.G.shape[0]
        self.ctx = ctx or default_context()
        self._bases = None

    @staticmethod
    def uniform_scalars(n, label):
        """Discrete logs of the n+1 points: SHAKE256(label || compress(G)) -> 64 B chunks ->
        from_uniform_bytes rule (group.rs:110-132): SHA3-256 -> LE scalar if < r, else
        SHA3-256("fallback" || chunk), else 1."""
        xof = hashlib.shake_256(label + GroupElement.generator().compress()).digest(64 * (n + 1))
        out = np.zeros((n + 1, 4), dtype=np.uint64)
        for i in range(n + 1):
            chunk = xof[64 * i: 64 * i + 64]
            v = int.from_bytes(hashlib.sha3_256(chunk).digest(), "little")
            if v >= R_MOD:
                v = int.from_bytes(hashlib.sha3_256(b"fallback" + chunk).digest(), "little")
                if v >= R_MOD:
                    v = 1
            out[i] = _limbs(v)
        return out

    @staticmethod
    def new(n, label, ctx=None):
        """commitments.rs:31-62."""
        ctx = ctx or default_context()
        canon = MultiCommitGens.uniform_scalars(n, label)
        mont = ctx.fr_from_canonical(canon)
        pts, inf = ctx.scalar_mul_batch(GroupElement.generator().xy, mont)
        assert not inf.any()
        return MultiCommitGens(pts[:n], pts[n], ctx)

    @staticmethod
    def from_generators(G, h, ctx=None):
        """commitments.rs:101-114."""
        return MultiCommitGens(G, h, ctx)

    def split_at(self, mid):
        """commitments.rs:78-98: both halves keep the same h."""
        return MultiCommitGens(self.G[:mid], self.h, self.ctx), MultiCommitGens(self.G[mid:], self.h, self.ctx)

    def scale(self, s):
        """commitments.rs:64-76: G_i <- s * G_i, h unchanged."""
        pts, inf = self.ctx.scale_points(self.G, None, s)
        assert not inf.any()
        return MultiCommitGens(pts, self.h, self.ctx)

    def device_bases(self):
        """Resident copy + window tables, built once and reused by every commit."""
        if self._bases is None:
            self._bases = self.ctx.bases(self.G, self.h)
        return self._bases

    def commit(self, scalars, blind):
        """<[Scalar] as Commitments>::commit (commitments.rs:144-154)."""
        scalars = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
        if scalars.shape[0] != self.n:
            raise AssertionError("assert_eq!(gens_n.n, self.len())")   # commitments.rs:146
        out, inf = self.ctx.commit(self.device_bases(), scalars, blind)
        return GroupElement(out, inf)


class DotProductProofGens:
    """nizk/mod.rs:404-415: MultiCommitGens::new(n + 1, label).split_at(n)."""

    def __init__(self, n, label, ctx=None):
        self.n = n
        self.gens_n, self.gens_1 = MultiCommitGens.new(n + 1, label, ctx).split_at(n)
        self._bases_ext = None

    def device_bases_ext(self):
        """gens_n (G, h) plus gens_1.G[0] resident with window tables: what the opening's bullet reduction uses."""
        if self._bases_ext is None:
            # kept apart from gens_n's own tables: a commit merges equal generators (two thirds of the reference's set),
            # the bullet reduction needs every generator as given
            self._bases_ext = self.gens_n.ctx.bases(self.gens_n.G, self.gens_n.h, g1=self.gens_1.G[0])
        return self._bases_ext


class PolyCommitmentGens:
    """hyrax.rs:20-31."""

    def __init__(self, num_vars, label, ctx=None):
        _, right = compute_factored_lens(num_vars)
        # The generators are a function of (n, label) alone (commitments.rs:31-62), and the Spark prover derives gens_ops,
        # gens_mem and gens_derefs from ONE label (sparse_mlpoly_full.rs:625-627): sets of equal n are the same points, so
        # they share one derivation and one resident copy (tables included) per context.
        ctx = ctx or default_context()
        cache = ctx.__dict__.setdefault("_gens_cache", {})
        key = (1 << right, bytes(label))
        if key not in cache:
            cache[key] = DotProductProofGens(1 << right, label, ctx)
        self.gens = cache[key]


class PolyCommitment:
    """hyrax.rs:38-52: C is the vector of row commitments (affine + infinity flags)."""

    def __init__(self, C, inf):
        self.C = C
        self.inf = inf

    def __len__(self):
        return self.C.shape[0]

    def element(self, i):
        return GroupElement(self.C[i], self.inf[i])

    def compressed(self):
        """What append_to_transcript feeds Merlin (hyrax.rs:44-52): 32 B per share."""
        return [self.element(i).compress() for i in range(len(self))]


class DensePolynomial:
    """hyrax.rs:155-324.  Z: uint64[len,4] Montgomery Fr, len a power of two."""

    def __init__(self, Z):
        self.Z = np.ascontiguousarray(Z, dtype=np.uint64).reshape(-1, 4)
        self.len = self.Z.shape[0]
        self.num_vars = log_2(self.len) if self.len > 0 else 0

    def get_num_vars(self):
        return self.num_vars

    def resident(self, ctx=None):
        """Uploads Z once; later commit_inner / bound calls reuse the device copy."""
        if getattr(self, "_poly", None) is None:
            self._poly = (ctx or default_context()).poly_upload(self.Z)
        return self._poly

    def commit_inner(self, blinds, gens, shard=None):
        """hyrax.rs:253-281: one batched GPU call instead of the rayon row loop.  shard = (rank, world, all_gather): rows are
        independent (:259-265), so with one process per GPU and the polynomial resident on each, this rank commits its
        contiguous block of rows (sbn_poly_commit_rows) and the blocks are all-gathered -- 65 bytes per row."""
        blinds = np.ascontiguousarray(blinds, dtype=np.uint64).reshape(-1, 4)
        L_size = blinds.shape[0]
        R_size = self.len // L_size
        if L_size * R_size != self.len:
            raise AssertionError("assert_eq!(L_size * R_size, self.Z.len())")    # hyrax.rs:258
        if gens.n != R_size:
            raise AssertionError("assert_eq!(gens_n.n, self.len())")             # commitments.rs:146
        zero_blinds = not blinds.any()
        if shard is not None and shard[1] > 1 and getattr(self, "_poly", None) is not None and L_size % shard[1] == 0:
            from .parallel import shard_rows
            rank, world, all_gather = shard
            first, n = shard_rows(L_size, world, rank)
            C, inf = self._poly.commit_rows(gens.device_bases(), first, n, R_size, None if zero_blinds else blinds[first:first + n])
            block = np.concatenate([C.view(np.uint8).reshape(n, 64), inf.reshape(n, 1)], axis=1)
            blocks = all_gather(block)
            allb = np.concatenate(blocks)
            return PolyCommitment(np.ascontiguousarray(allb[:, :64]).view(np.uint64).reshape(L_size, 8), np.ascontiguousarray(allb[:, 64]))
        if getattr(self, "_poly", None) is not None:
            C, inf = self._poly.commit(gens.device_bases(), L_size, R_size, None if zero_blinds else blinds)
        else:
            C, inf = gens.ctx.hyrax_commit(gens.device_bases(), self.Z, L_size, R_size, None if zero_blinds else blinds)
        return PolyCommitment(C, inf)

    def commit(self, gens, blinds=None):
        """hyrax.rs:283-308.  `blinds` stands for the random tape: None = zero blinds
        (random_tape = None), otherwise the L_size scalars the tape would have produced."""
        n = self.len
        ell = self.get_num_vars()
        if n != 1 << ell:
            raise AssertionError("assert_eq!(n, ell.pow2())")                    # hyrax.rs:290
        left, right = compute_factored_lens(ell)
        L_size, R_size = 1 << left, 1 << right
        if blinds is None:
            blinds = np.zeros((L_size, 4), dtype=np.uint64)
        blinds = np.ascontiguousarray(blinds, dtype=np.uint64).reshape(-1, 4)
        if blinds.shape[0] != L_size:
            raise AssertionError("blinds.len() == L_size")
        return self.commit_inner(blinds, gens.gens.gens_n), blinds

    def bound(self, L, ctx=None):
        """hyrax.rs:311-324."""
        left, right = compute_factored_lens(self.get_num_vars())
        if getattr(self, "_poly", None) is not None:
            return self._poly.bound(L, 1 << left, 1 << right)
        ctx = ctx or default_context()
        return ctx.bound(self.Z, L, 1 << left, 1 << right)


# ------------------------------------------------------------------------------------------------
# Fr helpers on the host (Python ints): Montgomery limbs <-> canonical values
# ------------------------------------------------------------------------------------------------
_RINV_R = pow(1 << 256, -1, R_MOD)


def fr_to_int(limbs):
    return _int(limbs) * _RINV_R % R_MOD


def fr_from_int(v):
    return np.array(_limbs((int(v) << 256) % R_MOD), dtype=np.uint64)


_BULK = 128     # vectors at least this long change form on the GPU (one kernel) instead of element by element


def _bulk_ctx():
    from . import lib as _lib
    return _lib._live_contexts[-1] if _lib._live_contexts else None


_hostlib = None


def _host_lib():
    """libsbn254's host-side conversions (sbn_fr_{to,from}_canonical_host) when the library has been built."""
    global _hostlib
    if _hostlib is None:
        try:
            from .lib import load_library
            _hostlib = load_library()
        except Exception:
            _hostlib = False
    return _hostlib or None


def _to_canonical_host(arr):
    arr = np.ascontiguousarray(arr, dtype=np.uint64).reshape(-1, 4)
    lib = _host_lib()
    if lib is None:
        return np.frombuffer(b"".join(fr_to_int(a).to_bytes(32, "little") for a in arr), dtype=np.uint64).reshape(-1, 4)
    canon = np.empty_like(arr)
    lib.sbn_fr_to_canonical_host(arr.ctypes.data_as(_C.c_void_p), _C.c_size_t(arr.shape[0]), canon.ctypes.data_as(_C.c_void_p))
    return canon


_TINY_TO, _TINY_FROM = 8, 16    # up to here Python integers beat a library call (2-6 us against 9-13 us per vector)


def fr_vec_to_ints(arr):
    arr = np.ascontiguousarray(arr, dtype=np.uint64).reshape(-1, 4)
    n = arr.shape[0]
    if n <= _TINY_TO:      # the one- to five-element vectors of the ZK sumcheck rounds and Sigma-protocols
        data = arr.tobytes()
        return [int.from_bytes(data[32 * i: 32 * i + 32], "little") * _RINV_R % R_MOD for i in range(n)]
    ctx = _bulk_ctx() if n >= _BULK else None
    if ctx is not None:
        canon = ctx.fr_to_canonical(arr)
    else:
        lib = _host_lib()
        if lib is None or n == 0:
            return [fr_to_int(a) for a in arr]
        canon = np.empty_like(arr)
        lib.sbn_fr_to_canonical_host(arr.ctypes.data_as(_C.c_void_p), _C.c_size_t(n), canon.ctypes.data_as(_C.c_void_p))
    data = canon.tobytes()
    return [int.from_bytes(data[32 * i: 32 * i + 32], "little") for i in range(n)]


def fr_vec_from_ints(vals):
    n = len(vals)
    if n <= _TINY_FROM:
        data = b"".join(((int(v) << 256) % R_MOD).to_bytes(32, "little") for v in vals)
        return np.frombuffer(data, dtype=np.uint64).reshape(-1, 4).copy()
    ctx = _bulk_ctx() if n >= _BULK else None
    lib = _host_lib() if ctx is None else None
    if (ctx is None and lib is None) or n == 0:
        out = np.zeros((n, 4), dtype=np.uint64)
        for i, v in enumerate(vals):
            out[i] = fr_from_int(v)
        return out
    data = b"".join((int(v) % R_MOD).to_bytes(32, "little") for v in vals)
    canon = np.frombuffer(data, dtype=np.uint64).reshape(-1, 4)
    if ctx is not None:
        return ctx.fr_from_canonical(canon)
    out = np.empty((n, 4), dtype=np.uint64)
    lib.sbn_fr_from_canonical_host(canon.ctypes.data_as(_C.c_void_p), _C.c_size_t(n), out.ctypes.data_as(_C.c_void_p))
    return out


class EqPolynomial:
    """hyrax.rs:340-383 on canonical ints."""

    def __init__(self, r):
        self.r = list(r)

    def evals(self):
        ell = len(self.r)
        ctx = _bulk_ctx() if (1 << ell) >= _BULK else None
        if ctx is not None:         # the doubling runs on the device (sbn_eq_evals); only the canonical values come back
            from .lib import eq_evals as _eq_evals
            return fr_vec_to_ints(_eq_evals(ctx, fr_vec_from_ints(self.r)))
        ev = [1] * (1 << ell)
        size = 1
        for j in range(ell):
            size *= 2
            for i in range(size - 1, -1, -2):
                s = ev[i // 2]
                ev[i] = s * self.r[j] % R_MOD
                ev[i - 1] = (s - ev[i]) % R_MOD
        return ev

    def compute_factored_evals(self):
        left, _ = compute_factored_lens(len(self.r))
        return EqPolynomial(self.r[:left]).evals(), EqPolynomial(self.r[left:]).evals()


class DotProductProofLog:
    """nizk/mod.rs:418-522 (prover).  Every group operation runs on the GPU through the C ABI: the (n+1)-point
    commitment Cx, the bullet reduction with device-resident generators, and the 2-point commitments."""

    def __init__(self, L_vec, R_vec, delta, beta, z1, z2):
        self.L_vec, self.R_vec, self.delta, self.beta, self.z1, self.z2 = L_vec, R_vec, delta, beta, z1, z2

    @staticmethod
    def prove(gens, transcript, random_tape, x_vec, blind_x, a_vec, y, blind_y):
        """x_vec, a_vec: lists of canonical ints, or Montgomery uint64[n, 4] arrays (what `bound` and the device eq tables
        hand over: the vectors then never pass through Python integers)."""
        ctx = gens.gens_n.ctx
        transcript.append_protocol_name(b"dot product proof (log)")
        n = len(x_vec)
        assert len(a_vec) == n and gens.n == n                              # mod.rs:451-452
        lg_n = log_2(n)
        d = random_tape.random_scalar(b"d")
        r_delta = random_tape.random_scalar(b"r_delta")
        r_beta = random_tape.random_scalar(b"r_delta")                      # sic (mod.rs:459)
        v1 = random_tape.random_vector(b"blinds_vec_1", lg_n)
        v2 = random_tape.random_vector(b"blinds_vec_2", lg_n)
        x_m = x_vec if isinstance(x_vec, np.ndarray) else fr_vec_from_ints(x_vec)
        a_m = a_vec if isinstance(a_vec, np.ndarray) else fr_vec_from_ints(a_vec)
        # (n+1)-point MSM, mod.rs:470: a single row, over the opening's tabulated set (the g1 column takes no scalar)
        out, inf = ctx.commit(gens.device_bases_ext(), x_m, fr_from_int(blind_x))
        Cx = GroupElement(out, inf)
        transcript.append_point(b"Cx", Cx.compress())
        Cy = gens.gens_1.commit(fr_vec_from_ints([y]), fr_from_int(blind_y))   # mod.rs:473, over gens_1's resident tables
        transcript.append_point(b"Cy", Cy.compress())
        if isinstance(a_vec, np.ndarray):
            transcript.append_scalars_canonical(b"a", ctx.fr_to_canonical(a_m) if n >= _BULK else _to_canonical_host(a_m))
        else:
            transcript.append_scalars(b"a", a_vec)
        r = transcript.challenge_scalar(b"r")
        # gens_1.scale(r) (mod.rs:481): Q = r * gens_1.G[0] never leaves the device -- the reduction takes r (q_scalar below)
        # and beta = d * Q + r_beta * h is committed as (d r) * gens_1.G[0] + r_beta * h, the same group element
        blind_Gamma = (blind_x + r * blind_y) % R_MOD
        # BulletReductionProof::prove (bullet.rs:24-126): G, a, b stay on the device; L, R and u cross the boundary
        # Q = r * gens_1.G[0]: handed over as the scalar r so every MSM of the reduction runs on the window tables
        st = ctx.bullet_begin(gens.device_bases_ext(), None, x_m, a_m, fr_from_int(blind_Gamma),
                              q_scalar=fr_from_int(r))
        L_vec, R_vec = [], []
        rhat = blind_Gamma
        for i in range(lg_n):
            (L, Li), (R, Ri) = st.round(fr_from_int(v1[i]), fr_from_int(v2[i]))
            L, R = GroupElement(L, Li), GroupElement(R, Ri)
            transcript.append_point(b"L", L.compress())
            transcript.append_point(b"R", R.compress())
            u = transcript.challenge_scalar(b"u")
            u_inv = pow(u, -1, R_MOD)
            st.fold(fr_from_int(u), fr_from_int(u_inv))
            rhat = (u * u * v1[i] + rhat + u_inv * u_inv * v2[i]) % R_MOD
            L_vec.append(L)
            R_vec.append(R)
        # delta = d * g_hat + r_delta * h (mod.rs:497-500) comes back with g_hat: a second row over the resident tables
        a_hat_m, b_hat_m, g_hat, g_inf, delta_xy, delta_inf = st.end_delta(fr_from_int(d), fr_from_int(r_delta))
        st.close()
        x_hat, a_hat = fr_to_int(a_hat_m), fr_to_int(b_hat_m)
        y_hat = x_hat * a_hat % R_MOD
        assert not g_inf
        delta = GroupElement(delta_xy, delta_inf)
        transcript.append_point(b"delta", delta.compress())
        beta = gens.gens_1.commit(fr_vec_from_ints([d * r % R_MOD]), fr_from_int(r_beta))   # mod.rs:503
        transcript.append_point(b"beta", beta.compress())
        c = transcript.challenge_scalar(b"c")
        z1 = (d + c * y_hat) % R_MOD
        z2 = (a_hat * (c * rhat + r_beta) + r_delta) % R_MOD
        return DotProductProofLog(L_vec, R_vec, delta, beta, z1, z2), Cx, Cy


class PolyEvalProof:
    """hyrax.rs:55-116 (prover)."""

    def __init__(self, proof):
        self.proof = proof

    @staticmethod
    def prove(poly, blinds, r, Zr, blind_Zr, gens, transcript, random_tape):
        """poly: DensePolynomial; blinds: uint64[L,4] or None; r: list of canonical ints; Zr / blind_Zr: ints."""
        ctx = gens.gens.gens_n.ctx
        transcript.append_protocol_name(b"polynomial evaluation proof")
        assert poly.get_num_vars() == len(r)                                 # hyrax.rs:77
        left, right = compute_factored_lens(len(r))
        L_size, R_size = 1 << left, 1 << right
        # EqPolynomial::compute_factored_evals (hyrax.rs:375-383) on the device; L, R, LZ stay Montgomery arrays
        from .lib import eq_evals as _eq_evals
        L_m = _eq_evals(ctx, fr_vec_from_ints(list(r[:left])))
        R_m = _eq_evals(ctx, fr_vec_from_ints(list(r[left:])))
        LZ = poly.bound(L_m, ctx)                                            # GPU, hyrax.rs:100
        LZ_blind = 0
        if blinds is not None:
            bl = fr_vec_to_ints(blinds)
            assert len(bl) == L_size                                         # hyrax.rs:88
            LZ_blind = sum(b * l for b, l in zip(bl, fr_vec_to_ints(L_m))) % R_MOD
        proof, _C_LR, C_Zr_prime = DotProductProofLog.prove(gens.gens, transcript, random_tape, LZ, LZ_blind, R_m, Zr,
                                                            blind_Zr or 0)
        return PolyEvalProof(proof), C_Zr_prime


__all__ = ["GroupElement", "MultiCommitGens", "DotProductProofGens", "PolyCommitmentGens", "PolyCommitment",
           "DensePolynomial", "compute_factored_lens", "log_2", "SbnError", "EqPolynomial", "DotProductProofLog",
           "PolyEvalProof", "fr_to_int", "fr_from_int", "fr_vec_to_ints", "fr_vec_from_ints"]
