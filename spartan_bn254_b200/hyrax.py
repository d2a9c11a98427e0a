"""Host-side mirror of the reference's Hyrax / Pedersen interface for the GPU hot path.

Same names, argument meaning and error behaviour as the Rust reference so the parity tests read like
the reference's own tests; every heavy operation goes through the libsbn254 C ABI (no CPU fallback).

  MultiCommitGens      reference commitments.rs:17-114
  DotProductProofGens  reference nizk/mod.rs:404-415
  PolyCommitmentGens   reference hyrax.rs:20-31
  DensePolynomial      reference hyrax.rs:155-324  (commit, commit_inner, bound)
  GroupElement         reference group.rs:20,98-175 (compress, msm_affine, vartime_multiscalar_mul)
"""
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
        """group.rs:135-140: ark compressed SW encoding (x LE, bit7 = y > -y, bit6 = infinity)."""
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


class PolyCommitmentGens:
    """hyrax.rs:20-31."""

    def __init__(self, num_vars, label, ctx=None):
        _, right = compute_factored_lens(num_vars)
        self.gens = DotProductProofGens(1 << right, label, ctx)


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

    def commit_inner(self, blinds, gens):
        """hyrax.rs:253-281: one batched GPU call instead of the rayon row loop."""
        blinds = np.ascontiguousarray(blinds, dtype=np.uint64).reshape(-1, 4)
        L_size = blinds.shape[0]
        R_size = self.len // L_size
        if L_size * R_size != self.len:
            raise AssertionError("assert_eq!(L_size * R_size, self.Z.len())")    # hyrax.rs:258
        if gens.n != R_size:
            raise AssertionError("assert_eq!(gens_n.n, self.len())")             # commitments.rs:146
        zero_blinds = not blinds.any()
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
        ctx = ctx or default_context()
        return ctx.bound(self.Z, L, 1 << left, 1 << right)


__all__ = ["GroupElement", "MultiCommitGens", "DotProductProofGens", "PolyCommitmentGens", "PolyCommitment",
           "DensePolynomial", "compute_factored_lens", "log_2", "SbnError"]
