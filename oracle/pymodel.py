"""TEST INFRASTRUCTURE ONLY -- independent Python big-int model of the hot path.

This file is the second, independent oracle (the first is oracle/bn254_oracle.c).
It restates, with Python integers and hashlib only, the reference behaviour of
Antiparadox/Spartan-BN254 for the Hyrax commit/open hot path.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg may import it; the
product (spartan_bn254_b200/) never does.

PARITY STATUS: *unpinned by reference fixtures* -- the reference holds no golden
vectors for this path and cannot be compiled here (no Rust
toolchain, arkworks not vendored).  Parity is pinned mathematically instead: an
MSM result is a unique group element, so canonical affine (x, y, inf) is
algorithm independent.  This model is cross-checked against the C oracle and
against public BN254 constants (2G from EIP-196 test data, group order).

Reference citations (file:line under /root/reference):
  group.rs:110-132     GroupElement::from_uniform_bytes   -> from_uniform_bytes()
  group.rs:135-140     GroupElement::compress             -> compress()
  group.rs:143-175     vartime_multiscalar_mul/msm_affine -> msm()
  commitments.rs:31-62 MultiCommitGens::new               -> multi_commit_gens()
  commitments.rs:144   <[Scalar] as Commitments>::commit  -> commit_row()
  nizk/mod.rs:412-415  DotProductProofGens::new           -> dotproduct_gens()
  hyrax.rs:253-308     DensePolynomial::commit(_inner)    -> hyrax_commit()
  hyrax.rs:311-324     DensePolynomial::bound             -> bound()
  hyrax.rs:355-383     EqPolynomial                       -> eq_evals(), factored_lens()
  nizk/bullet.rs:24    BulletReductionProof::prove        -> bullet_prove()
  scalar.rs:75-95      Scalar::to_bytes/from_bytes
  transcript.rs:56-67  challenge_scalar (64 B LE mod r)
"""
import hashlib

P = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47  # Fq
R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001  # Fr
B = 3
G = (1, 2)
INF = None  # affine identity

MONT_R = 1 << 256


# ------------------------------------------------------------------ curve
def is_on_curve(pt):
    if pt is INF:
        return True
    x, y = pt
    return (y * y - x * x * x - B) % P == 0


def neg(pt):
    if pt is INF:
        return INF
    return (pt[0], (-pt[1]) % P)


def add(p1, p2):
    if p1 is INF:
        return p2
    if p2 is INF:
        return p1
    x1, y1 = p1
    x2, y2 = p2
    if x1 == x2:
        if (y1 + y2) % P == 0:
            return INF
        lam = 3 * x1 * x1 * pow(2 * y1, -1, P) % P
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, P) % P
    x3 = (lam * lam - x1 - x2) % P
    y3 = (lam * (x1 - x3) - y1) % P
    return (x3, y3)


# Jacobian internals for speed (pure-Python affine adds cost one pow() each).
def _jdbl(X, Y, Z):
    if Z == 0 or Y == 0:
        return (1, 1, 0)
    A = X * X % P
    Bq = Y * Y % P
    C = Bq * Bq % P
    D = 2 * ((X + Bq) * (X + Bq) - A - C) % P
    E = 3 * A % P
    F = E * E % P
    X3 = (F - 2 * D) % P
    Y3 = (E * (D - X3) - 8 * C) % P
    Z3 = 2 * Y * Z % P
    return (X3, Y3, Z3)


def _jadd_affine(X1, Y1, Z1, x2, y2):
    if Z1 == 0:
        return (x2, y2, 1)
    Z1Z1 = Z1 * Z1 % P
    U2 = x2 * Z1Z1 % P
    S2 = y2 * Z1 * Z1Z1 % P
    H = (U2 - X1) % P
    r = (S2 - Y1) % P
    if H == 0:
        if r == 0:
            return _jdbl(X1, Y1, Z1)
        return (1, 1, 0)
    HH = H * H % P
    HHH = H * HH % P
    V = X1 * HH % P
    X3 = (r * r - HHH - 2 * V) % P
    Y3 = (r * (V - X3) - Y1 * HHH) % P
    Z3 = Z1 * H % P
    return (X3, Y3, Z3)


def _jaffine(X, Y, Z):
    if Z == 0:
        return INF
    zi = pow(Z, -1, P)
    zi2 = zi * zi % P
    return (X * zi2 % P, Y * zi2 * zi % P)


def mul(k, pt):
    """k * pt, double-and-add (k reduced mod r)."""
    k %= R
    if pt is INF or k == 0:
        return INF
    acc = (1, 1, 0)
    for bit in bin(k)[2:]:
        acc = _jdbl(*acc)
        if bit == "1":
            acc = _jadd_affine(*acc, pt[0], pt[1])
    return _jaffine(*acc)


def msm(scalars, points):
    """Sum_i scalars[i] * points[i]  (group.rs:171-175).  Length mismatch -> identity
    (mirrors `.unwrap_or_default()` swallowing the arkworks error)."""
    if len(scalars) != len(points):
        return INF
    acc = (1, 1, 0)
    for s, pt in zip(scalars, points):
        q = mul(s, pt)
        if q is not INF:
            acc = _jadd_affine(*acc, q[0], q[1])
    return _jaffine(*acc)


def compress(pt):
    """ark-serialize compressed SW point: x LE 32 B, bit7 of byte31 = y > p-y, bit6 = inf
    (group.rs:135-140)."""
    if pt is INF:
        b = bytearray(32)
        b[31] |= 0x40
        return bytes(b)
    x, y = pt
    b = bytearray(x.to_bytes(32, "little"))
    if y > (P - y) % P:
        b[31] |= 0x80
    return bytes(b)


# ------------------------------------------------------------------ generators
def scalar_from_bytes(b32):
    """Scalar::from_bytes (scalar.rs:87-95): LE, None if >= r."""
    v = int.from_bytes(b32, "little")
    return v if v < R else None


def uniform_bytes_to_scalar(chunk64):
    """Returns (scalar, kind) per group.rs:110-132; kind in primary/fallback/one."""
    h = hashlib.sha3_256(chunk64).digest()
    s = scalar_from_bytes(h)
    if s is not None:
        return s, "primary"
    h2 = hashlib.sha3_256(b"fallback" + chunk64).digest()
    s = scalar_from_bytes(h2)
    if s is not None:
        return s, "fallback"
    return 1, "one"


def gen_scalars(n, label):
    """Discrete logs (w.r.t. G) of the n+1 points of MultiCommitGens::new(n,label)
    (commitments.rs:31-62): SHAKE256(label || compress(G)) -> (n+1) x 64 B."""
    xof = hashlib.shake_256(label + compress(G)).digest(64 * (n + 1))
    return [uniform_bytes_to_scalar(xof[64 * i: 64 * i + 64]) for i in range(n + 1)]


def multi_commit_gens(n, label):
    """-> (G_list[n], h) affine."""
    sc = gen_scalars(n, label)
    pts = [mul(s, G) for s, _ in sc]
    return pts[:n], pts[n]


def dotproduct_gens(n, label):
    """DotProductProofGens::new (nizk/mod.rs:412-415): MultiCommitGens::new(n+1).split_at(n)
    -> gens_n = (G[0..n], h), gens_1 = ([G[n]], h)."""
    Gs, h = multi_commit_gens(n + 1, label)
    return (Gs[:n], h), ([Gs[n]], h)


# ------------------------------------------------------------------ Hyrax
def factored_lens(ell):
    return ell // 2, ell - ell // 2


def commit_row(row, blind, gens_n):
    Gs, h = gens_n
    assert len(Gs) == len(row)
    return msm(list(row) + [blind], list(Gs) + [h])


def hyrax_commit(Z, blinds, gens_n):
    """DensePolynomial::commit_inner (hyrax.rs:253-281)."""
    L = len(blinds)
    Rs = len(Z) // L
    assert L * Rs == len(Z)
    return [commit_row(Z[Rs * i: Rs * (i + 1)], blinds[i], gens_n) for i in range(L)]


def bound(Z, Lvec, ell):
    """DensePolynomial::bound (hyrax.rs:311-324)."""
    l, r = factored_lens(ell)
    Ls, Rs = 1 << l, 1 << r
    return [sum(Lvec[j] * Z[j * Rs + i] for j in range(Ls)) % R for i in range(Rs)]


def eq_evals(r):
    """EqPolynomial::evals (hyrax.rs:355-369)."""
    ell = len(r)
    ev = [1] * (1 << ell)
    size = 1
    for j in range(ell):
        size *= 2
        for i in range(size - 1, -1, -2):
            s = ev[i // 2]
            ev[i] = s * r[j] % R
            ev[i - 1] = (s - ev[i]) % R
    return ev


def factored_evals(r):
    l, _ = factored_lens(len(r))
    return eq_evals(r[:l]), eq_evals(r[l:])


# ------------------------------------------------------------------ bullet reduction
def bullet_prove(Q, G_vec, H, a_vec, b_vec, blind, blinds_vec, challenges):
    """BulletReductionProof::prove (nizk/bullet.rs:24-126) with the Fiat-Shamir
    challenges u_i supplied by the caller (the transcript is outside the GPU path)."""
    n = len(G_vec)
    assert len(a_vec) == n and len(b_vec) == n and n & (n - 1) == 0
    Gs, a, b = list(G_vec), list(a_vec), list(b_vec)
    dot = lambda x, y: sum(p * q for p, q in zip(x, y)) % R
    Gamma = add(add(msm(a, Gs), mul(dot(a, b), Q)), mul(blind, H))
    blind_Gamma = blind
    Ls, Rs = [], []
    i = 0
    while n > 1:
        n //= 2
        aL, aR, bL, bR, GL, GR = a[:n], a[n:], b[:n], b[n:], Gs[:n], Gs[n:]
        cL, cR = dot(aL, bR), dot(aR, bL)
        bl, br = blinds_vec[i]
        Lp = add(add(msm(aL, GR), mul(cL, Q)), mul(bl, H))
        Rp = add(add(msm(aR, GL), mul(cR, Q)), mul(br, H))
        u = challenges[i]
        ui = pow(u, -1, R)
        Gs = [add(mul(ui, gl), mul(u, gr)) for gl, gr in zip(GL, GR)]
        a = [(u * x + ui * y) % R for x, y in zip(aL, aR)]
        b = [(ui * x + u * y) % R for x, y in zip(bL, bR)]
        blind_Gamma = (u * u * bl + blind_Gamma + ui * ui * br) % R
        Ls.append(Lp)
        Rs.append(Rp)
        i += 1
    return Ls, Rs, Gamma, a[0], b[0], Gs[0], blind_Gamma


# ------------------------------------------------------------------ synthetic inputs
class SplitMix64:
    def __init__(self, seed):
        self.s = seed & 0xFFFFFFFFFFFFFFFF

    def next(self):
        self.s = (self.s + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return z ^ (z >> 31)

    def scalar(self):
        v = 0
        for i in range(4):
            v |= self.next() << (64 * i)
        return v % R


def to_mont_limbs(v, mod):
    m = v * MONT_R % mod
    return [(m >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]


def from_mont_limbs(limbs, mod):
    m = sum(int(l) << (64 * i) for i, l in enumerate(limbs))
    return m * pow(MONT_R, -1, mod) % mod
