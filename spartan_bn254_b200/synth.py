"""Seeded synthetic inputs for tests and benchmarks, numpy only.

SplitMix64(seed) emits u64s; a scalar is four consecutive outputs as LE limbs reduced mod r.  The
reduced value is used directly as the MONTGOMERY representation (a bijection of Fr, so the scalars
are still uniform) -- no modular multiplication is needed on the host."""
import numpy as np

R_LIMBS = np.array([0x43E1F593F0000001, 0x2833E84879B97091, 0xB85045B68181585D, 0x30644E72E131A029], dtype=np.uint64)


def splitmix64(seed, n, offset=0):
    idx = np.arange(offset + 1, offset + n + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def _geq_r(v):
    """v: uint64[n,4] -> bool[n], v >= r (lexicographic from the top limb)."""
    ge = np.ones(v.shape[0], dtype=bool)
    decided = np.zeros(v.shape[0], dtype=bool)
    for k in (3, 2, 1, 0):
        gt = v[:, k] > R_LIMBS[k]
        lt = v[:, k] < R_LIMBS[k]
        ge = np.where(~decided & lt, False, ge)
        decided |= gt | lt
    return ge


def _sub_r(v, mask):
    out = v.copy()
    borrow = np.zeros(v.shape[0], dtype=np.uint64)
    with np.errstate(over="ignore"):
        for k in range(4):
            a = v[:, k]
            t = a - R_LIMBS[k]
            b1 = (a < R_LIMBS[k]).astype(np.uint64)
            t2 = t - borrow
            b2 = (t < borrow).astype(np.uint64)
            out[:, k] = np.where(mask, t2, a)
            borrow = b1 | b2
    return out


def reduce_mod_r(v):
    """uint64[n,4] (any 256-bit values) -> values mod r (2^256 / r < 6, so <= 5 subtractions)."""
    v = np.ascontiguousarray(v, dtype=np.uint64).reshape(-1, 4).copy()
    for _ in range(6):
        m = _geq_r(v)
        if not m.any():
            break
        v = _sub_r(v, m)
    return v


def uniform_scalars(seed, n, offset=0):
    """S-uniform: n scalars, Montgomery limbs uint64[n,4]."""
    return reduce_mod_r(splitmix64(seed, 4 * n, 4 * offset).reshape(n, 4))


def small_scalars_canonical(seed, n, bits=21):
    """S-small: canonical values < 2^bits (convert with Context.fr_from_canonical)."""
    out = np.zeros((n, 4), dtype=np.uint64)
    out[:, 0] = splitmix64(seed, n) & np.uint64((1 << bits) - 1)
    return out


def derefs_scalars(ell, seed_table=2, seed_addr=3, table_bits=None):
    """S-derefs (mirrors sparse_mlpoly_full.rs:89-100,295): 6 segments of N = 2^(ell-3) gathers
    T[a] from a table of uniform scalars, the last quarter of every segment = T[0], and the final
    quarter of Z zero."""
    n = 1 << ell
    N = n // 8
    tb = table_bits if table_bits is not None else max(1, min(21, ell - 4))
    T = uniform_scalars(seed_table, 1 << tb)
    addr = (splitmix64(seed_addr, 6 * N) % np.uint64(max(1, (1 << tb) // 2))).astype(np.int64)
    Z = np.zeros((n, 4), dtype=np.uint64)
    for s in range(6):
        seg = T[addr[s * N:(s + 1) * N]].copy()
        seg[(3 * N) // 4:] = T[0]
        Z[s * N:(s + 1) * N] = seg
    return Z


def distinct_generators(ctx, n, seed=5):
    """G-distinct: k_j * G with k_j uniform (seed 5); returns (G uint64[n,8], h uint64[8])."""
    from .hyrax import GroupElement
    k = uniform_scalars(seed, n + 1)
    pts, inf = ctx.scalar_mul_batch(GroupElement.generator().xy, k)
    assert not inf.any()
    return pts[:n].copy(), pts[n].copy()
