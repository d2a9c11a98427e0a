"""GPU parity tests (run on a B200 with -m gpu): the CUDA Hyrax commit, called through the C ABI, must be
bit-exact (affine x, y, infinity flag) against the CPU oracle and the committed golden vectors."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "hyrax_golden.json")))


def h2i(s):
    return int(s, 16)


def test_generators_match_reference_rule(ctx, orc):
    """MultiCommitGens::new (commitments.rs:31-62): SHAKE256 / SHA3-256 rule + GPU scalar muls."""
    from spartan_bn254_b200.hyrax import MultiCommitGens, DotProductProofGens
    for label in (b"gens_r1cs_sat", b"gens_r1cs_eval", b"test"):
        gens = MultiCommitGens.new(40, label, ctx)
        G, h = orc.multi_commit_gens(label, 40)
        assert np.array_equal(gens.G, G) and np.array_equal(gens.h, h)
    g = GOLD["generators"]["gens_r1cs_eval"]
    gens = MultiCommitGens.new(16, b"gens_r1cs_eval", ctx)
    got = orc.points_to_ints(gens.G[:6], [0] * 6)
    assert got == [(h2i(p[0]), h2i(p[1])) for p in g["points"]]
    # DotProductProofGens::new(n) = MultiCommitGens::new(n+1).split_at(n)  (nizk/mod.rs:412-415)
    d = DotProductProofGens(8, b"gens_r1cs_eval", ctx)
    Gn, hh, G1 = orc.dotproduct_gens(b"gens_r1cs_eval", 8)
    assert np.array_equal(d.gens_n.G, Gn) and np.array_equal(d.gens_n.h, hh)
    assert np.array_equal(d.gens_1.G[0], G1) and np.array_equal(d.gens_1.h, hh)


def test_commit_golden_vector(ctx, orc):
    from spartan_bn254_b200.hyrax import DensePolynomial, PolyCommitmentGens
    g = GOLD["hyrax_commit_4x8"]
    gens = PolyCommitmentGens(5, g["label"].encode(), ctx)
    Z = orc.to_mont([h2i(z) for z in g["Z"]])
    bl = orc.to_mont([h2i(b) for b in g["blinds"]])
    comm, _ = DensePolynomial(Z).commit(gens, bl)
    exp = [None if p is None else (h2i(p[0]), h2i(p[1])) for p in g["C"]]
    assert orc.points_to_ints(comm.C, comm.inf) == exp
    assert [c.hex() for c in comm.compressed()] == g["C_compressed"]
    assert comm.inf[3] == 1


SHAPES = [
    # (ell, gens kind, blinds, scalars)
    (2, "ref", True, "uniform"),
    (7, "ref", False, "uniform"),
    (12, "ref", True, "uniform"),        # cfg0 witness: 64 x 64, random blinds (r1csproof.rs:210-237)
    (15, "ref", False, "derefs"),        # cfg0 derefs: 128 x 256, zero blinds, last quarter of rows zero
    (16, "ref", False, "derefs"),        # 256 x 256
    (13, "distinct", True, "uniform"),
    (14, "ref", False, "small"),         # comb_ops-like small scalars
]


@pytest.mark.parametrize("ell,gens_kind,use_blinds,kind", SHAPES)
def test_commit_matches_oracle(ctx, orc, ell, gens_kind, use_blinds, kind):
    from spartan_bn254_b200 import synth
    from spartan_bn254_b200.hyrax import DensePolynomial, PolyCommitmentGens, MultiCommitGens, compute_factored_lens
    l, r = compute_factored_lens(ell)
    L, R = 1 << l, 1 << r
    if gens_kind == "ref":
        gens = PolyCommitmentGens(ell, b"gens_r1cs_eval", ctx)
    else:
        G, h = synth.distinct_generators(ctx, R)
        gens = PolyCommitmentGens.__new__(PolyCommitmentGens)
        gens.gens = type("D", (), {})()
        gens.gens.gens_n = MultiCommitGens(G, h, ctx)
    if kind == "uniform":
        Z = synth.uniform_scalars(1, L * R)
    elif kind == "derefs":
        Z = synth.derefs_scalars(ell)
    else:
        Z = ctx.fr_from_canonical(synth.small_scalars_canonical(8, L * R))
    blinds = synth.uniform_scalars(4, L) if use_blinds else None
    comm, _ = DensePolynomial(Z).commit(gens, blinds)
    gn = gens.gens.gens_n
    C, inf = orc.hyrax_commit(gn.G, gn.h, Z, L, R, blinds)
    assert np.array_equal(comm.inf, inf)
    assert np.array_equal(comm.C, C)
    if kind == "derefs":
        assert comm.inf[3 * L // 4:].all() and not comm.inf[: 3 * L // 4].any()


def test_commit_adversarial_rows(ctx, orc):
    """Inputs the reference never tests: all-zero rows, scalar 0 / 1 / r-1, rows that
    cancel to the identity through duplicate generators, a blind that cancels the row."""
    from spartan_bn254_b200 import synth
    from spartan_bn254_b200.hyrax import MultiCommitGens
    R = 32
    gens = MultiCommitGens.new(R, b"gens_r1cs_eval", ctx)      # contains many copies of G
    sc, kinds = orc.gen_scalars(b"gens_r1cs_eval", R)
    ones = [i for i in range(R) if kinds[i] == 2]
    assert len(ones) >= 4
    rmod = h2i(GOLD["constants"]["r"])
    L = 8
    Z = [[0] * R for _ in range(L)]
    Z[1][0] = 1
    Z[2][ones[0]] = 5; Z[2][ones[1]] = rmod - 5            # 5G - 5G = identity
    Z[3] = [rmod - 1] * R
    Z[4][ones[0]] = 7; Z[4][ones[1]] = 7; Z[4][ones[2]] = 7  # same base three times (P + P paths)
    Z[5] = [1] * R
    Z[6][3] = 123456789
    Zm = orc.to_mont([v for row in Z for v in row])
    blinds = [0] * L
    blinds[6] = 42
    bm = orc.to_mont(blinds)
    C, inf = ctx.hyrax_commit(gens.device_bases(), Zm, L, R, bm)
    Co, info = orc.hyrax_commit(gens.G, gens.h, Zm, L, R, bm)
    assert np.array_equal(inf, info) and np.array_equal(C, Co)
    assert inf[0] == 1 and inf[2] == 1


def test_commit_chunking_and_window_override(ctx, orc):
    """Results do not depend on the pipeline chunk size or the window width."""
    from spartan_bn254_b200 import Context, synth
    L, R = 64, 128
    G, h = synth.distinct_generators(ctx, R)
    Z = synth.uniform_scalars(2, L * R)
    bl = synth.uniform_scalars(3, L)
    Co, info = orc.hyrax_commit(G, h, Z, L, R, bl)
    for chunk, c in ((7, 0), (64, 5), (1000, 9), (16, 12)):
        c2 = Context(0)
        c2.set("chunk_rows", chunk)
        c2.set("window_bits", c)
        b = c2.bases(G, h)
        if c:
            assert b.window_bits == c
        C, inf = c2.hyrax_commit(b, Z, L, R, bl)
        assert np.array_equal(C, Co) and np.array_equal(inf, info), (chunk, c)
        b.close()
        c2.close()


@pytest.mark.parametrize("rounds", [1, 2, 3])
@pytest.mark.parametrize("gens_kind,kind,ell", [("ref", "uniform", 12), ("ref", "derefs", 15), ("distinct", "uniform", 13),
                                               ("ref", "small", 14)])
def test_batched_affine_rounds_match_oracle(orc, rounds, gens_kind, kind, ell):
    """The batched-affine pre-reduction (ba_kernels.cuh) is normally chosen only for chunks of > 8 M list entries; forced
    here on small shapes.  The reference generators repeat the same point (group.rs:110-132), so the P + P, P + (-P) and
    identity cases of the affine rounds are all exercised; results must stay bit-exact."""
    from spartan_bn254_b200 import Context, synth
    from spartan_bn254_b200.hyrax import MultiCommitGens, compute_factored_lens
    c2 = Context(0)
    c2.set("ba_rounds", rounds)
    l, r = compute_factored_lens(ell)
    L, R = 1 << l, 1 << r
    if gens_kind == "ref":
        g = MultiCommitGens.new(R, b"gens_r1cs_eval", c2)
        G, h = g.G, g.h
    else:
        G, h = synth.distinct_generators(c2, R)
    if kind == "uniform":
        Z = synth.uniform_scalars(5, L * R)
    elif kind == "derefs":
        Z = synth.derefs_scalars(ell)
    else:
        Z = c2.fr_from_canonical(synth.small_scalars_canonical(8, L * R))
    blinds = synth.uniform_scalars(6, L)
    bases = c2.bases(G, h)
    C, inf = c2.hyrax_commit(bases, Z, L, R, blinds)
    Co, info = orc.hyrax_commit(G, h, Z, L, R, blinds)
    assert np.array_equal(inf, info) and np.array_equal(C, Co)
    # an adversarial row set on the same context: cancellations and repeated bases
    if gens_kind == "ref":
        rmod = h2i(GOLD["constants"]["r"])
        sc, kinds = orc.gen_scalars(b"gens_r1cs_eval", R)
        ones = [i for i in range(R) if kinds[i] == 2]
        Zr = [[0] * R for _ in range(4)]
        Zr[0][ones[0]] = 5; Zr[0][ones[1]] = rmod - 5
        Zr[1] = [7] * R
        Zr[2][ones[0]] = 9; Zr[2][ones[1]] = 9; Zr[2][ones[2]] = 9; Zr[2][ones[3]] = rmod - 9
        Zm = orc.to_mont([v for row in Zr for v in row])
        C, inf = c2.hyrax_commit(bases, Zm, 4, R, None)
        Co, info = orc.hyrax_commit(G, h, Zm, 4, R, None)
        assert np.array_equal(inf, info) and np.array_equal(C, Co)
        assert inf[0] == 1 and inf[3] == 1
    bases.close()
    c2.close()


def test_shape_errors(ctx):
    """Reference preconditions (commitments.rs:146, hyrax.rs:258) surface as SBN_ERR_SHAPE / AssertionError."""
    from spartan_bn254_b200 import SbnError, synth
    from spartan_bn254_b200.hyrax import DensePolynomial, MultiCommitGens
    G, h = synth.distinct_generators(ctx, 16)
    gens = MultiCommitGens(G, h, ctx)
    Z = synth.uniform_scalars(1, 64)
    with pytest.raises(SbnError) as e:
        ctx.hyrax_commit(gens.device_bases(), Z, 8, 8, None)       # R_size != gens.n
    assert e.value.status == -2
    with pytest.raises(AssertionError):
        DensePolynomial(Z).commit_inner(np.zeros((8, 4), dtype=np.uint64), gens)


def test_full_size_linearity_property(ctx):
    """cfg1 size (1024 x 1024): commit(Z1 + Z2) == commit(Z1) + commit(Z2) row by row, checked with the
    GPU's own point addition through a 2-term MSM, plus the oracle on a sample of rows."""
    import oracle as orc
    from spartan_bn254_b200 import synth
    L = R = 1024
    G, h = synth.distinct_generators(ctx, R)
    bases = ctx.bases(G, h)
    Z1 = synth.uniform_scalars(21, L * R)
    Z2 = synth.uniform_scalars(22, L * R)
    C1, i1 = ctx.hyrax_commit(bases, Z1, L, R, None)
    C2, i2 = ctx.hyrax_commit(bases, Z2, L, R, None)
    assert not i1.any() and not i2.any()
    rows = np.arange(0, L, 64)
    Zs = Z1.reshape(L, R, 4)[rows].reshape(-1, 4)
    Co, info = orc.hyrax_commit(G, h, Zs, len(rows), R, None, threads=0)
    assert np.array_equal(C1[rows], Co)
    # linearity through the oracle's point addition on the sampled rows
    Zsum = synth.reduce_mod_r(_add256(Z1.reshape(L, R, 4)[rows].reshape(-1, 4), Z2.reshape(L, R, 4)[rows].reshape(-1, 4)))
    Csum, isum = ctx.hyrax_commit(bases, np.tile(Zsum, (L // len(rows), 1)), L, R, None)
    for k, rrow in enumerate(rows):
        o = np.zeros(8, dtype=np.uint64)
        oi = np.zeros(1, dtype=np.uint8)
        orc.lib().orc_g1_add_affine(C1[rrow].ctypes.data, 0, C2[rrow].ctypes.data, 0, o.ctypes.data, oi.ctypes.data)
        assert np.array_equal(o, Csum[k]) and oi[0] == 0


def _add256(a, b):
    """limb-wise 256-bit add of uint64[n,4] arrays, both < r < 2^254 (no overflow out of limb 3)."""
    out = np.zeros_like(a)
    carry = np.zeros(a.shape[0], dtype=np.uint64)
    with np.errstate(over="ignore"):
        for k in range(4):
            t = a[:, k] + b[:, k]
            c1 = (t < a[:, k]).astype(np.uint64)
            t2 = t + carry
            c2 = (t2 < t).astype(np.uint64)
            out[:, k] = t2
            carry = c1 | c2
    return out


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 14, 15])
def test_short_commitments_match_oracle(ctx, orc, n):
    """The Sigma-protocols' short vectors (commitments.rs:118-154 over gens_1 / gens_3 / gens_4, sumcheck.rs:559-649):
    generator sets of at most 16 points take the tabulated one-launch path (small_kernels.cuh).  Checked against the oracle
    and against the general pipeline on random rows and on digit-boundary scalars (0, 1, r - 1, 0x80 / 0x7f / 0x81 bytes,
    which sit on the signed-digit carry), with and without blinds, with the reference's duplicate generators."""
    from spartan_bn254_b200 import Context, synth
    from spartan_bn254_b200.hyrax import DotProductProofGens, MultiCommitGens
    rmod = h2i(GOLD["constants"]["r"])
    gens = MultiCommitGens.new(n, b"gens_r1cs_sat", ctx)
    edge = [0, 1, rmod - 1, rmod - 2, int.from_bytes(b"\x80" * 31, "little"), int.from_bytes(b"\x7f" * 31, "little"),
            int.from_bytes(b"\x81" * 31, "little"), 128, 127, 129, 1 << 253, (1 << 253) + (1 << 8) - 128]
    for L in (1, 2, 41, 64, 65):              # 65 rows fall back to the pipeline
        Z = synth.uniform_scalars(11 + L, L * n)
        ez = orc.to_mont([edge[(i * 5 + 3 * (i // n)) % len(edge)] for i in range(L * n)])
        for Zm, bl in ((Z, synth.uniform_scalars(5, L)), (Z, None), (ez, orc.to_mont([edge[(7 * i) % len(edge)] for i in range(L)]))):
            C, inf = ctx.hyrax_commit(gens.device_bases(), Zm, L, n, bl)
            Co, info = orc.hyrax_commit(gens.G, gens.h, Zm, L, n, bl)
            assert np.array_equal(inf, info) and np.array_equal(C, Co), (n, L)
    # the same rows through the general pipeline, and through a set that carries gens_1's generator as well
    c2 = Context(0)
    c2.set("small_commit_path", 0)
    b2 = c2.bases(gens.G, gens.h)
    Z = synth.uniform_scalars(3, 8 * n)
    bl = synth.uniform_scalars(4, 8)
    C, inf = ctx.hyrax_commit(gens.device_bases(), Z, 8, n, bl)
    C2, inf2 = c2.hyrax_commit(b2, Z, 8, n, bl)
    assert np.array_equal(C, C2) and np.array_equal(inf, inf2)
    b2.close()
    c2.close()
    d = DotProductProofGens(n, b"gens_r1cs_sat", ctx)
    C3, inf3 = ctx.hyrax_commit(d.device_bases_ext(), Z, 8, n, bl)
    Co, info = orc.hyrax_commit(d.gens_n.G, d.gens_n.h, Z, 8, n, bl)
    assert np.array_equal(C3, Co) and np.array_equal(inf3, info)
    # a single scalar that cancels against the blind: x G + (r - x) G when h equals the generator
    one = MultiCommitGens.from_generators(gens.G[:1].reshape(1, 8), gens.G[0], ctx)
    C, inf = ctx.hyrax_commit(one.device_bases(), orc.to_mont([5]), 1, 1, orc.to_mont([rmod - 5]))
    assert inf[0] == 1


@pytest.mark.parametrize("n", [16, 63, 64, 65, 200, 1024])
def test_tabulated_few_row_commits_match_oracle(ctx, orc, n):
    """Single rows and row pairs over an opening's generator set (Cx of nizk/mod.rs:470, the L / R rows of bullet.rs:75-76)
    take the tabulated sum-of-table-points path (small_kernels.cuh, k_tab_commit_*): checked against the oracle and against
    the bucket pipeline for 1..5 rows (5 rows fall back to the pipeline), random and digit-boundary scalars, with and
    without blinds, row lengths n and n + 1 (the extra column is gens_1's generator)."""
    from spartan_bn254_b200 import Context, synth
    from spartan_bn254_b200.hyrax import DotProductProofGens
    rmod = h2i(GOLD["constants"]["r"])
    d = DotProductProofGens(n, b"gens_r1cs_eval", ctx)
    bases = d.device_bases_ext()
    edge = [0, 1, rmod - 1, rmod - 2, int.from_bytes(b"\x80" * 31, "little"), int.from_bytes(b"\x7f" * 31, "little"),
            int.from_bytes(b"\x81" * 31, "little"), 128, 127, 129, 1 << 253, (1 << 253) + (1 << 8) - 128]
    c2 = Context(0)
    c2.set("small_commit_path", 0)
    b2 = c2.bases(d.gens_n.G, d.gens_n.h, g1=d.gens_1.G[0])
    for L in (1, 2, 3, 4, 5):
        Z = synth.uniform_scalars(20 + L, L * n)
        ez = orc.to_mont([edge[(i * 7 + 5 * (i // n)) % len(edge)] for i in range(L * n)])
        for Zm, bl in ((Z, synth.uniform_scalars(6, L)), (Z, None), (ez, orc.to_mont([edge[(3 * i + 1) % len(edge)] for i in range(L)]))):
            C, inf = ctx.hyrax_commit(bases, Zm, L, n, bl)
            Co, info = orc.hyrax_commit(d.gens_n.G, d.gens_n.h, Zm, L, n, bl)
            assert np.array_equal(inf, info) and np.array_equal(C, Co), (n, L)
            C2, inf2 = c2.hyrax_commit(b2, Zm, L, n, bl)
            assert np.array_equal(C, C2) and np.array_equal(inf, inf2), (n, L)
    b2.close()
    c2.close()


@pytest.mark.parametrize("gens_kind,L,R,chunk", [("ref", 4, 32, 0), ("ref", 64, 64, 24), ("distinct", 32, 16, 0), ("distinct", 8, 200, 3),
                                                 ("ref", 16, 1024, 0), ("distinct", 300, 33, 128)])
def test_tabulated_sum_commits_match_oracle(orc, gens_kind, L, R, chunk):
    """Many-row commits as sums of tabulated digit multiples (mult_kernels.cuh), forced on small shapes with
    mult_min_rows = 1: against the oracle and against the bucket pipeline (mult_max_mb = 0), on the reference's generators
    (duplicates merged first) and on distinct ones, with blinds, zero rows, digit-boundary scalars, odd chunking, and through
    both the host-pointer and the resident-polynomial entry points."""
    from spartan_bn254_b200 import Context, synth
    from spartan_bn254_b200.hyrax import MultiCommitGens
    from spartan_bn254_b200.lib import Poly
    rmod = h2i(GOLD["constants"]["r"])
    cm, cp = Context(0), Context(0)
    cm.set("mult_min_rows", 1)
    cp.set("mult_max_mb", 0)
    if chunk:
        cm.set("chunk_rows", chunk)
    if gens_kind == "ref":
        g = MultiCommitGens.new(R, b"gens_r1cs_eval", cm)
        G, h = g.G, g.h
    else:
        G, h = synth.distinct_generators(cm, R)
    bm, bp = cm.bases(G, h), cp.bases(G, h)
    edge = [0, 1, rmod - 1, rmod - 2, int.from_bytes(b"\x80" * 31, "little"), int.from_bytes(b"\x7f" * 31, "little"),
            int.from_bytes(b"\x81" * 31, "little"), 1 << 12, (1 << 12) - 1, (1 << 12) + 1, 1 << 253, (1 << 15) + 1, 1 << 15]
    Z = synth.uniform_scalars(31, L * R)
    Z[R: 2 * R] = 0                                   # an all-zero row -> identity
    Z[2 * R: 3 * R] = orc.to_mont([edge[i % len(edge)] for i in range(R)])
    for bl in (synth.uniform_scalars(32, L), None):
        C, inf = cm.hyrax_commit(bm, Z, L, R, bl)
        Co, info = orc.hyrax_commit(G, h, Z, L, R, bl)
        assert np.array_equal(inf, info) and np.array_equal(C, Co)
        C2, inf2 = cp.hyrax_commit(bp, Z, L, R, bl)
        assert np.array_equal(C, C2) and np.array_equal(inf, inf2)
    assert inf[1] == 1                                # the zero row without a blind
    assert cm.last_commit_profile()["reduce"]["launches"] == 0 and cp.last_commit_profile()["reduce"]["launches"] > 0
    poly = Poly(cm, Z)
    C3, inf3 = poly.commit(bm, L, R, None)
    assert np.array_equal(C3, C) and np.array_equal(inf3, inf)
    poly.close()
    for b in (bm, bp):
        b.close()
    cm.close()
    cp.close()
