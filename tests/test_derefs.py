"""Device-side derefs: eq tables and the gather are built in HBM from resident address vectors;
the commitment must equal the oracle's commit of the host-built polynomial (sparse_mlpoly_full.rs:245-304, 1713-1724)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("batch,N,nx,ny", [(3, 64, 6, 6), (3, 256, 5, 9), (1, 16, 4, 4), (2, 512, 9, 9)])
def test_derefs_commit_matches_oracle(ctx, orc, batch, N, nx, ny):
    from spartan_bn254_b200 import synth
    from spartan_bn254_b200.lib import Addrs
    from spartan_bn254_b200.hyrax import MultiCommitGens
    rng = np.random.default_rng(11 + N)
    row = rng.integers(0, 1 << nx, size=(batch, N), dtype=np.uint32)
    col = rng.integers(0, 1 << ny, size=(batch, N), dtype=np.uint32)
    row[:, 3 * N // 4:] = 0                      # padding slots all point at address 0 (sparse_mlpoly_full.rs:89-100)
    rx = synth.uniform_scalars(21, nx)
    ry = synth.uniform_scalars(22, ny)
    used = 2 * batch * N
    ell = (used - 1).bit_length()
    L, R = 1 << (ell // 2), 1 << (ell - ell // 2)
    gens = MultiCommitGens.new(R, b"gens_r1cs_eval", ctx)
    addrs = Addrs(ctx, row, col)
    C, inf, poly = addrs.derefs_commit(gens.device_bases(), rx, ry)
    # oracle: eq tables, gather, merge + zero padding, Hyrax commit with zero blinds
    mem_rx, mem_ry = orc.eq_evals(rx), orc.eq_evals(ry)
    Z = np.zeros((1 << ell, 4), dtype=np.uint64)
    Z[: batch * N] = mem_rx[row.reshape(-1)]
    Z[batch * N: used] = mem_ry[col.reshape(-1)]
    assert np.array_equal(poly.download(), Z)
    Co, info = orc.hyrax_commit(gens.G, gens.h, Z, L, R, None)
    assert np.array_equal(inf, info) and np.array_equal(C, Co)
    # the resident polynomial serves the opening's bound (hyrax.rs:311-324) without another upload
    Lv = synth.uniform_scalars(23, L)
    assert np.array_equal(poly.bound(Lv, L, R), orc.bound(Z, Lv, L, R))
    poly.close()
    addrs.close()


def test_derefs_address_out_of_range_is_a_shape_error(ctx):
    from spartan_bn254_b200 import SbnError, synth
    from spartan_bn254_b200.lib import Addrs
    from spartan_bn254_b200.hyrax import MultiCommitGens
    row = np.full((1, 16), 40, dtype=np.uint32)          # 40 >= 2^5
    col = np.zeros((1, 16), dtype=np.uint32)
    gens = MultiCommitGens.new(8, b"gens_r1cs_eval", ctx)
    addrs = Addrs(ctx, row, col)
    with pytest.raises(SbnError) as e:
        addrs.derefs_commit(gens.device_bases(), synth.uniform_scalars(1, 5), synth.uniform_scalars(2, 5))
    assert e.value.status == -2
    addrs.close()


@pytest.mark.parametrize("ell,seg", [(10, None), (12, 9), (16, 13), (1, None)])
def test_resident_polynomial_evaluate(ctx, orc, ell, seg):
    """DensePolynomial::evaluate (hyrax.rs:217-222), whole polynomial and an inner segment (the hash layer evaluates the
    polynomials merged into derefs / comb_ops one by one, sparse_mlpoly_full.rs:907-976)."""
    from spartan_bn254_b200 import SbnError, synth
    Z = synth.uniform_scalars(70 + ell, 1 << ell)
    poly = ctx.poly_upload(Z)
    r = synth.uniform_scalars(71, ell)
    assert np.array_equal(poly.evaluate(r), orc.evaluate(Z, r))
    if seg is not None:
        off = 3 << seg
        rs = synth.uniform_scalars(72, seg)
        assert np.array_equal(poly.evaluate(rs, offset=off), orc.evaluate(Z[off: off + (1 << seg)], rs))
        with pytest.raises(SbnError):
            poly.evaluate(rs, offset=(1 << ell) - 5)
    poly.close()
