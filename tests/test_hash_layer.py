"""Hash layer + product circuits of the memory-checking network built on the GPU (sparse_mlpoly_full.rs:745-841)."""
import random

import numpy as np
import pytest

R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001


def test_vectorised_timestamps_match_the_reference_walk():
    """AddrTimestamps::new walks the operations sequentially (sparse_mlpoly_full.rs:220-236); the host mirror counts with a
    stable sort.  Compared with the oracle's line-by-line restatement, including heavily repeated addresses."""
    import product_model as pm
    from spartan_bn254_b200.sparse_mlpoly import AddrTimestamps
    rnd = random.Random(5)
    for num_cells, n, batch in ((8, 16, 3), (64, 64, 1), (4, 32, 2), (256, 128, 3)):
        ops = [[rnd.randrange(num_cells) if rnd.random() < 0.8 else 0 for _ in range(n)] for _ in range(batch)]
        want_read, want_audit = pm.addr_timestamps(num_cells, ops)
        got = AddrTimestamps(num_cells, np.array(ops, dtype=np.uint32))
        assert got.read_ts.tolist() == want_read and got.audit_ts.tolist() == want_audit


@pytest.mark.gpu
@pytest.mark.parametrize("nr,N,batch", [(4, 16, 1), (6, 64, 3), (8, 64, 3), (5, 256, 2)])
def test_hash_layer_circuits_match_oracle(ctx, orc, nr, N, batch):
    import product_model as pm
    from spartan_bn254_b200 import synth
    from spartan_bn254_b200.hyrax import fr_vec_to_ints, fr_to_int
    from spartan_bn254_b200.sparse_mlpoly import SparkAddresses, PolyEvalNetwork
    M = 1 << nr
    rng = np.random.default_rng(nr * 100 + N)
    row = rng.integers(0, M, size=(batch, N), dtype=np.uint32)
    col = rng.integers(0, M, size=(batch, N), dtype=np.uint32)
    col[:, N // 2:] = 0
    spark = SparkAddresses(ctx, M, row, col)
    rx, ry = synth.uniform_scalars(31, nr), synth.uniform_scalars(32, nr)
    gam = synth.uniform_scalars(33, 2)
    net = PolyEvalNetwork(spark, rx, ry, (gam[0], gam[1]))
    r_hash, r_ms = fr_to_int(gam[0]), fr_to_int(gam[1])
    for side, layers, addr, r in ((0, net.row_layers, row, rx), (1, net.col_layers, col, ry)):
        mem = pm.eq_evals(fr_vec_to_ints(r))
        read_ts, audit_ts = pm.addr_timestamps(M, addr.tolist())
        derefs = [[mem[a] for a in inst] for inst in addr.tolist()]
        init, reads, writes, audit = pm.build_hash_layer(mem, addr.tolist(), derefs, read_ts, audit_ts, r_hash, r_ms)
        pl = layers.prod_layer
        assert fr_vec_to_ints(pl.init.gpu.layer(0)) == init
        assert fr_vec_to_ints(pl.audit.gpu.layer(0)) == audit
        for k in range(batch):
            assert fr_vec_to_ints(pl.read_vec[k].gpu.layer(0)) == reads[k]
            assert fr_vec_to_ints(pl.write_vec[k].gpu.layer(0)) == writes[k]
        for got, poly in zip(pl.all(), [init] + reads + writes + [audit]):
            assert got.evaluate() == pm.ProductCircuit(poly).evaluate()
        # the multiset identity the reference debug-asserts (:823-828): init * writes == reads * audit
        lhs, rhs = pl.init.evaluate(), pl.audit.evaluate()
        for k in range(batch):
            lhs = lhs * pl.write_vec[k].evaluate() % R
            rhs = rhs * pl.read_vec[k].evaluate() % R
        assert lhs == rhs
        for c in pl.all():
            c.close()
    spark.close()


@pytest.mark.gpu
def test_hash_layer_full_size_multiset_identity(ctx):
    """2^18 cells, 3 x 2^20 operations per side: the memory-checking identity init * prod(writes) == prod(reads) * audit
    holds for the circuits built on the device (a size-independent property of correct hashing + timestamps)."""
    from spartan_bn254_b200 import synth
    from spartan_bn254_b200.sparse_mlpoly import SparkAddresses, Layers
    nr, N, batch = 18, 1 << 20, 3
    rng = np.random.default_rng(9)
    row = rng.integers(0, 1 << nr, size=(batch, N), dtype=np.uint32)
    row[:, 3 * N // 4:] = 0
    spark = SparkAddresses(ctx, 1 << nr, row, row[::-1].copy())
    gam = synth.uniform_scalars(34, 2)
    pl = Layers(spark, 0, synth.uniform_scalars(35, nr), (gam[0], gam[1])).prod_layer
    lhs, rhs = pl.init.evaluate(), pl.audit.evaluate()
    for k in range(batch):
        lhs = lhs * pl.write_vec[k].evaluate() % R
        rhs = rhs * pl.read_vec[k].evaluate() % R
    assert lhs == rhs and lhs != 0
    for c in pl.all():
        c.close()
    spark.close()
