import numpy as np


def test_splitmix_matches_python_model():
    import pymodel as pm
    from spartan_bn254_b200 import synth
    sm = pm.SplitMix64(1)
    ref = [sm.next() for _ in range(16)]
    assert [int(x) for x in synth.splitmix64(1, 16)] == ref
    sm = pm.SplitMix64(9)
    exp = [sm.scalar() for _ in range(50)]
    got = synth.uniform_scalars(9, 50)
    assert [sum(int(got[i, k]) << (64 * k) for k in range(4)) for i in range(50)] == exp


def test_derefs_shape():
    from spartan_bn254_b200 import synth
    Z = synth.derefs_scalars(12)
    assert Z.shape == (4096, 4)
    assert not Z[3072:].any()              # final quarter zero (sparse_mlpoly_full.rs:295 padding)
    assert Z[:3072].any()
    N = 512
    assert np.array_equal(Z[N - 1], Z[N - 2])   # tail of a segment repeats T[0]
