"""spartan_bn254_b200 -- B200-native (sm_100a) backend for the Hyrax commit / opening hot path of
Antiparadox/Spartan-BN254.  The product is libsbn254.so (hand-written CUDA behind the C ABI in
include/sbn254.h); this package is the thin Python host mirror used by tests and benchmarks.
There is no CPU fallback: importing `lib` without the built CUDA library raises."""
from .lib import Context, Bases, SbnError, load_library  # noqa: F401
from .hyrax import (  # noqa: F401
    DensePolynomial, MultiCommitGens, PolyCommitmentGens, DotProductProofGens, PolyCommitment,
    GroupElement, compute_factored_lens,
)
