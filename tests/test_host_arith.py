"""CPU tests of the DEVICE arithmetic: fp.cuh / ec.cuh are __host__ __device__ with the PTX carry
primitives emulated on the host, so the exact Montgomery / XYZZ algorithms that run on the GPU are
compiled with g++ and checked against the C oracle."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name,defs", [("fp_host_test", []), ("ec_host_test", []),
                                       ("fp_host_test", ["-DSBN_HOST_FAST_FP=1"]), ("ec_host_test", ["-DSBN_HOST_FAST_FP=1"])])
def test_host_emulated_device_arithmetic(orc, tmp_path, name, defs):
    """Without defines: the 32-bit DEVICE algorithm with the PTX carry flag emulated.  With SBN_HOST_FAST_FP: the 4 x 64-bit
    product the library's own host code uses (transcript challenges, UniPoly arithmetic, point compression)."""
    exe = str(tmp_path / name)
    build_dir = os.path.join(ROOT, "oracle", "_build")
    subprocess.check_call(["g++", "-O2", "-std=c++17"] + defs + ["-o", exe, os.path.join(ROOT, "tests", "host", name + ".cpp"),
                           "-L" + build_dir, "-loracle", "-Wl,-rpath," + build_dir])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "FAIL" not in out.stdout


def test_product_keccak_matches_hashlib(tmp_path):
    """FIPS-202 in the C++ host mirror (used for MultiCommitGens::new) against hashlib."""
    import hashlib
    exe = str(tmp_path / "keccak_test")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "host", "keccak_test.cpp")])
    for m, r in ((b"abc", 1), (b"", 1), (b"0123456789", 30), (b"x", 136), (b"y", 135)):
        out = subprocess.run([exe, m.decode(), str(r)], capture_output=True, text=True).stdout.split()
        msg = m * r
        assert out[0] == hashlib.sha3_256(msg).hexdigest()
        assert out[1] == hashlib.shake_256(msg).hexdigest(200)


def test_signed_window_recoding_of_the_tabulated_sum_path(tmp_path):
    """digits.cuh (what k_mult_entries runs per scalar) compiled for the host: for window widths 8..16 the signed digits
    recompose to the scalar, stay within [-2^(c-1), 2^(c-1)], number ceil(255 / c), and no carry leaves the top digit --
    on random scalars below r and on the values that sit on the borrow boundary of every window."""
    import random
    R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
    exe = str(tmp_path / "digits_host_test")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "host", "digits_host_test.cpp")])
    rnd = random.Random(5)
    vals = [0, 1, 2, R - 1, R - 2, (1 << 253), (1 << 253) + 1, int.from_bytes(b"\x80" * 31, "little"),
            int.from_bytes(b"\x7f" * 31, "little"), int.from_bytes(b"\x81" * 31, "little"), (1 << 254) - 1 - (1 << 200)]
    vals += [sum((1 << (c - 1)) << (k * c) for k in range(254 // c)) % R for c in range(8, 17)]          # every window exactly at 2^(c-1)
    vals += [sum(((1 << (c - 1)) + 1) << (k * c) for k in range(254 // c)) % R for c in range(8, 17)]    # ... and one above it
    vals += [rnd.randrange(R) for _ in range(40)]
    out = subprocess.run([exe] + ["%064x" % v for v in vals], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0
    lines = out.stdout.split("\n")
    idx = 0
    for v in vals:
        for c in range(8, 17):
            f = [int(x) for x in lines[idx].split()]
            idx += 1
            assert f[0] == c and f[1] == -(-255 // c) and len(f) == f[1] + 3
            digits, carry = f[2:-1], f[-1]
            assert carry == 0
            assert all(abs(d) <= 1 << (c - 1) for d in digits)
            assert sum(d << (k * c) for k, d in enumerate(digits)) == v, (hex(v), c)
