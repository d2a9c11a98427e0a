"""CPU tests of the DEVICE arithmetic: fp.cuh / ec.cuh are __host__ __device__ with the PTX carry
primitives emulated on the host, so the exact Montgomery / XYZZ algorithms that run on the GPU are
compiled with g++ and checked against the C oracle."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name", ["fp_host_test", "ec_host_test"])
def test_host_emulated_device_arithmetic(orc, tmp_path, name):
    exe = str(tmp_path / name)
    build_dir = os.path.join(ROOT, "oracle", "_build")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "host", name + ".cpp"),
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
