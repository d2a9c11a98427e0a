"""Builds libsbn254.so (hand-written sm_100a CUDA + C ABI) in-tree with nvcc."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_lib")
OUT = os.path.join(OUT_DIR, "libsbn254.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "--expt-relaxed-constexpr",
    "-shared", "-Xcompiler", "-fPIC",
]


def sources():
    """Every file under csrc/ (the host headers in csrc/host/ are compiled into the library too) plus the C-ABI header."""
    out = [os.path.join(HERE, "..", "include", "sbn254.h")]
    for d, _, files in os.walk(CSRC):
        out += [os.path.join(d, f) for f in sorted(files)]
    return out


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [
        "-o", OUT, os.path.join(CSRC, "sbn254.cu")]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
