"""CPU tests of the drop-in boundary: libsbn254.so loads and exports every symbol declared in
include/sbn254.h (no compute calls -- there is no GPU here), and fails loudly without a device."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "sbn254.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sbn_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    from spartan_bn254_b200 import build
    from spartan_bn254_b200.lib import load_library, EXPORTS
    build.build()
    lib = load_library()
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/sbn254.h but not exported"
    assert sorted(EXPORTS) == syms


def test_no_cpu_fallback():
    """Without a CUDA device the product refuses to run instead of computing on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from spartan_bn254_b200 import Context, SbnError
    with pytest.raises(SbnError):
        Context(0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "spartan_bn254_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower().replace("# noqa", ""), f"{f} references the oracle"


def test_strerror_and_shapes_without_gpu():
    from spartan_bn254_b200.lib import load_library
    lib = load_library()
    assert lib.sbn_strerror(0) == b"ok"
    assert lib.sbn_strerror(-2) == b"shape precondition violated"
    assert lib.sbn_version() >= 1
