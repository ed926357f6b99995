"""The C-ABI shared library: it loads, exports every symbol include/b200zk.h declares,
and refuses to compute without a CUDA device (no CPU fallback).  No compute calls here."""
import ctypes
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def built_lib():
    sys.path.insert(0, str(ROOT / "anon-aadhaar-halo2_b200"))
    import build as zkbuild  # anon-aadhaar-halo2_b200/build.py

    return zkbuild.build()


def test_library_exports_every_declared_symbol(built_lib):
    import b200zk

    lib = ctypes.CDLL(str(built_lib))
    names = b200zk.header_symbols()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"include/b200zk.h declares symbols the library lacks: {missing}"
    assert lib.b200zk_abi_version() == 1


def test_header_compiles_as_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "b200zk.h"\nint main(void){return (int)sizeof(&b200zk_msm_g1) == 0;}\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", str(ROOT / "include"), str(src)],
                   check=True)


def test_no_cpu_fallback_without_device(built_lib):
    """On a box without a GPU every compute entry point must fail loudly."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    import b200zk

    a = np.zeros((8, 4), dtype=np.uint64)
    with pytest.raises(b200zk.B200zkError, match="no CUDA device|no CPU fallback|CUDA"):
        b200zk.best_fft(a, 1, 3)
    with pytest.raises(b200zk.B200zkError):
        b200zk.best_multiexp(a, np.zeros((8, 8), dtype=np.uint64))


def test_product_does_not_import_oracle():
    pkg = ROOT / "anon-aadhaar-halo2_b200"
    offenders = []
    for path in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + list(pkg.rglob("*.hpp")):
        text = path.read_text()
        if re.search(r"^\s*(import|from)\s+oracle\b", text, flags=re.M) or "halo2_oracle" in text or "c_oracle" in text:
            offenders.append(str(path))
    assert not offenders, offenders


def test_rust_ffi_block_declares_every_header_symbol():
    """anon-aadhaar-halo2_b200/rust/b200zk-sys/src/lib.rs cannot be compiled here (no cargo), so at
    least keep its hand-written `extern "C"` block in step with include/b200zk.h."""
    import re
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    header = set(re.findall(r"\b(b200zk_[a-z0-9_]+)\s*\(", (root / "include" / "b200zk.h").read_text()))
    rust = set(re.findall(r"pub fn (b200zk_[a-z0-9_]+)", (root / "anon-aadhaar-halo2_b200" / "rust" / "b200zk-sys" / "src" / "lib.rs").read_text()))
    assert header == rust, (sorted(header - rust), sorted(rust - header))
