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


# ---- the Rust `extern "C"` block against the header, argument by argument ------------------------
_C_TO_RUST = {
    "int": "c_int", "uint32_t": "u32", "uint64_t": "u64", "size_t": "usize", "double": "f64", "float": "f32",
    "uint8_t": "u8", "int32_t": "i32", "void": "c_void", "char": "c_char",
    "b200zk_graph": "b200zk_graph", "b200zk_quotient_env": "b200zk_quotient_env",
    "b200zk_src": "b200zk_src", "b200zk_calc": "b200zk_calc",
}


def _c_type_to_rust(decl: str) -> str:
    """`const uint64_t omega[4]` / `void* const* dest_bases` / `size_t n` -> the Rust FFI type."""
    decl = decl.strip()
    array = bool(re.search(r"\[\d*\]\s*$", decl))
    decl = re.sub(r"\[\d*\]\s*$", "", decl).strip()
    # drop the parameter name (last identifier not followed by * and not a type keyword)
    m = re.match(r"^(.*?)([A-Za-z_][A-Za-z0-9_]*)$", decl)
    if m and m.group(2) not in _C_TO_RUST and m.group(2) != "const":
        decl = m.group(1).strip()
    tokens = re.findall(r"const|\*|[A-Za-z_][A-Za-z0-9_]*", decl)
    base = next(t for t in tokens if t in _C_TO_RUST)
    ty = _C_TO_RUST[base]
    # pointer levels, left to right; `const` binds to what precedes it (or the base type when leading)
    i = tokens.index(base)
    const_here = "const" in tokens[:i] or (i + 1 < len(tokens) and tokens[i + 1] == "const")
    rest = [t for t in tokens[i + 1:]]
    if rest and rest[0] == "const":
        rest = rest[1:]
    for j, t in enumerate(rest):
        if t == "*":
            ty = ("*const " if const_here else "*mut ") + ty
            const_here = j + 1 < len(rest) and rest[j + 1] == "const"
    if array:
        ty = ("*const " if const_here or "const" in tokens[:i] else "*mut ") + ty
    return ty


def _header_prototypes():
    text = re.sub(r"/\*.*?\*/", "", (ROOT / "include" / "b200zk.h").read_text(), flags=re.S)
    protos = {}
    for ret, name, args in re.findall(r"([A-Za-z_][A-Za-z0-9_ \*]*?)\s*\b(b200zk_[a-z0-9_]+)\s*\(([^;{}]*?)\)\s*;", text):
        args = " ".join(args.split())
        params = [] if args in ("", "void") else [_c_type_to_rust(a) for a in args.split(",")]
        protos[name] = (params, None if ret.strip() == "void" else _c_type_to_rust(ret.strip() + " x"))
    return protos


def _rust_prototypes():
    text = (ROOT / "anon-aadhaar-halo2_b200" / "rust" / "b200zk-sys" / "src" / "lib.rs").read_text()
    protos = {}
    for name, args, ret in re.findall(r"pub fn (b200zk_[a-z0-9_]+)\(([^)]*)\)\s*(?:->\s*([^;]+))?;", text):
        params = [" ".join(a.split(":", 1)[1].split()) for a in args.split(",") if a.strip()]
        protos[name] = (params, " ".join(ret.split()) if ret else None)
    return protos


def test_c_type_mapping_examples():
    assert _c_type_to_rust("const uint64_t omega[4]") == "*const u64"
    assert _c_type_to_rust("uint64_t out_xyz[12]") == "*mut u64"
    assert _c_type_to_rust("void* const* dest_bases") == "*const *mut c_void"
    assert _c_type_to_rust("const void* const* d_values") == "*const *const c_void"
    assert _c_type_to_rust("void** ptr_out") == "*mut *mut c_void"
    assert _c_type_to_rust("const b200zk_graph* graph") == "*const b200zk_graph"
    assert _c_type_to_rust("size_t n") == "usize" and _c_type_to_rust("double* x") == "*mut f64"


def test_rust_ffi_argument_types_and_order_match_the_header():
    """No cargo here, so the hand-written Rust declarations are checked mechanically: every entry point has
    the same number of parameters, in the same order, with the Rust type the C type maps to, and the same
    return type."""
    c, r = _header_prototypes(), _rust_prototypes()
    assert set(c) == set(r)
    assert len(c) >= 70
    bad = {name: (c[name], r[name]) for name in c if c[name] != r[name]}
    assert not bad, bad
