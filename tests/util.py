"""Shared helpers for the test suite."""
from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path

import numpy as np

from oracle import bn254 as bn

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = ROOT / "tests" / "golden"
CSRC = ROOT / "anon-aadhaar-halo2_b200" / "csrc"


def omega_for(k: int) -> int:
    w = bn.FR_ROOT_OF_UNITY
    for _ in range(k, bn.FR_S):
        w = w * w % bn.R
    return w


def fr1(x: int) -> np.ndarray:
    """canonical int -> (4,) u64 Montgomery limbs"""
    return bn.fr_array_from_canonical([x])[0]


def jac_affine(limbs12):
    return bn.g1_jacobian_limbs_to_affine(limbs12)


def build_hostlib(name: str) -> ctypes.CDLL:
    """Compile tests/hostlib/<name>.cpp (host emulation bodies of the CUDA headers)."""
    src = ROOT / "tests" / "hostlib" / f"{name}.cpp"
    out = ROOT / "tests" / "hostlib" / f"lib{name}.so"
    deps = [src] + list(CSRC.glob("*.cuh"))
    if not out.exists() or any(d.stat().st_mtime > out.stat().st_mtime for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", "-I", str(CSRC), str(src),
                        "-o", str(out)], check=True)
    return ctypes.CDLL(str(out))
