"""ctypes front-end of oracle/halo2_oracle.c (the C restatement / restated CPU baseline).

TEST INFRASTRUCTURE ONLY — see the header of oracle/halo2_oracle.c.  Arrays use the wire
layout (uint64 limbs, Montgomery form), identical to what crosses the product's C ABI.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "_build" / "libhalo2_oracle.so"
_lib = None


def build(force: bool = False) -> Path:
    src = HERE / "halo2_oracle.c"
    if force or not LIB.exists() or LIB.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-s", "-C", str(HERE)] + (["-B"] if force else []), check=True)
    return LIB


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(LIB))
    return _lib


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def _p(a: np.ndarray) -> C.c_void_p:
    assert a.dtype == np.uint64 and a.flags.c_contiguous
    return C.c_void_p(a.ctypes.data)


def best_multiexp(coeffs: np.ndarray, bases: np.ndarray, threads: int | None = None) -> np.ndarray:
    """arithmetic.rs best_multiexp -> 12-limb Jacobian (Montgomery)."""
    assert coeffs.shape[0] == bases.shape[0]
    out = np.zeros(12, dtype=np.uint64)
    lib().orc_best_multiexp(_p(coeffs), _p(bases), C.c_size_t(coeffs.shape[0]), threads or host_threads(), _p(out))
    return out


def g1_to_affine(jac12: np.ndarray) -> np.ndarray:
    out = np.zeros(8, dtype=np.uint64)
    lib().orc_g1_to_affine(_p(np.ascontiguousarray(jac12)), _p(out))
    return out


def best_fft(a: np.ndarray, omega: np.ndarray, log_n: int, threads: int | None = None) -> np.ndarray:
    out = np.ascontiguousarray(a).copy()
    lib().orc_best_fft(_p(out), _p(omega), log_n, threads or host_threads())
    return out


def ifft(a, omega_inv, log_n, divisor, threads=None) -> np.ndarray:
    out = np.ascontiguousarray(a).copy()
    lib().orc_ifft(_p(out), _p(omega_inv), log_n, _p(divisor), threads or host_threads())
    return out


def coeff_to_extended(a, k, ext_k, ext_omega, zeta, threads=None) -> np.ndarray:
    out = np.zeros((1 << ext_k, 4), dtype=np.uint64)
    lib().orc_coeff_to_extended(_p(np.ascontiguousarray(a)), k, _p(out), ext_k, _p(ext_omega), _p(zeta),
                                threads or host_threads())
    return out


def extended_to_coeff(a, ext_k, ext_omega_inv, divisor, zeta, keep, threads=None) -> np.ndarray:
    out = np.ascontiguousarray(a).copy()
    lib().orc_extended_to_coeff(_p(out), ext_k, _p(ext_omega_inv), _p(divisor), _p(zeta), threads or host_threads())
    return out[:keep].copy()


def divide_by_vanishing(h, ext_k, t_eval, threads=None) -> np.ndarray:
    out = np.ascontiguousarray(h).copy()
    t_eval = np.ascontiguousarray(t_eval)
    lib().orc_divide_by_vanishing(_p(out), ext_k, _p(t_eval), C.c_size_t(t_eval.shape[0]), threads or host_threads())
    return out


def gen_scalars(seed: int, n: int, start: int = 0) -> np.ndarray:
    out = np.zeros((n, 4), dtype=np.uint64)
    lib().orc_gen_scalars(_p(out), C.c_size_t(n), C.c_uint64(seed), C.c_size_t(start))
    return out


def gen_points(seed: int, n: int, start: int = 0, threads=None) -> np.ndarray:
    out = np.zeros((n, 8), dtype=np.uint64)
    lib().orc_gen_points(_p(out), C.c_size_t(n), C.c_uint64(seed), C.c_size_t(start), threads or host_threads())
    return out


def g1_generator_mul(scalars: np.ndarray, threads=None) -> np.ndarray:
    """[s_i] G for Montgomery scalars (n, 4) -> affine Montgomery points (n, 8)."""
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64)
    out = np.zeros((scalars.shape[0], 8), dtype=np.uint64)
    lib().orc_g1_generator_mul(_p(scalars), C.c_size_t(scalars.shape[0]), _p(out), threads or host_threads())
    return out
