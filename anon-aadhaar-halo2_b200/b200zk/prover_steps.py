"""Host-side mirror of the prover loops either side of the hot path (SURVEY.md section 8 f):
`batch_invert`, `eval_polynomial`, `kate_division` ([DEP] halo2_proofs/src/arithmetic.rs, ff
`BatchInvert`) and the grand products of the permutation and lookup arguments
([DEP] halo2_proofs/src/plonk/{permutation,lookup}/prover.rs).  Same names and argument meaning
as upstream; arrays are (n, 4) uint64 Montgomery limbs, all array work happens on the device.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import check, load
from .api import FR_MODULUS, FR_ROOT_OF_UNITY, _as_fr_scalar, _check_vec, _ptr, fr_limbs
from .quotient import FR_DELTA, DeviceColumn


def batch_invert(a: np.ndarray) -> None:
    """`a.iter_mut().batch_invert()`: in place, zeros stay zero."""
    _check_vec(a, 4, "a")
    check(load().b200zk_batch_invert(_ptr(a), a.shape[0]))


def eval_polynomial(poly: np.ndarray, point) -> np.ndarray:
    """`eval_polynomial(&poly, point)` -> 4 limbs."""
    return eval_polynomial_many(poly[None, ...], [point])[0]


def eval_polynomial_many(polys: np.ndarray, points) -> np.ndarray:
    """`polys[c]` evaluated at `points[c]` in one device pipeline; polys is (count, n, 4)."""
    assert polys.dtype == np.uint64 and polys.ndim == 3 and polys.shape[2] == 4
    count, n = polys.shape[0], polys.shape[1]
    pts = np.stack([_as_fr_scalar(p) for p in points]) if count else np.zeros((0, 4), np.uint64)
    assert pts.shape[0] == count
    out = np.zeros((count, 4), dtype=np.uint64)
    if count == 0:
        return out
    col = DeviceColumn.from_host(np.ascontiguousarray(polys).reshape(-1, 4)) if n else None
    check(load().b200zk_eval_polynomial_dev(C.c_void_p(col.ptr if col else 0), n, count, n, _ptr(pts), _ptr(out), None))
    if col:
        col.free()
    return out


def kate_division(a: np.ndarray, b) -> np.ndarray:
    """`kate_division(&a, b)`: a(X) / (X - b), len(a) - 1 coefficients."""
    _check_vec(a, 4, "a")
    assert a.shape[0] >= 1
    n = a.shape[0]
    if n == 1:
        return np.zeros((0, 4), dtype=np.uint64)
    src, dst = DeviceColumn.from_host(a), DeviceColumn(n - 1)
    bl = _as_fr_scalar(b)                       # keep the limbs alive across the call
    check(load().b200zk_kate_division_dev(C.c_void_p(src.ptr), n, _ptr(bl), C.c_void_p(dst.ptr), None))
    out = dst.to_host()
    src.free(); dst.free()
    return out


def _ptr_array(cols):
    return (C.c_void_p * len(cols))(*[c.ptr for c in cols])


def permutation_products(values, sigma, chunk_len: int, k: int, beta, gamma, blinding_factors: int,
                         blinds: np.ndarray = None) -> np.ndarray:
    """The z polynomials of `permutation::Argument::commit` (Lagrange basis): `values[j]` and
    `sigma[j]` are column j's values and permutation polynomial, (2^k, 4) each.  Returns
    (sets, 2^k, 4).  `blinds` (sets, blinding_factors, 4): the scalars for the blinded rows."""
    n, n_cols = 1 << k, len(values)
    assert len(sigma) == n_cols and all(v.shape == (n, 4) for v in values) and all(s.shape == (n, 4) for s in sigma)
    n_sets = -(-n_cols // chunk_len)
    omega = pow(FR_ROOT_OF_UNITY, 1 << (28 - k), FR_MODULUS)
    dv = [DeviceColumn.from_host(v) for v in values]
    ds = [DeviceColumn.from_host(s) for s in sigma]
    z = DeviceColumn(n_sets * n)
    if blinds is not None:
        blinds = np.ascontiguousarray(blinds, dtype=np.uint64)
        assert blinds.shape == (n_sets, blinding_factors, 4)
    bl, gl, wl, dl = _as_fr_scalar(beta), _as_fr_scalar(gamma), fr_limbs(omega), fr_limbs(FR_DELTA)   # keep alive
    pv, psg = _ptr_array(dv), _ptr_array(ds)
    check(load().b200zk_permutation_product_dev(pv, psg, n_cols, chunk_len, k, _ptr(bl), _ptr(gl), _ptr(wl), _ptr(dl),
                                                blinding_factors,
                                                _ptr(blinds) if blinds is not None else None, C.c_void_p(z.ptr), None))
    out = z.to_host().reshape(n_sets, n, 4)
    for c in dv + ds + [z]:
        c.free()
    return out


def lookup_products(compressed_inputs, compressed_tables, permuted_inputs, permuted_tables, k: int, beta, gamma,
                    blinding_factors: int, blinds: np.ndarray = None) -> np.ndarray:
    """`lookup::prover::Permuted::commit_product` for a list of lookups -> (count, 2^k, 4)."""
    n, count = 1 << k, len(compressed_inputs)
    groups = [compressed_inputs, compressed_tables, permuted_inputs, permuted_tables]
    assert all(len(g) == count for g in groups)
    dev = [[DeviceColumn.from_host(a) for a in g] for g in groups]
    z = DeviceColumn(count * n)
    if blinds is not None:
        blinds = np.ascontiguousarray(blinds, dtype=np.uint64)
        assert blinds.shape == (count, blinding_factors, 4)
    bl, gl = _as_fr_scalar(beta), _as_fr_scalar(gamma)                       # keep alive across the call
    ptrs = [_ptr_array(g) for g in dev]
    check(load().b200zk_lookup_product_dev(ptrs[0], ptrs[1], ptrs[2], ptrs[3], count, k, _ptr(bl), _ptr(gl),
                                           blinding_factors, _ptr(blinds) if blinds is not None else None,
                                           C.c_void_p(z.ptr), None))
    out = z.to_host().reshape(count, n, 4)
    for g in dev:
        for c in g:
            c.free()
    z.free()
    return out


def permute_expression_pairs(inputs: np.ndarray, tables: np.ndarray, k: int, blinding_factors: int,
                             blinds: np.ndarray = None):
    """`lookup::prover::permute_expression_pair` for a batch of lookups: inputs / tables are
    (count, 2^k, 4); returns (permuted_inputs, permuted_tables) of the same shape.  `blinds`
    (count, 2, blinding_factors + 1, 4): the random scalars of the last rows (input, then table).
    Raises B200zkError("...ConstraintSystemFailure...") when an input value is not in the table."""
    n = 1 << k
    assert inputs.dtype == np.uint64 and inputs.shape == tables.shape and inputs.shape[1:] == (n, 4)
    count = inputs.shape[0]
    din = DeviceColumn.from_host(np.ascontiguousarray(inputs).reshape(-1, 4))
    dtb = DeviceColumn.from_host(np.ascontiguousarray(tables).reshape(-1, 4))
    oin, otb = DeviceColumn(count * n), DeviceColumn(count * n)
    if blinds is not None:
        blinds = np.ascontiguousarray(blinds, dtype=np.uint64)
        assert blinds.shape == (count, 2, blinding_factors + 1, 4)
    try:
        check(load().b200zk_permute_expression_pair_dev(C.c_void_p(din.ptr), C.c_void_p(dtb.ptr), n, count, k,
                                                        blinding_factors, _ptr(blinds) if blinds is not None else None,
                                                        C.c_void_p(oin.ptr), C.c_void_p(otb.ptr), n, None))
        return oin.to_host().reshape(count, n, 4), otb.to_host().reshape(count, n, 4)
    finally:
        for c in (din, dtb, oin, otb):
            c.free()


def linear_combination(polys, coeffs) -> np.ndarray:
    """sum_j coeffs[j] * polys[j] (the `acc * y + poly` chains of the multi-open provers)."""
    count = len(polys)
    assert count == len(coeffs) and count >= 1
    n = polys[0].shape[0]
    dev = [DeviceColumn.from_host(p) for p in polys]
    out = DeviceColumn(n)
    cl = np.stack([_as_fr_scalar(c) for c in coeffs])
    ptrs = _ptr_array(dev)
    check(load().b200zk_linear_combination_dev(ptrs, _ptr(cl), count, n, C.c_void_p(out.ptr), None))
    res = out.to_host()
    for c in dev + [out]:
        c.free()
    return res
