"""Host-side mirror of `plonk::evaluation::{GraphEvaluator, Evaluator::evaluate_h}`
([DEP] halo2_proofs/src/plonk/evaluation.rs @ v2023_01_20) over the C ABI.

The Rust shim flattens upstream's `GraphEvaluator` (constants, rotations, calculations)
into the `b200zk_graph` encoding; this module is the same glue in Python so the parity
tests can drive the device path exactly as `create_proof` would: advice / instance
polynomials arrive in coefficient form, are extended on the device
(`coeff_to_extended`), and the gate, permutation and lookup kernels fold their terms
into one extended column in upstream's order.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from ._lib import check, load
from .api import EvaluationDomain, FR_MODULUS, FR_ZETA, _ptr, fr_limbs

FR_DELTA = pow(7, 1 << 28, FR_MODULUS)   # halo2curves Fr::DELTA (reference contract.sol:440)


class Src(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("a", C.c_uint32), ("b", C.c_uint32)]


class Calc(C.Structure):
    _fields_ = [("op", C.c_uint32), ("target", C.c_uint32), ("x", Src), ("y", Src),
                ("parts_off", C.c_uint32), ("parts_len", C.c_uint32)]


class GraphC(C.Structure):
    _fields_ = [("constants", C.c_void_p), ("n_constants", C.c_uint32),
                ("rotations", C.c_void_p), ("n_rotations", C.c_uint32),
                ("calcs", C.c_void_p), ("n_calcs", C.c_uint32),
                ("parts", C.c_void_p), ("n_parts", C.c_uint32),
                ("n_intermediates", C.c_uint32)]


class EnvC(C.Structure):
    _fields_ = [("fixed", C.c_void_p), ("n_fixed", C.c_uint32),
                ("advice", C.c_void_p), ("n_advice", C.c_uint32),
                ("instance", C.c_void_p), ("n_instance", C.c_uint32),
                ("challenges", C.c_void_p), ("n_challenges", C.c_uint32),
                ("beta", C.c_uint64 * 4), ("gamma", C.c_uint64 * 4), ("theta", C.c_uint64 * 4),
                ("y", C.c_uint64 * 4), ("k", C.c_uint32), ("ext_k", C.c_uint32),
                ("range_begin", C.c_uint64), ("range_len", C.c_uint64)]


assert C.sizeof(Src) == 12 and C.sizeof(Calc) == 40


class DeviceColumn:
    """A device-resident vector of Fr (b200zk_dev_* handle)."""

    def __init__(self, n_elems: int):
        self._lib = load()
        h = C.c_uint64(0)
        check(self._lib.b200zk_dev_alloc(n_elems, C.byref(h)))
        self.handle = h.value
        self.n = n_elems

    @classmethod
    def from_host(cls, a: np.ndarray) -> "DeviceColumn":
        a = np.ascontiguousarray(a, dtype=np.uint64)
        assert a.ndim == 2 and a.shape[1] == 4
        col = cls(a.shape[0])
        check(col._lib.b200zk_dev_upload(col.handle, 0, _ptr(a), a.shape[0]))
        return col

    @classmethod
    def view(cls, parent: "DeviceColumn", offset: int, n_elems: int) -> "DeviceColumn":
        """A second handle onto parent[offset : offset + n_elems] (no copy)."""
        col = cls.__new__(cls)
        col._lib = load()
        h = C.c_uint64(0)
        check(col._lib.b200zk_dev_view(parent.handle, offset, n_elems, C.byref(h)))
        col.handle, col.n, col._parent = h.value, n_elems, parent
        return col

    def to_host(self) -> np.ndarray:
        out = np.zeros((self.n, 4), dtype=np.uint64)
        check(self._lib.b200zk_dev_download(self.handle, 0, _ptr(out), self.n))
        return out

    @property
    def ptr(self) -> int:
        return int(self._lib.b200zk_dev_ptr(self.handle) or 0)

    def free(self) -> None:
        if getattr(self, "handle", 0):
            self._lib.b200zk_dev_free(self.handle)
            self.handle = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


@dataclass
class FlatGraph:
    """`GraphEvaluator` flattened: constants (n,4) u64 Montgomery, rotations int32,
    calcs (n,10) u32 rows [op, target, x.kind, x.a, x.b, y.kind, y.a, y.b, parts_off,
    parts_len], parts (m,3) u32."""
    constants: np.ndarray
    rotations: np.ndarray
    calcs: np.ndarray
    parts: np.ndarray
    num_intermediates: int

    def as_c(self):
        self._keep = [np.ascontiguousarray(self.constants, dtype=np.uint64),
                      np.ascontiguousarray(self.rotations, dtype=np.int32),
                      np.ascontiguousarray(self.calcs, dtype=np.uint32),
                      np.ascontiguousarray(self.parts, dtype=np.uint32)]
        k = self._keep
        return GraphC(k[0].ctypes.data, k[0].shape[0], k[1].ctypes.data, k[1].shape[0], k[2].ctypes.data,
                      k[2].shape[0], k[3].ctypes.data, k[3].shape[0], self.num_intermediates)


def _handles(cols) -> np.ndarray:
    return np.array([c.handle for c in cols], dtype=np.uint64)


@dataclass
class ProvingKeyCosets:
    """The parts of `ProvingKey` that evaluate_h reads (all extended-domain columns)."""
    fixed_cosets: list
    l0: DeviceColumn
    l_last: DeviceColumn
    l_active_row: DeviceColumn
    permutation_cosets: list            # pk.permutation.cosets
    permutation_columns: list           # cs.permutation.columns as ("fixed"|"advice"|"instance", index)
    degree: int                         # cs.degree()
    blinding_factors: int               # cs.blinding_factors()


@dataclass
class LookupCommitted:
    """`lookup::prover::Committed`: the three polynomials in coefficient form."""
    product_poly: np.ndarray
    permuted_input_poly: np.ndarray
    permuted_table_poly: np.ndarray


class Evaluator:
    """`plonk::evaluation::Evaluator`: custom-gate graph + one graph per lookup."""

    def __init__(self, custom_gates: FlatGraph, lookups=()):
        self.custom_gates = custom_gates
        self.lookups = list(lookups)
        self._lib = load()

    def _env(self, domain, fixed, advice, instance, challenges, beta, gamma, theta, y, idx_range=None):
        keep = [_handles(fixed), _handles(advice), _handles(instance),
                np.ascontiguousarray(challenges, dtype=np.uint64).reshape(-1, 4)]
        env = EnvC()
        env.fixed, env.n_fixed = keep[0].ctypes.data, len(fixed)
        env.advice, env.n_advice = keep[1].ctypes.data, len(advice)
        env.instance, env.n_instance = keep[2].ctypes.data, len(instance)
        env.challenges, env.n_challenges = keep[3].ctypes.data, keep[3].shape[0]
        for name, v in (("beta", beta), ("gamma", gamma), ("theta", theta), ("y", y)):
            getattr(env, name)[:] = [int(x) for x in np.asarray(v, dtype=np.uint64)]
        env.k, env.ext_k = domain.k, domain.extended_k
        env.range_begin, env.range_len = (idx_range[0], idx_range[1] - idx_range[0]) if idx_range else (0, 0)
        env._keep = keep
        return env

    def _extend(self, domain: EvaluationDomain, polys) -> list:
        """coeff_to_extended on the device for a list of coefficient-form columns."""
        out = []
        for p in polys:
            src = DeviceColumn.from_host(p)
            dst = DeviceColumn(domain.extended_len())
            check(self._lib.b200zk_coeff_to_extended_dev(C.c_void_p(src.ptr), domain.n, C.c_void_p(dst.ptr),
                                                         domain.extended_len(), 1, domain.k, domain.extended_k,
                                                         _ptr(domain.extended_omega), _ptr(domain.g_coset), None))
            out.append(dst)
            src.free()
        return out

    def evaluate_h(self, domain: EvaluationDomain, pk: ProvingKeyCosets, advice_polys, instance_polys, challenges,
                   y, beta, gamma, theta, lookups=(), permutation_products=(), idx_range=None) -> np.ndarray:
        """One circuit instance of `Evaluator::evaluate_h`.  advice / instance / lookup /
        permutation-product polynomials are host arrays in coefficient form, as
        `create_proof` holds them; returns the extended-domain numerator (host array).
        With ``idx_range=(begin, end)`` only that slice of the extended domain is evaluated
        (one rank's share of a multi-GPU evaluation) and only that slice is returned."""
        lib = self._lib
        advice = self._extend(domain, advice_polys)
        instance = self._extend(domain, instance_polys)
        env = self._env(domain, pk.fixed_cosets, advice, instance, challenges, beta, gamma, theta, y, idx_range)
        values = DeviceColumn(domain.extended_len())
        zeros = np.zeros((domain.extended_len(), 4), np.uint64)   # domain.empty_extended()
        check(lib.b200zk_dev_upload(values.handle, 0, _ptr(zeros), domain.extended_len()))
        g = self.custom_gates.as_c()
        check(lib.b200zk_quotient_graph(C.byref(g), C.byref(env), values.handle, values.handle))
        if len(permutation_products):
            products = self._extend(domain, permutation_products)
            kinds = np.array([{"fixed": 2, "advice": 3, "instance": 4}[k] for (k, _) in pk.permutation_columns],
                             dtype=np.uint32)
            idxs = np.array([i for (_, i) in pk.permutation_columns], dtype=np.uint32)
            sig = _handles(pk.permutation_cosets)
            ph = _handles(products)
            zeta, delta = fr_limbs(FR_ZETA), fr_limbs(FR_DELTA)   # keep alive across the call
            check(lib.b200zk_quotient_permutation(
                C.byref(env), values.handle, _ptr32(kinds), _ptr32(idxs), _ptr(sig), len(kinds), _ptr(ph), len(products),
                pk.degree - 2, pk.blinding_factors, pk.l0.handle, pk.l_last.handle, pk.l_active_row.handle,
                _ptr(domain.extended_omega), _ptr(zeta), _ptr(delta)))
            for c in products:
                c.free()
        for n, lk in enumerate(lookups):
            prod, pin, ptab = self._extend(domain, [lk.product_poly, lk.permuted_input_poly, lk.permuted_table_poly])
            table = DeviceColumn(domain.extended_len())
            lg = self.lookups[n].as_c()
            check(lib.b200zk_quotient_graph(C.byref(lg), C.byref(env), 0, table.handle))
            check(lib.b200zk_quotient_lookup(C.byref(env), values.handle, table.handle, prod.handle, pin.handle,
                                             ptab.handle, pk.l0.handle, pk.l_last.handle, pk.l_active_row.handle))
            for c in (prod, pin, ptab, table):
                c.free()
        out = values.to_host()
        if idx_range:
            out = out[idx_range[0]:idx_range[1]].copy()
        for c in advice + instance + [values]:
            c.free()
        return out


def _ptr32(a: np.ndarray) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data)
