"""Host-side mirror of the halo2_proofs items whose bodies the C ABI replaces.

Only per-call *scalars* (domain constants such as omega, zeta, n^-1) are derived here,
with Python integers, exactly as upstream derives them on the CPU in
``EvaluationDomain::new`` ([DEP] halo2_proofs/src/poly/domain.rs).  All array work goes
through libb200zk.so; nothing here touches ``oracle/``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import check, load

# BN254 scalar field (reference solidity_verifier_contract/contract.sol:211) and the
# halo2curves 0.3.1 `bn256::Fr` constants.
FR_MODULUS = 21888242871839275222246405745257275088548364400416034343698204186575808495617
FR_S = 28
FR_ROOT_OF_UNITY = pow(7, (FR_MODULUS - 1) >> FR_S, FR_MODULUS)
FR_ZETA = 0x30644E72E131A029048B6E193FD84104CC37A73FEC2BC5E9B8CA0B2D36636F23
_MONT_R = 1 << 256


def fr_limbs(x: int) -> np.ndarray:
    """Canonical integer -> 4 u64 Montgomery limbs (the `bn256::Fr` wire layout)."""
    v = (int(x) % FR_MODULUS) * _MONT_R % FR_MODULUS
    return np.frombuffer(v.to_bytes(32, "little"), dtype="<u8").copy()


def _as_fr_scalar(x) -> np.ndarray:
    if isinstance(x, (int, np.integer)):
        return fr_limbs(int(x))
    a = np.ascontiguousarray(x, dtype=np.uint64)
    assert a.shape == (4,), "field scalar must be 4 u64 limbs"
    return a


def _ptr(a: np.ndarray) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data)


def _check_vec(a, width: int, name: str) -> np.ndarray:
    assert isinstance(a, np.ndarray) and a.dtype == np.uint64, f"{name}: expected uint64 ndarray"
    assert a.ndim == 2 and a.shape[1] == width, f"{name}: expected shape (n, {width})"
    assert a.flags.c_contiguous, f"{name}: must be C-contiguous"
    return a


def fresh(a: np.ndarray) -> np.ndarray:
    """A newly allocated host buffer about to be handed to the library: no device mirror of an earlier
    buffer at the same address may survive (the contract of b200zk_mirror_enable; the Rust fork does
    this in `Polynomial`'s `Drop`).  A no-op while mirrors are off."""
    check(load().b200zk_mirror_invalidate(C.c_void_p(a.ctypes.data), a.nbytes))
    return a


def init(device: int = -1) -> None:
    check(load().b200zk_init(device))


def shutdown() -> None:
    check(load().b200zk_shutdown())


class pinned:
    """Page-lock a numpy buffer for the duration of a `with` block (b200zk_host_register): what a
    Rust shim does once for the `Vec<Fr>`s a proof keeps, so host-pointer calls transfer at the full
    PCIe rate and the MSM upload pipeline can overlap its copies."""

    def __init__(self, a: np.ndarray):
        assert a.flags.c_contiguous
        self.a = a

    def __enter__(self):
        check(load().b200zk_host_register(C.c_void_p(self.a.ctypes.data), self.a.nbytes))
        return self.a

    def __exit__(self, *exc):
        check(load().b200zk_host_unregister(C.c_void_p(self.a.ctypes.data)))
        return False


def host_alloc_fr(n: int, width: int = 4) -> np.ndarray:
    """An (n, width) uint64 array in page-locked memory from b200zk_host_alloc; release it with
    host_free."""
    p = C.c_void_p(0)
    check(load().b200zk_host_alloc(max(n * width * 8, 8), C.byref(p)))
    buf = (C.c_uint64 * (n * width)).from_address(p.value)
    a = np.frombuffer(buf, dtype=np.uint64).reshape(n, width)
    _host_allocs[a.ctypes.data] = p.value
    return a


def host_free(a: np.ndarray) -> None:
    check(load().b200zk_host_free(C.c_void_p(_host_allocs.pop(a.ctypes.data))))


_host_allocs = {}


def kernel_launches() -> int:
    return int(load().b200zk_kernel_launches())


def modmul_peak(iters: int = 4096) -> float:
    out = C.c_double(0.0)
    check(load().b200zk_modmul_peak(iters, C.byref(out)))
    return out.value


# ------------------------------------------------------------------ arithmetic.rs
def best_fft(a: np.ndarray, omega, log_n: int) -> None:
    """`arithmetic::best_fft::<Fr>(a, omega, log_n)`: in place, natural order."""
    _check_vec(a, 4, "a")
    assert a.shape[0] == 1 << log_n  # upstream: assert_eq!(n, 1 << log_n)
    w = _as_fr_scalar(omega)
    check(load().b200zk_ntt(_ptr(a), log_n, _ptr(w)))


def best_multiexp(coeffs: np.ndarray, bases: np.ndarray) -> np.ndarray:
    """`arithmetic::best_multiexp::<G1Affine>(coeffs, bases) -> G1` (12 limbs, Jacobian)."""
    _check_vec(coeffs, 4, "coeffs")
    _check_vec(bases, 8, "bases")
    assert coeffs.shape[0] == bases.shape[0]  # upstream: assert_eq!(coeffs.len(), bases.len())
    out = np.zeros(12, dtype=np.uint64)
    check(load().b200zk_msm_g1(_ptr(coeffs), _ptr(bases), coeffs.shape[0], _ptr(out)))
    return out


def g1_sum(points: np.ndarray) -> np.ndarray:
    """Fold of Jacobian points (the `results.iter().fold(identity, +)` of best_multiexp)."""
    _check_vec(points, 12, "points")
    out = np.zeros(12, dtype=np.uint64)
    check(load().b200zk_g1_sum(_ptr(points), points.shape[0], _ptr(out)))
    return out


# ------------------------------------------------------------------ poly/domain.rs
class EvaluationDomain:
    """`poly::EvaluationDomain::<Fr>::new(j, k)` and its transforms."""

    def __init__(self, j: int, k: int):
        p = FR_MODULUS
        self.k = k
        self.n = 1 << k
        self.quotient_poly_degree = j - 1
        ext_k = k
        while (1 << ext_k) < self.n * self.quotient_poly_degree:
            ext_k += 1
        assert ext_k <= FR_S, "extended domain exceeds the field's two-adicity"
        self.extended_k = ext_k
        w = FR_ROOT_OF_UNITY
        for _ in range(ext_k, FR_S):
            w = w * w % p
        self.extended_omega_int = w
        for _ in range(k, ext_k):
            w = w * w % p
        self.omega_int = w
        self.extended_omega = fr_limbs(self.extended_omega_int)
        self.extended_omega_inv = fr_limbs(pow(self.extended_omega_int, -1, p))
        self.omega = fr_limbs(self.omega_int)
        self.omega_inv = fr_limbs(pow(self.omega_int, -1, p))
        self.g_coset = fr_limbs(FR_ZETA)
        self.g_coset_inv = fr_limbs(FR_ZETA * FR_ZETA % p)
        self.ifft_divisor = fr_limbs(pow(self.n, -1, p))
        self.extended_ifft_divisor = fr_limbs(pow(1 << ext_k, -1, p))
        # t(X) = X^n - 1 on the coset, period 2^(ext_k - k), stored inverted like upstream
        orig = pow(FR_ZETA, self.n, p)
        step = pow(self.extended_omega_int, self.n, p)
        t, cur = [], orig
        while True:
            t.append(cur)
            cur = cur * step % p
            if cur == orig:
                break
        assert len(t) == 1 << (ext_k - k)
        self.t_evaluations = np.stack([fr_limbs(pow((v - 1) % p, -1, p)) for v in t])

    def extended_len(self) -> int:
        return 1 << self.extended_k

    def lagrange_to_coeff(self, a: np.ndarray) -> np.ndarray:
        _check_vec(a, 4, "a")
        assert a.shape[0] == self.n
        out = fresh(a.copy())
        check(load().b200zk_intt(_ptr(out), self.k, _ptr(self.omega_inv), _ptr(self.ifft_divisor)))
        return out

    def lagrange_to_coeff_many(self, cols: np.ndarray) -> np.ndarray:
        """Batched form: cols is (count, n, 4)."""
        assert cols.dtype == np.uint64 and cols.ndim == 3 and cols.shape[1:] == (self.n, 4)
        out = np.ascontiguousarray(cols).copy()
        check(load().b200zk_intt_many(_ptr(out), self.n, out.shape[0], self.k, _ptr(self.omega_inv),
                                      _ptr(self.ifft_divisor)))
        return out

    def coeff_to_extended(self, a: np.ndarray) -> np.ndarray:
        _check_vec(a, 4, "a")
        assert a.shape[0] == self.n
        out = fresh(np.zeros((self.extended_len(), 4), dtype=np.uint64))
        check(load().b200zk_coeff_to_extended(_ptr(a), self.k, _ptr(out), self.extended_k,
                                              _ptr(self.extended_omega), _ptr(self.g_coset)))
        return out

    def coeff_to_extended_many(self, cols: np.ndarray) -> np.ndarray:
        assert cols.dtype == np.uint64 and cols.ndim == 3 and cols.shape[1:] == (self.n, 4)
        cols = np.ascontiguousarray(cols)
        out = np.zeros((cols.shape[0], self.extended_len(), 4), dtype=np.uint64)
        check(load().b200zk_coeff_to_extended_many(_ptr(cols), self.n, _ptr(out), self.extended_len(),
                                                   cols.shape[0], self.k, self.extended_k,
                                                   _ptr(self.extended_omega), _ptr(self.g_coset)))
        return out

    def extended_to_coeff(self, a: np.ndarray) -> np.ndarray:
        _check_vec(a, 4, "a")
        assert a.shape[0] == self.extended_len()
        keep = self.n * self.quotient_poly_degree
        out = fresh(np.zeros((keep, 4), dtype=np.uint64))
        check(load().b200zk_extended_to_coeff(_ptr(a), self.extended_k, _ptr(self.extended_omega_inv),
                                              _ptr(self.extended_ifft_divisor), _ptr(self.g_coset), _ptr(out),
                                              keep))
        return out

    def divide_by_vanishing_poly(self, h: np.ndarray) -> np.ndarray:
        _check_vec(h, 4, "h")
        assert h.shape[0] == self.extended_len()
        out = fresh(h.copy())
        check(load().b200zk_divide_by_vanishing(_ptr(out), self.extended_k, _ptr(self.t_evaluations),
                                                self.t_evaluations.shape[0]))
        return out


# ------------------------------------------------------- poly/kzg/commitment.rs
class ParamsKZG:
    """`ParamsKZG<Bn256>` restricted to what the hot path uses: the two base tables.

    `g` and `g_lagrange` are uploaded to the device once (they are fixed for the life
    of the params) and evicted when the object is dropped.
    """

    def __init__(self, g: np.ndarray, g_lagrange: np.ndarray, precompute_windows: bool = True):
        _check_vec(g, 8, "g")
        _check_vec(g_lagrange, 8, "g_lagrange")
        assert g.shape[0] == g_lagrange.shape[0]
        self.n = g.shape[0]
        self._lib = load()
        self._pre = 1 if precompute_windows else 0
        self._h_g = self._register(g)
        self._h_gl = self._register(g_lagrange)

    def _register(self, bases: np.ndarray) -> int:
        h = C.c_uint64(0)
        check(self._lib.b200zk_bases_register_ex(_ptr(bases), bases.shape[0], self._pre, C.byref(h)))
        return h.value

    def _msm_many(self, handle: int, polys: np.ndarray) -> np.ndarray:
        assert polys.dtype == np.uint64 and polys.ndim == 3 and polys.shape[2] == 4
        assert polys.shape[1] <= self.n
        polys = np.ascontiguousarray(polys)
        out = np.zeros((polys.shape[0], 12), dtype=np.uint64)
        check(self._lib.b200zk_msm_g1_registered_many(handle, _ptr(polys), polys.shape[1], polys.shape[0],
                                                      polys.shape[1], _ptr(out)))
        return out

    def commit_many(self, polys: np.ndarray) -> np.ndarray:
        """`polys.iter().map(|p| params.commit(p, _))` in one device pipeline; polys is (count, len, 4)."""
        return self._msm_many(self._h_g, polys)

    def commit_lagrange_many(self, polys: np.ndarray) -> np.ndarray:
        return self._msm_many(self._h_gl, polys)

    def _msm(self, handle: int, poly: np.ndarray) -> np.ndarray:
        _check_vec(poly, 4, "poly")
        assert poly.shape[0] <= self.n  # upstream: assert!(bases.len() >= size)
        out = np.zeros(12, dtype=np.uint64)
        check(self._lib.b200zk_msm_g1_registered(handle, _ptr(poly), poly.shape[0], _ptr(out)))
        return out

    def commit(self, poly: np.ndarray) -> np.ndarray:
        """`ParamsKZG::commit(&poly, _blind) -> G1` (the blind is ignored upstream)."""
        return self._msm(self._h_g, poly)

    def commit_lagrange(self, poly: np.ndarray) -> np.ndarray:
        """`ParamsKZG::commit_lagrange(&poly, _blind) -> G1`."""
        return self._msm(self._h_gl, poly)

    # ---- poly/kzg/commitment.rs ParamsKZG::{write, read} (v2023_01_20 framing): k as u32 LE, then
    # g[0..n) and g_lagrange[0..n) as `G1Affine::to_bytes` (32 bytes each), then g2 and s_g2 as
    # `G2Affine::to_bytes` (64 bytes each).  The G2 points never touch the hot path: they are
    # carried through as opaque bytes.  Compression and the n square roots of decompression run
    # on the device (b200zk_g1_affine_to_bytes / b200zk_g1_affine_from_bytes).
    @staticmethod
    def write_bytes(g: np.ndarray, g_lagrange: np.ndarray, g2: bytes, s_g2: bytes) -> bytes:
        n = g.shape[0]
        assert n & (n - 1) == 0 and g_lagrange.shape == g.shape and len(g2) == 64 and len(s_g2) == 64
        k = n.bit_length() - 1
        return (k.to_bytes(4, "little") + g1_affine_to_bytes(g).tobytes() + g1_affine_to_bytes(g_lagrange).tobytes()
                + bytes(g2) + bytes(s_g2))

    @classmethod
    def read_bytes(cls, data: bytes, precompute_windows: bool = True):
        """Returns (ParamsKZG, g, g_lagrange, g2_bytes, s_g2_bytes)."""
        k = int.from_bytes(data[:4], "little")
        n = 1 << k
        assert len(data) == 4 + 64 * n + 128, "truncated params"   # upstream: read_exact fails
        body = np.frombuffer(data, dtype=np.uint8, count=64 * n, offset=4).reshape(2, n, 32)
        g = g1_affine_from_bytes(body[0])
        g_lagrange = g1_affine_from_bytes(body[1])
        return cls(g, g_lagrange, precompute_windows), g, g_lagrange, data[4 + 64 * n: 68 + 64 * n], data[68 + 64 * n:]

    def close(self) -> None:
        for name in ("_h_g", "_h_gl"):
            h = getattr(self, name, 0)
            if h:
                self._lib.b200zk_bases_evict(h)
                setattr(self, name, 0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------- halo2curves derive/curve.rs
def g1_to_bytes(points: np.ndarray) -> np.ndarray:
    """`G1Affine::from(p).to_bytes()` for Jacobian results (count, 12) -> (count, 32) uint8."""
    _check_vec(points, 12, "points")
    out = np.zeros((points.shape[0], 32), dtype=np.uint8)
    check(load().b200zk_g1_to_bytes(_ptr(points), points.shape[0], _ptr(out)))
    return out


def g1_affine_to_bytes(points: np.ndarray) -> np.ndarray:
    """`G1Affine::to_bytes` for (count, 8) affine points (what `ParamsKZG::write` emits)."""
    _check_vec(points, 8, "points")
    out = np.zeros((points.shape[0], 32), dtype=np.uint8)
    check(load().b200zk_g1_affine_to_bytes(_ptr(points), points.shape[0], _ptr(out)))
    return out


def g1_to_evm_bytes(points: np.ndarray) -> np.ndarray:
    """The 64-byte big-endian (x, y) encoding of the EVM transcript / calldata."""
    _check_vec(points, 12, "points")
    out = np.zeros((points.shape[0], 64), dtype=np.uint8)
    check(load().b200zk_g1_to_evm_bytes(_ptr(points), points.shape[0], _ptr(out)))
    return out


def g1_affine_from_bytes(data: np.ndarray) -> np.ndarray:
    """`G1Affine::from_bytes` for (count, 32) uint8 -> (count, 8); raises on an invalid point
    (upstream returns CtOption::none, which `ParamsKZG::read` unwraps)."""
    assert isinstance(data, np.ndarray) and data.dtype == np.uint8 and data.ndim == 2 and data.shape[1] == 32
    data = np.ascontiguousarray(data)
    out = np.zeros((data.shape[0], 8), dtype=np.uint64)
    check(load().b200zk_g1_affine_from_bytes(_ptr(data), data.shape[0], _ptr(out)))
    return out
