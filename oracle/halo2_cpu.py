"""CPU restatement (Python big integers) of the halo2_proofs hot-path functions.

TEST INFRASTRUCTURE ONLY — see the header of ``oracle/bn254.py``.  PARITY PIN: no
upstream output vectors exist (the reference never calls any of these), so
this file restates the *published algorithm* of the pinned dependency, anchored on
mathematical uniqueness (a DFT over Fr / a group element of G1) and the constants the
reference pins in ``solidity_verifier_contract/contract.sol``; end to end it is pinned by
that contract accepting the proofs built from these functions
(``tests/test_square_proof_oracle.py``).

Upstream being restated ([DEP] halo2_proofs 0.2.0 @ v2023_01_20, ``Cargo.lock:469-471``):
  * ``halo2_proofs/src/arithmetic.rs``      best_fft, recursive_butterfly_arithmetic,
                                            best_multiexp, multiexp_serial
  * ``halo2_proofs/src/poly/domain.rs``     EvaluationDomain::{new, lagrange_to_coeff,
                                            coeff_to_extended, extended_to_coeff,
                                            divide_by_vanishing_poly, rotate_extended}
  * ``halo2_proofs/src/poly/kzg/commitment.rs``  ParamsKZG::{commit, commit_lagrange}

All values are canonical Python ints; conversion to/from the Montgomery wire layout
is in ``oracle/bn254.py``.
"""
from __future__ import annotations

import math

from . import bn254 as bn

R = bn.R


# ------------------------------------------------------------------------------ NTT
def dft_naive(a, omega):
    """Definition: out[i] = sum_j a[j] * omega^(i*j).  O(n^2); small n only."""
    n = len(a)
    pw = [pow(omega, i, R) for i in range(n)]
    return [sum(a[j] * pw[(i * j) % n] for j in range(n)) % R for i in range(n)]


def _bitreverse(x: int, bits: int) -> int:
    r = 0
    for _ in range(bits):
        r = (r << 1) | (x & 1)
        x >>= 1
    return r


def best_fft(a, omega, log_n):
    """arithmetic.rs best_fft: bit-reverse permutation, n/2 precomputed twiddles,
    radix-2 decimation-in-time layers with chunk = 2, 4, ...  Natural order in and out.
    Returns a new list (upstream works in place)."""
    n = len(a)
    assert n == 1 << log_n
    a = list(a)
    for k in range(n):
        rk = _bitreverse(k, log_n)
        if k < rk:
            a[k], a[rk] = a[rk], a[k]
    tw = [1] * max(n // 2, 1)
    for i in range(1, n // 2):
        tw[i] = tw[i - 1] * omega % R
    chunk, tchunk = 2, n // 2
    for _ in range(log_n):
        half = chunk // 2
        for base in range(0, n, chunk):
            for i in range(half):
                t = a[base + half + i] * tw[i * tchunk] % R
                u = a[base + i]
                a[base + i] = (u + t) % R
                a[base + half + i] = (u - t) % R
        chunk *= 2
        tchunk //= 2
    return a


class EvaluationDomain:
    """poly/domain.rs EvaluationDomain::new(j, k) and the transforms on it."""

    def __init__(self, j: int, k: int):
        self.k = k
        self.n = 1 << k
        self.quotient_poly_degree = j - 1
        ext_k = k
        while (1 << ext_k) < self.n * self.quotient_poly_degree:
            ext_k += 1
        self.extended_k = ext_k
        self.extended_n = 1 << ext_k
        w = bn.FR_ROOT_OF_UNITY
        for _ in range(ext_k, bn.FR_S):
            w = w * w % R
        self.extended_omega = w
        self.extended_omega_inv = pow(w, -1, R)
        for _ in range(k, ext_k):
            w = w * w % R
        self.omega = w
        self.omega_inv = pow(w, -1, R)
        self.g_coset = bn.FR_ZETA
        self.g_coset_inv = bn.FR_ZETA * bn.FR_ZETA % R
        self.ifft_divisor = pow(self.n, -1, R)
        self.extended_ifft_divisor = pow(self.extended_n, -1, R)
        # t(X) = X^n - 1 on the zeta-coset: period 2^(ext_k - k); stored inverted
        orig = pow(bn.FR_ZETA, self.n, R)
        step = pow(self.extended_omega, self.n, R)
        t, cur = [], orig
        while True:
            t.append(cur)
            cur = cur * step % R
            if cur == orig:
                break
        assert len(t) == 1 << (ext_k - k)
        self.t_evaluations = [pow((v - 1) % R, -1, R) for v in t]

    # -- transforms ---------------------------------------------------------------
    def lagrange_to_coeff(self, a):
        assert len(a) == self.n
        out = best_fft(a, self.omega_inv, self.k)
        return [v * self.ifft_divisor % R for v in out]

    def coeff_to_lagrange(self, a):
        assert len(a) == self.n
        return best_fft(a, self.omega, self.k)

    def _distribute_powers_zeta(self, a, into_coset: bool):
        pw = [self.g_coset, self.g_coset_inv] if into_coset else [self.g_coset_inv, self.g_coset]
        return [v if i % 3 == 0 else v * pw[i % 3 - 1] % R for i, v in enumerate(a)]

    def coeff_to_extended(self, a):
        assert len(a) == self.n
        a = self._distribute_powers_zeta(a, True)
        a = a + [0] * (self.extended_n - self.n)
        return best_fft(a, self.extended_omega, self.extended_k)

    def extended_to_coeff(self, a):
        assert len(a) == self.extended_n
        a = best_fft(a, self.extended_omega_inv, self.extended_k)
        a = [v * self.extended_ifft_divisor % R for v in a]
        a = self._distribute_powers_zeta(a, False)
        return a[: self.n * self.quotient_poly_degree]

    def divide_by_vanishing_poly(self, h):
        assert len(h) == self.extended_n
        m = len(self.t_evaluations)
        return [v * self.t_evaluations[i % m] % R for i, v in enumerate(h)]

    def rotate_extended(self, a, rotation: int):
        """poly/domain.rs rotate_extended: rotate by rotation * 2^(ext_k - k)."""
        s = (rotation << (self.extended_k - self.k)) % self.extended_n
        return a[s:] + a[:s]

    def eval_coeff_on_coset(self, coeffs):
        """Definition check for coeff_to_extended: a(zeta * w_ext^j) by Horner."""
        out = []
        for j in range(self.extended_n):
            x = bn.FR_ZETA * pow(self.extended_omega, j, R) % R
            acc = 0
            for c in reversed(coeffs):
                acc = (acc * x + c) % R
            out.append(acc)
        return out


# ------------------------------------------------------------------------------ MSM
def _window_size(nbases: int) -> int:
    if nbases < 4:
        return 1
    if nbases < 32:
        return 3
    return math.ceil(math.log(nbases))


def _get_at(segment: int, c: int, repr_bytes: bytes) -> int:
    skip_bits = segment * c
    skip_bytes = skip_bits // 8
    if skip_bytes >= 32:
        return 0
    v = repr_bytes[skip_bytes:skip_bytes + 8].ljust(8, b"\0")
    tmp = int.from_bytes(v, "little") >> (skip_bits - skip_bytes * 8)
    return tmp % (1 << c)


def multiexp_serial(coeffs, bases, acc):
    """arithmetic.rs multiexp_serial on Jacobian accumulator ``acc`` (tuple)."""
    reprs = [int(c % R).to_bytes(32, "little") for c in coeffs]
    c = _window_size(len(bases))
    segments = 256 // c + 1
    for seg in reversed(range(segments)):
        for _ in range(c):
            acc = bn._jac_double(acc)
        buckets = [(0, 1, 0)] * ((1 << c) - 1)
        for rb, base in zip(reprs, bases):
            d = _get_at(seg, c, rb)
            if d != 0:
                buckets[d - 1] = bn._jac_add_affine(buckets[d - 1], base)
        running = (0, 1, 0)
        for b in reversed(buckets):
            running = _jac_add(running, b)
            acc = _jac_add(acc, running)
    return acc


def _jac_add(P, Qp):
    if P[2] == 0:
        return Qp
    if Qp[2] == 0:
        return P
    # via affine of Q: exact and simple (this is a model, not a timer)
    return bn._jac_add_affine(P, bn._jac_to_affine(Qp))


def best_multiexp(coeffs, bases, num_threads: int = 8):
    """arithmetic.rs best_multiexp: len/threads chunks, each multiexp_serial, folded.
    Returns the affine point (or None): callers compare in affine."""
    assert len(coeffs) == len(bases)
    n = len(coeffs)
    if n > num_threads:
        chunk = n // num_threads
        total = (0, 1, 0)
        for s in range(0, n, chunk):
            part = multiexp_serial(coeffs[s:s + chunk], bases[s:s + chunk], (0, 1, 0))
            total = _jac_add(total, part)
        return bn._jac_to_affine(total)
    return bn._jac_to_affine(multiexp_serial(coeffs, bases, (0, 1, 0)))


def commit_lagrange(g_lagrange, poly):
    """poly/kzg/commitment.rs ParamsKZG::commit_lagrange: MSM against g_lagrange[..len]."""
    return best_multiexp(list(poly), g_lagrange[: len(poly)])


def commit(g, poly):
    """poly/kzg/commitment.rs ParamsKZG::commit: MSM against g[..len]."""
    return best_multiexp(list(poly), g[: len(poly)])
