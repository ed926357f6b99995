"""csrc/field.cuh and csrc/ec.cuh compiled for the host (their plain-C emulation of the
same limb schedule the PTX path runs) against Python big integers."""
import ctypes
import random

import numpy as np
import pytest

from oracle import bn254 as bn
from util import build_hostlib


@pytest.fixture(scope="module")
def flib():
    return build_hostlib("field_host")


@pytest.fixture(scope="module")
def elib():
    return build_hostlib("ec_host")


def _call(lib, fn, *ints):
    bufs = [(ctypes.c_uint32 * 8).from_buffer_copy(int(x).to_bytes(32, "little")) for x in ints]
    out = (ctypes.c_uint32 * 8)()
    getattr(lib, fn)(*bufs, out)
    return int.from_bytes(bytes(out), "little")


@pytest.mark.parametrize("p,pre", [(bn.R, "fr"), (bn.Q, "fq")])
def test_montgomery_field_ops(flib, p, pre):
    rnd = random.Random(7)
    rinv = pow(1 << 256, -1, p)
    edge = [0, 1, 2, p - 1, p - 2, (1 << 256) % p, (p - 1) // 2, (1 << 253) % p, p - (1 << 32), (1 << 32) - 1,
            (1 << 224) - 1, (1 << 64) - 1, 1 << 64]
    vals = edge + [rnd.randrange(p) for _ in range(200)]
    for a in vals:
        for b in rnd.sample(vals, 8) + edge:
            assert _call(flib, f"h_{pre}_mul", a, b) == a * b * rinv % p
            assert _call(flib, f"h_{pre}_add", a, b) == (a + b) % p
            assert _call(flib, f"h_{pre}_sub", a, b) == (a - b) % p
        assert _call(flib, f"h_{pre}_neg", a) == (-a) % p
    # lazy representatives: inputs anywhere in [0, 2p), raw outputs stay in [0, 2p)
    lazy = vals[:40] + [v + p for v in vals[:40]]
    for a in lazy:
        for b in rnd.sample(lazy, 6) + [p, p - 1 + p, 2 * p - 1, 0]:
            for op, exp in (("mul", a * b * rinv), ("add", a + b), ("sub", a - b)):
                raw = _call(flib, f"h_{pre}_{op}_raw", a, b)
                assert raw < 2 * p and raw % p == exp % p, (op, hex(a), hex(b))
                assert _call(flib, f"h_{pre}_{op}", a, b) == exp % p
        assert bool(getattr(flib, f"h_{pre}_is_zero")((ctypes.c_uint32 * 8).from_buffer_copy(int(a).to_bytes(32, "little")))) == (a % p == 0)
    for a in vals[:24]:
        am = a * (1 << 256) % p
        if a:
            assert _call(flib, f"h_{pre}_inv", am) == pow(a, -1, p) * (1 << 256) % p
    assert _call(flib, f"h_{pre}_one") == (1 << 256) % p


@pytest.mark.parametrize("p,pre", [(bn.R, "fr"), (bn.Q, "fq")])
def test_dedicated_square_is_the_same_representative_as_the_product(flib, p, pre):
    """sqr() (36 + 64 + 8 multiplier operations) returns bit for bit what a * a returns, for every
    representative in [0, 2p): the host bodies run the same limb schedule as the PTX path."""
    rnd = random.Random(11)
    rinv = pow(1 << 256, -1, p)
    vals = [0, 1, 2, p - 1, p, p + 1, 2 * p - 1, (1 << 32) - 1, 1 << 32, (1 << 64) - 1, (1 << 224) - 1, (1 << 254) - 1,
            int("ffffffff" * 8, 16) % (2 * p), p - (1 << 32), 2 * p - (1 << 224)]
    vals += [rnd.randrange(2 * p) for _ in range(600)]
    vals += [((1 << 32) - 1) << (32 * i) for i in range(8)] + [(((1 << 32) - 1) << (32 * i)) | ((1 << 32) - 1) for i in range(8)]
    for a in vals:
        a %= 2 * p
        raw = _call(flib, f"h_{pre}_sqr_raw", a)
        assert raw == _call(flib, f"h_{pre}_mul_raw", a, a), hex(a)
        assert raw < 2 * p and raw % p == a * a * rinv % p


def test_mont_conversion(flib):
    for a in (0, 1, 5, bn.R - 1, 1 << 200):
        m = _call(flib, "h_fr_to_mont", a)
        assert m == bn.to_mont(a, bn.R)
        assert _call(flib, "h_fr_from_mont", m) == a


def _aff(elib, acc, norm=0):
    out = np.zeros(12, dtype=np.uint64)
    elib.h_xyzz_to_jac(acc.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p), norm)
    if norm:
        z = bn.from_mont(bn.array_to_ints(out)[2], bn.Q)
        assert z in (0, 1)
    return bn.g1_jacobian_limbs_to_affine(out)


def _add_aff(elib, acc, p):
    a = np.ascontiguousarray(bn.g1_affine_array_from_points([p]))
    elib.h_xyzz_add_affine(acc.ctypes.data_as(ctypes.c_void_p), a.ctypes.data_as(ctypes.c_void_p))


def _build(elib, pts):
    acc = np.zeros(16, dtype=np.uint64)
    for p in pts:
        _add_aff(elib, acc, p)
    return acc


def _add(elib, a, b):
    c = a.copy()
    elib.h_xyzz_add(c.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p))
    return c


def test_xyzz_mixed_add_all_cases(elib):
    pts = bn.seeded_g1_points(11, 12)
    acc = np.zeros(16, dtype=np.uint64)
    ref = None
    for p in pts[:10] + [pts[3], None, bn.g1_neg(pts[0])]:
        _add_aff(elib, acc, p)
        ref = bn.g1_add(ref, p)
        assert _aff(elib, acc) == ref and _aff(elib, acc, 1) == ref
    acc = _build(elib, [pts[0], pts[0]])                      # P + P -> doubling branch
    assert _aff(elib, acc) == bn.g1_add(pts[0], pts[0])
    _add_aff(elib, acc, bn.g1_neg(bn.g1_add(pts[0], pts[0])))  # P + (-P) -> identity
    assert _aff(elib, acc) is None and _aff(elib, acc, 1) is None


def test_xyzz_full_add_all_cases(elib):
    pts = bn.seeded_g1_points(12, 9)
    A, B = _build(elib, pts[:5]), _build(elib, pts[5:9])
    rA = rB = None
    for p in pts[:5]:
        rA = bn.g1_add(rA, p)
    for p in pts[5:9]:
        rB = bn.g1_add(rB, p)
    assert _aff(elib, _add(elib, A, B)) == bn.g1_add(rA, rB)
    A2 = _build(elib, list(reversed(pts[:5])))               # same point, other representative
    assert _aff(elib, _add(elib, A, A2)) == bn.g1_add(rA, rA)
    An = _build(elib, [bn.g1_neg(p) for p in reversed(pts[:5])])
    assert _aff(elib, _add(elib, A, An)) is None
    Z = np.zeros(16, dtype=np.uint64)
    assert _aff(elib, _add(elib, A, Z)) == rA and _aff(elib, _add(elib, Z, A)) == rA
    D = A.copy()
    elib.h_xyzz_dbl(D.ctypes.data_as(ctypes.c_void_p))
    assert _aff(elib, D) == bn.g1_add(rA, rA)
