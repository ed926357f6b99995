"""Field arithmetic on the device against Python integers (the known-answer test of
halo2curves `bn256::{Fr, Fq}` — Mul, Add, Sub, square, invert, neg, double, Montgomery conversions, pow — element
by element through `b200zk_field_op`).  Operands cover every representative a kernel can hold: the library keeps
values in [0, 2p) between kernels (lazy reduction), so canonical and non-canonical limbs of the same residue, 0, p,
1, p - 1, 2p - 1, all-ones low limbs and random values all go through every operation; results must be the
canonical limbs of the exact integer result."""
import ctypes as C
import random

import numpy as np
import pytest

from oracle import bn254 as bn

pytestmark = pytest.mark.gpu

MUL, ADD, SUB, SQR, INV, NEG, DBL, FROM_MONT, TO_MONT, POW = range(10)
R_INV = {p: pow(bn.MONT_R, -1, p) for p in (bn.R, bn.Q)}


def _operands(p: int, n: int, seed: int):
    rng = random.Random(seed)
    edge = [0, 1, 2, p - 1, p - 2, p, p + 1, 2 * p - 1, 2 * p - 2, (1 << 64) - 1, (1 << 128) - 1, (1 << 192) - 1,
            (1 << 253) - 1, (1 << 254) - 1, p >> 1, (p >> 1) + 1, bn.MONT_R % p, (bn.MONT_R % p) + p,
            0xFFFFFFFF, 0xFFFFFFFF00000000, (1 << 224) - (1 << 32)]
    edge = [e for e in edge if e < 2 * p]
    xs = edge + [rng.randrange(2 * p) for _ in range(n - len(edge))]
    rng.shuffle(xs)
    return xs


def _run(zk, field: int, op: int, a, b=None):
    lib = zk.load()
    from b200zk.api import _ptr
    A = bn.ints_to_array(a)
    out = np.zeros_like(A)
    B = bn.ints_to_array(b) if b is not None else None
    zk.check(lib.b200zk_field_op(field, op, _ptr(A), _ptr(B) if B is not None else None, len(a), _ptr(out)))
    return bn.array_to_ints(out)


@pytest.mark.parametrize("field,p", [(0, bn.R), (1, bn.Q)])
def test_field_operations_match_python_integers(zk, field, p):
    n = 1 << 16
    a, b = _operands(p, n, 17 + field), _operands(p, n, 29 + field)
    ri = R_INV[p]
    assert _run(zk, field, MUL, a, b) == [x * y * ri % p for x, y in zip(a, b)]        # Montgomery product
    assert _run(zk, field, ADD, a, b) == [(x + y) % p for x, y in zip(a, b)]
    assert _run(zk, field, SUB, a, b) == [(x - y) % p for x, y in zip(a, b)]
    assert _run(zk, field, SQR, a) == [x * x * ri % p for x in a]
    assert _run(zk, field, NEG, a) == [(-x) % p for x in a]
    assert _run(zk, field, DBL, a) == [2 * x % p for x in a]
    assert _run(zk, field, FROM_MONT, a) == [x * ri % p for x in a]
    assert _run(zk, field, TO_MONT, a) == [x * bn.MONT_R % p for x in a]


@pytest.mark.parametrize("field,p", [(0, bn.R), (1, bn.Q)])
def test_field_inverse_and_pow(zk, field, p):
    n = 1 << 12
    a = _operands(p, n, 41 + field)
    # inverse in the Montgomery domain: a = x R  ->  x^-1 R; 0 -> 0
    got = _run(zk, field, INV, a)
    want = [(pow(x * R_INV[p] % p, -1, p) * bn.MONT_R % p) if x % p else 0 for x in a]
    assert got == want
    # a * a^-1 = 1 (Montgomery one) through the device product as well
    one = bn.MONT_R % p
    prod = _run(zk, field, MUL, a, got)
    assert all(v == (one if x % p else 0) for v, x in zip(prod, a))
    # pow with a full-width exponent
    rng = random.Random(5 + field)
    e = [0, 1, 2, p - 1, p - 2, (1 << 256) - 1] + [rng.randrange(1 << 256) for _ in range(n - 6)]
    got = _run(zk, field, POW, a, e)
    assert got == [pow(x * R_INV[p] % p, k, p) * bn.MONT_R % p for x, k in zip(a, e)]


def test_field_op_argument_checks(zk):
    lib = zk.load()
    from b200zk.api import _ptr
    a = np.zeros((4, 4), dtype=np.uint64)
    assert lib.b200zk_field_op(2, MUL, _ptr(a), _ptr(a), 4, _ptr(a)) != 0
    assert lib.b200zk_field_op(0, 99, _ptr(a), _ptr(a), 4, _ptr(a)) != 0
    assert lib.b200zk_field_op(0, MUL, _ptr(a), None, 4, _ptr(a)) != 0
    assert lib.b200zk_field_op(0, SQR, _ptr(a), None, 0, _ptr(a)) == 0
