"""GPU parity of the G1 wire formats (csrc/encoding.cu) against oracle/bn254.py, byte for byte."""
import random

import numpy as np
import pytest

from oracle import bn254 as bn
from oracle import c_oracle as co

pytestmark = pytest.mark.gpu


def _jacobian_cases(n, seed):
    """n points as (12-limb Jacobian with random z, affine oracle point), with an identity and
    +-G mixed in."""
    rnd = random.Random(seed)
    pts = [None, (1, 2), (1, bn.Q - 2)] + [bn.g1_mul((1, 2), rnd.randrange(1, bn.R)) for _ in range(n - 3)]
    rows = []
    for p in pts:
        if p is None:
            rows.append([0, bn.to_mont(1, bn.Q), 0])
        else:
            z = rnd.randrange(1, bn.Q)
            rows.append([bn.to_mont(p[0] * z * z % bn.Q, bn.Q), bn.to_mont(p[1] * z * z * z % bn.Q, bn.Q),
                         bn.to_mont(z, bn.Q)])
    arr = bn.ints_to_array([v for r in rows for v in r], width=3)
    return np.ascontiguousarray(arr.reshape(len(pts), 12)), pts


def test_to_bytes_and_evm_bytes(zk):
    arr, pts = _jacobian_cases(40, 1)
    got = zk.g1_to_bytes(arr)
    assert [bytes(r) for r in got] == [bn.g1_to_bytes(p) for p in pts]
    evm = zk.g1_to_evm_bytes(arr)
    assert [bytes(r) for r in evm] == [bn.g1_to_evm_bytes(p) for p in pts]


def test_affine_round_trip_through_bytes(zk):
    n = 3000
    aff = co.gen_points(0xE7C, n)
    aff[5] = 0                                             # an identity in the table
    enc = zk.g1_affine_to_bytes(aff)
    pts = bn.g1_affine_array_to_points(aff)
    for i in list(range(0, n, 97)) + [5]:
        assert bytes(enc[i]) == bn.g1_to_bytes(pts[i])
    back = zk.g1_affine_from_bytes(enc)
    assert np.array_equal(back, aff)


def test_from_bytes_rejects_bad_points(zk):
    good = np.frombuffer(bn.g1_to_bytes((1, 2)), dtype=np.uint8)
    not_canonical = np.frombuffer((bn.Q + 1).to_bytes(32, "little"), dtype=np.uint8)
    x = next(x for x in range(2, 50) if pow((x ** 3 + 3) % bn.Q, (bn.Q - 1) // 2, bn.Q) != 1)
    off_curve = np.frombuffer(x.to_bytes(32, "little"), dtype=np.uint8)
    with pytest.raises(zk.B200zkError, match="point 1.*canonical"):
        zk.g1_affine_from_bytes(np.stack([good, not_canonical]))
    with pytest.raises(zk.B200zkError, match="point 2.*curve"):
        zk.g1_affine_from_bytes(np.stack([good, good, off_curve]))
