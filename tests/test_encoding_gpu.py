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


def test_params_kzg_write_read_round_trip(zk):
    """ParamsKZG::write / read framing around the device point codecs: the re-read tables are
    identical and commit to the same point."""
    n = 1 << 9
    g, gl = co.gen_points(0x9A1, n), co.gen_points(0x9A2, n)
    g2, s_g2 = bytes(range(64)), bytes(range(64, 128))
    blob = zk.ParamsKZG.write_bytes(g, gl, g2, s_g2)
    assert len(blob) == 4 + 64 * n + 128 and blob[:4] == (9).to_bytes(4, "little")
    assert blob[4:36] == bn.g1_to_bytes(bn.g1_affine_array_to_points(g[:1])[0])
    params, g_back, gl_back, g2_back, s_back = zk.ParamsKZG.read_bytes(blob)
    assert np.array_equal(g_back, g) and np.array_equal(gl_back, gl) and g2_back == g2 and s_back == s_g2
    poly = co.gen_scalars(0x9A3, n)
    want = bn.g1_jacobian_limbs_to_affine(co.best_multiexp(poly, gl))
    assert bn.g1_jacobian_limbs_to_affine(params.commit_lagrange(poly)) == want
    params.close()
    with pytest.raises(AssertionError, match="truncated"):
        zk.ParamsKZG.read_bytes(blob[:-1])
