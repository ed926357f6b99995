"""Known answers that pin the oracle's G1 arithmetic and wire formats (oracle/bn254.py)."""
import random

import pytest

from oracle import bn254 as bn

# 2 * (1, 2) on alt_bn128: the public EIP-196 / py_ecc known answer
TWO_G = (1368015179489954701390400359078579693043519447331113978918064868415326638035,
         9918110051302171585080402603319702774565515993150576347155970296011118125764)


def test_public_known_answer_for_doubling_the_generator():
    assert bn.g1_mul((1, 2), 2) == TWO_G == bn.g1_add((1, 2), (1, 2))
    assert bn.g1_is_on_curve(TWO_G)
    # group order: r * G = identity (r from reference contract.sol:211)
    assert bn.g1_mul((1, 2), bn.R) is None


def test_to_bytes_layout():
    assert bn.g1_to_bytes(None) == bytes(32)
    g = bn.g1_to_bytes((1, 2))
    assert g == bytes([1]) + bytes(31)                      # y = 2 is even: flag clear
    neg = bn.g1_to_bytes((1, bn.Q - 2))                     # -G: y odd
    assert neg[0] == 1 and neg[31] == 0x80 and neg[1:31] == bytes(30)
    assert bn.g1_to_evm_bytes((1, 2)) == (1).to_bytes(32, "big") + (2).to_bytes(32, "big")


def test_bytes_round_trip_and_rejections():
    rnd = random.Random(6)
    for _ in range(20):
        p = bn.g1_mul((1, 2), rnd.randrange(1, bn.R))
        assert bn.g1_from_bytes(bn.g1_to_bytes(p)) == p
    assert bn.g1_from_bytes(bytes(32)) is None
    with pytest.raises(ValueError, match="canonical"):
        bn.g1_from_bytes((bn.Q + 1).to_bytes(32, "little"))
    # x = 4: 4^3 + 3 = 67 is a non-residue mod q?  find one deterministically
    x = next(x for x in range(2, 50) if pow((x ** 3 + 3) % bn.Q, (bn.Q - 1) // 2, bn.Q) != 1)
    with pytest.raises(ValueError, match="curve"):
        bn.g1_from_bytes(x.to_bytes(32, "little"))
