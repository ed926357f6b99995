"""CPU: the end-to-end pin of the oracle.  A complete proof of the reference's SquareCircuit
(src/signal.rs), assembled from the oracle's MSM / NTT / quotient / prover-step functions, must be
accepted by the transliteration of the reference's Solidity verifier (oracle/sol_verifier.py <-
solidity_verifier_contract/contract.sol), and every kind of tampering must be rejected.  Also pins
the verifier's own building blocks (Keccak-256, the alt_bn128 pairing) by published answers and
group laws."""
import hashlib

import pytest

import square_proof as sp
from oracle import bn254 as bn
from oracle import bn254_pairing as pr
from oracle import sol_verifier as sv
from oracle.keccak import keccak256
from util import GOLDEN


def test_keccak256_published_digests():
    assert keccak256(b"").hex() == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"
    assert keccak256(b"abc").hex() == "4e03657aea45a94fc7d47ba826c8d667c0d1e6e33a64a036ec44f58fa12d6c45"
    # padding edge cases around the 136-byte rate: one block minus one byte, exactly one block
    assert keccak256(b"a" * 135) != keccak256(b"a" * 136)
    assert len(keccak256(b"a" * 136)) == 32 and len(keccak256(b"a" * 300)) == 32
    assert keccak256(b"abc") != hashlib.sha3_256(b"abc").digest()        # not the NIST padding


def test_pairing_group_laws():
    g1, g2 = bn.G1_GEN, pr.G2_GEN
    assert pr.g2_is_on_curve(g2)
    assert pr.g2_add(pr.g2_mul(g2, bn.R - 1), g2) is None                 # order r
    e = pr.pairing(g1, g2)
    assert e != pr.F12_ONE and pr.f12_pow(e, bn.R) == pr.F12_ONE          # non-degenerate, in mu_r
    a, b = 0x1234567890ABCDEF1234567, 0xFEDCBA0987654321
    assert pr.pairing(bn.g1_mul(g1, a), pr.g2_mul(g2, b)) == pr.f12_pow(e, a * b % bn.R)
    assert pr.pairing_check([(bn.g1_mul(g1, a), pr.g2_mul(g2, b)), (bn.g1_neg(bn.g1_mul(g1, a * b % bn.R)), g2)])
    assert not pr.pairing_check([(bn.g1_mul(g1, a), pr.g2_mul(g2, b)), (bn.g1_neg(bn.g1_mul(g1, a * b % bn.R + 1)), g2)])
    assert pr.pairing_check([(None, g2), (g1, None)])


@pytest.fixture(scope="module")
def reference_case():
    """The reference's own test values (src/signal.rs:93-103): k = 4, signal_hash = 5, public
    input 25."""
    srs = sp.setup(4)
    be = sp.OracleBackend(srs)
    asg = sp.Assignment(4, [5], [25])
    pk = sp.keygen(be, asg)
    proof = sp.create_proof(be, pk, asg, sp.Rng(7))
    return srs, pk, proof


def test_reference_square_circuit_proof_is_accepted_by_the_contract(reference_case):
    srs, pk, proof = reference_case
    assert len(proof) == 0x460                                            # contract.sol:221
    assert sv.verify_proof(sp.vk_code(pk, srs, 1), proof, [25])


def test_proof_matches_committed_golden(reference_case):
    _, _, proof = reference_case
    assert proof == (GOLDEN / "square_proof_k4.bin").read_bytes()


@pytest.mark.parametrize("where", ["advice commitment", "permutation product", "quotient piece", "evaluation", "W", "W'"])
def test_tampered_proof_is_rejected(reference_case, where):
    srs, pk, proof = reference_case
    vk = sp.vk_code(pk, srs, 1)
    # offsets inside the proof (calldata offset - 0x84)
    off = {"advice commitment": 0x00, "permutation product": 0x80, "quotient piece": 0x180, "evaluation": 0x200 + 0x20 * 7,
           "W": 0x3e0, "W'": 0x420}[where]
    bad = bytearray(proof)
    if where == "evaluation":
        bad[off + 31] ^= 1
    else:                      # replace the point by another curve point so the on-curve check passes
        other = bn.g1_mul(bn.G1_GEN, 0xC0FFEE + off)
        bad[off:off + 64] = bn.g1_to_evm_bytes(other)
    assert not sv.verify_proof(vk, bytes(bad), [25])
    off_curve = bytearray(proof)
    off_curve[0x3f] ^= 1                                                   # y of the first commitment: not on the curve
    assert not sv.verify_proof(vk, bytes(off_curve), [25])


def test_wrong_public_input_or_key_is_rejected(reference_case):
    srs, pk, proof = reference_case
    vk = sp.vk_code(pk, srs, 1)
    assert not sv.verify_proof(vk, proof, [26])
    assert not sv.verify_proof(vk, proof, [])                              # instance count (contract.sol:226-227)
    other_srs = sp.setup(4, seed=0xBAD5EED)
    assert not sv.verify_proof(sp.vk_code(pk, other_srs, 1), proof, [25])  # another s_g2


def test_unsatisfied_witness_is_rejected():
    """a_1 != a_0^2 on an enabled row: the numerator is not divisible by X^n - 1, so the truncated
    h(X) no longer satisfies the contract's quotient identity at x."""
    srs = sp.setup(4)
    be = sp.OracleBackend(srs)
    asg = sp.Assignment(4, [5], [25])
    pk = sp.keygen(be, asg)
    vk = sp.vk_code(pk, srs, 1)
    assert sv.verify_proof(vk, sp.create_proof(be, pk, asg, sp.Rng(7)), [25])
    assert not sv.verify_proof(vk, sp.create_proof(be, pk, asg, sp.Rng(7), witness_error=1), [25])


@pytest.mark.parametrize("k,rows", [(5, 20), (6, 58)])
def test_many_rows_and_copy_constraint(k, rows):
    """More content than the reference's single row: `rows` squares, and the copy constraint the
    reference leaves commented out (src/signal.rs:72-73: advice[1] row 0 == instance row 0), which
    makes the permutation a real one (sigma != identity)."""
    srs = sp.setup(k)
    be = sp.OracleBackend(srs)
    hashes = [3 + 11 * i for i in range(rows)]
    copies = [(("advice", 1, 0), ("instance", 0, 0)), (("advice", 0, 2), ("advice", 0, 2)),
              (("advice", 0, 1), ("instance", 0, 1))]
    asg = sp.Assignment(k, hashes, [hashes[0] ** 2, hashes[1]], copies)
    pk = sp.keygen(be, asg)
    assert pk.permutations[1][0] != pow(bn.FR_DELTA, 1, bn.R)             # sigma is not the identity any more
    proof = sp.create_proof(be, pk, asg, sp.Rng(11 + k))
    vk = sp.vk_code(pk, srs, 2)
    assert sv.verify_proof(vk, proof, asg.instances)
    assert not sv.verify_proof(vk, proof, [hashes[0] ** 2 + 1, hashes[1]])
    # a public input that violates the copy constraint cannot be proven
    broken = sp.Assignment(k, hashes, [hashes[0] ** 2 + 1, hashes[1]], copies)
    proof2 = sp.create_proof(be, pk, broken, sp.Rng(11 + k))
    assert not sv.verify_proof(vk, proof2, broken.instances)


def test_committed_golden_proof_verifies_under_committed_key():
    """Fixture only (no prover): the golden proof and key the GPU test compares against."""
    proof = (GOLDEN / "square_proof_k4.bin").read_bytes()
    vk = (GOLDEN / "square_proof_k4.vk").read_bytes()
    assert sv.verify_proof(vk, proof, [25])
    assert not sv.verify_proof(vk, proof, [24])
