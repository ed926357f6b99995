"""GPU: a complete proof of the reference's SquareCircuit assembled from the product's outputs
(every commitment, transform, quotient evaluation, grand product, polynomial evaluation and
division goes through the C ABI) is (a) byte-identical to the proof assembled from the oracle's
outputs under the same fixed-seed blinding stream and to the committed golden proof, and (b)
accepted by the transliteration of the reference's Solidity verifier."""
import pytest

import square_proof as sp
from oracle import sol_verifier as sv
from util import GOLDEN

pytestmark = pytest.mark.gpu


def _prove(be, asg, seed):
    pk = sp.keygen(be, asg)
    return pk, sp.create_proof(be, pk, asg, sp.Rng(seed))


def test_reference_case_matches_golden_and_verifies(zk):
    """src/signal.rs:93-103: k = 4, signal_hash = 5, public input 25."""
    srs = sp.setup(4)
    dev = sp.DeviceBackend(srs, zk)
    try:
        pk, proof = _prove(dev, sp.Assignment(4, [5], [25]), 7)
    finally:
        dev.close()
    assert proof == (GOLDEN / "square_proof_k4.bin").read_bytes()
    assert sp.vk_code(pk, srs, 1) == (GOLDEN / "square_proof_k4.vk").read_bytes()
    assert sv.verify_proof(sp.vk_code(pk, srs, 1), proof, [25])


@pytest.mark.parametrize("k,rows", [(5, 20), (7, 100), (10, 1000)])
def test_device_proof_is_byte_identical_to_oracle_proof(zk, k, rows):
    srs = sp.setup(k)
    hashes = [3 + 11 * i for i in range(rows)]
    copies = [(("advice", 1, 0), ("instance", 0, 0)), (("advice", 0, 1), ("instance", 0, 1))]
    asg = sp.Assignment(k, hashes, [hashes[0] ** 2, hashes[1]], copies)
    dev = sp.DeviceBackend(srs, zk)
    try:
        pk_d, proof_d = _prove(dev, asg, 100 + k)
    finally:
        dev.close()
    pk_o, proof_o = _prove(sp.OracleBackend(srs), asg, 100 + k)
    assert sp.vk_code(pk_d, srs, 2) == sp.vk_code(pk_o, srs, 2)            # keygen commitments
    assert proof_d == proof_o
    assert sv.verify_proof(sp.vk_code(pk_d, srs, 2), proof_d, asg.instances)


def test_prover_shape_domain_proof_verifies(zk):
    """k = 15, the RSA-SHA256 circuit's domain size (src/lib.rs:444): every usable row holds a
    square; the device proof is accepted by the contract and a wrong public input is not."""
    k = 15
    srs = sp.setup(k)
    n_rows = (1 << k) - 6
    hashes = [(0x9E3779B97F4A7C15 * (i + 1)) % sp.R for i in range(n_rows)]
    copies = [(("advice", 1, 0), ("instance", 0, 0))]
    asg = sp.Assignment(k, hashes, [hashes[0] ** 2 % sp.R], copies)
    dev = sp.DeviceBackend(srs, zk)
    try:
        pk, proof = _prove(dev, asg, 15)
    finally:
        dev.close()
    vk = sp.vk_code(pk, srs, 1)
    assert sv.verify_proof(vk, proof, asg.instances)
    assert not sv.verify_proof(vk, proof, [asg.instances[0] + 1])
