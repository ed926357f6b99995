"""Generate tests/golden/square_proof_k4.bin (+ .vk): the proof of the reference's own test case
(src/signal.rs:93-103: k = 4, signal_hash = 5, public input 25) assembled from the ORACLE's
hot-path functions under the fixed-seed blinding stream Rng(7) and the local SRS setup(4), and
accepted by the transliterated contract.  The CUDA path must reproduce these bytes.

    python tests/make_square_golden.py
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import square_proof as sp                      # noqa: E402
from oracle import sol_verifier as sv          # noqa: E402

if __name__ == "__main__":
    srs = sp.setup(4)
    be = sp.OracleBackend(srs)
    asg = sp.Assignment(4, [5], [25])
    pk = sp.keygen(be, asg)
    proof = sp.create_proof(be, pk, asg, sp.Rng(7))
    vk = sp.vk_code(pk, srs, 1)
    assert sv.verify_proof(vk, proof, [25])
    (ROOT / "tests" / "golden" / "square_proof_k4.bin").write_bytes(proof)
    (ROOT / "tests" / "golden" / "square_proof_k4.vk").write_bytes(vk)
    print("wrote", len(proof), "proof bytes and", len(vk), "vk bytes")
