"""oracle/sol_verifier.py is a hand transliteration of the reference's
solidity_verifier_contract/contract.sol; this test ties it to that file mechanically (CPU, in the
build container: /root/reference does not travel to the GPU box, where the test skips).

  * every `uint256 internal constant` of the contract (:6-66) exists under the same name with the
    same value; q, r (:210-211), the proof / vk lengths (:221, :307) and delta (:440) are read
    from the cited lines;
  * every block of the transliteration cites the contract lines it restates, and every memory /
    calldata pointer literal and every named pointer it uses occurs inside exactly those lines;
  * inside the quotient identity (:439-511) the calldata offsets of the evaluations are read in
    the contract's order, and inside the transcript section the proof is absorbed in the
    contract's section sizes.
"""
import re
from pathlib import Path

import pytest

from oracle import sol_verifier as sv

CONTRACT = Path("/root/reference/solidity_verifier_contract/contract.sol")
pytestmark = pytest.mark.skipif(not CONTRACT.exists(), reason="the reference tree is only present in the build container")

PTR = r"\b[A-Z][A-Z0-9_]*_(?:MPTR|CPTR)\b"


def _sol():
    return CONTRACT.read_text().split("\n")


def _py():
    return Path(sv.__file__).read_text().split("\n")


def test_named_constants_equal_the_contract():
    sol = _sol()
    decl = {}
    for ln in sol[5:66]:                                         # contract.sol:6-66
        m = re.match(r"\s*uint256 internal constant\s+(\w+)\s*=\s*(0x[0-9a-fA-F]+);", ln)
        if m:
            decl[m.group(1)] = int(m.group(2), 16)
    assert len(decl) == 55
    for name, value in decl.items():
        assert getattr(sv, name) == value, name
    assert int(re.search(r"let q := (\d+)", sol[209]).group(1)) == sv.Q          # :210
    assert int(re.search(r"let r := (\d+)", sol[210]).group(1)) == sv.R          # :211
    assert int(re.search(r"eq\((0x[0-9a-f]+), calldataload\(PROOF_LEN_CPTR\)\)", sol[220]).group(1), 16) == sv.PROOF_LEN   # :221
    assert int(re.search(r"extcodecopy\(vk, VK_MPTR, 0x00, (0x[0-9a-f]+)\)", sol[306]).group(1), 16) == sv.VK_LEN          # :307
    assert int(re.search(r"let delta := (\d+)", sol[439]).group(1)) == sv.DELTA  # :440


def _blocks():
    """(python line number, code, cited (first, last) contract lines) for every code line of
    verify_proof; a `contract.sol:A-B` comment opens a block, a bare `:A-B` comment narrows it."""
    py = _py()
    start = next(i for i, ln in enumerate(py) if ln.startswith("def verify_proof"))
    cur = None
    for i, ln in enumerate(py[start:], start + 1):
        code, _, com = ln.partition("#")
        m = re.search(r"contract\.sol:(\d+)(?:-(\d+))?", com) or re.search(r"(?<![\w.]):(\d+)(?:-(\d+))?", com)
        if m:
            cur = (int(m.group(1)), int(m.group(2) or m.group(1)))
        if cur and code.strip():
            yield i, code, cur


def test_every_pointer_of_the_transliteration_occurs_in_the_cited_lines():
    sol = _sol()
    checked = 0
    for i, code, (a, b) in _blocks():
        assert 1 <= a <= b <= len(sol), (i, a, b)
        text = "\n".join(sol[a - 1:b])
        values = {int(h, 16) for h in re.findall(r"0x[0-9a-fA-F]+", text)}
        values |= {int(d) for d in re.findall(r"(?<![0-9a-zA-Z_x])\d+(?![0-9a-zA-Z_x])", text)}
        names = set(re.findall(PTR, text))
        for h in re.findall(r"0x[0-9a-fA-F]+", code):
            checked += 1
            assert int(h, 16) in values, f"sol_verifier.py:{i}: {h} not in contract.sol:{a}-{b}"
        for nm in re.findall(PTR, code):
            checked += 1
            assert nm in names, f"sol_verifier.py:{i}: {nm} not in contract.sol:{a}-{b}"
    assert checked >= 400


def _section(first, last):
    return [(i, code) for i, code, (a, b) in _blocks() if first <= a and b <= last]


def test_quotient_identity_reads_the_evaluations_in_the_contract_order():
    sol = _sol()
    want = re.findall(r"calldataload\((0x[0-9a-f]+)\)", "\n".join(sol[438:511]))          # :439-511
    got = [h for _, code in _section(439, 511) for h in re.findall(r"calldataload\((0x[0-9a-f]+)\)", code)]
    assert [int(h, 16) for h in got] == [int(h, 16) for h in want] and len(want) >= 20
    # ... and folds with y exactly as often
    folds_sol = len(re.findall(r"quotient_eval_numer := addmod\(mulmod\(quotient_eval_numer, y, r\)", "\n".join(sol[438:511])))
    folds_py = sum(code.count("addmod(mulmod(quotient_eval_numer, y, r)") for _, code in _section(439, 511))
    assert folds_sol == folds_py == 7


def test_transcript_absorbs_the_proof_in_the_contract_sections():
    sol = _sol()
    want = re.findall(r"let proof_cptr_end := add\(proof_cptr, (0x[0-9a-f]+)\)", "\n".join(sol[215:352]))
    got = [h for _, code in _section(216, 352) for h in re.findall(r"proof_cptr_end = proof_cptr \+ (0x[0-9a-f]+)", code)]
    assert [int(h, 16) for h in got] == [int(h, 16) for h in want] == [0x80, 0x100, 0x80, 0x1e0]
    squeezes_sol = len(re.findall(r":= squeeze_challenge(?:_cont)?\(", "\n".join(sol[215:352])))
    squeezes_py = sum(len(re.findall(r"= squeeze_challenge(?:_cont)?\(", code)) for _, code in _section(216, 352))
    assert squeezes_sol == squeezes_py == 8
