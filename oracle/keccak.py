"""ORACLE (test infrastructure, never on the product path): Keccak-256 as the EVM's KECCAK256
opcode computes it (original Keccak padding 0x01 ... 0x80, rate 136 bytes) — Python's hashlib
only has the NIST SHA-3 padding.  Used by the transliteration of the reference's Solidity
verifier (solidity_verifier_contract/contract.sol:93, :105, :793 hash with `keccak256`) and by the
EVM transcript of the SquareCircuit proof harness.

Pinned in tests/test_square_proof_oracle.py by the published digests of "" and "abc".
"""
from __future__ import annotations

_MASK = (1 << 64) - 1
_RC = []
_ROT = [[0] * 5 for _ in range(5)]


def _init():
    # round constants from the degree-8 LFSR of the Keccak specification
    r = 1
    for _ in range(24):
        rc = 0
        for j in range(7):
            if r & 1:
                rc ^= 1 << ((1 << j) - 1)
            r = ((r << 1) ^ (0x71 if r & 0x80 else 0)) & 0xFF
        _RC.append(rc)
    # rotation offsets: (t + 1)(t + 2) / 2 along the walk (x, y) -> (y, 2x + 3y)
    x, y = 1, 0
    for t in range(24):
        _ROT[x][y] = ((t + 1) * (t + 2) // 2) % 64
        x, y = y, (2 * x + 3 * y) % 5


_init()


def _rol(v, n):
    return ((v << n) | (v >> (64 - n))) & _MASK if n else v


def _f1600(a):
    for rc in _RC:
        c = [a[x][0] ^ a[x][1] ^ a[x][2] ^ a[x][3] ^ a[x][4] for x in range(5)]
        d = [c[(x - 1) % 5] ^ _rol(c[(x + 1) % 5], 1) for x in range(5)]
        a = [[a[x][y] ^ d[x] for y in range(5)] for x in range(5)]
        b = [[0] * 5 for _ in range(5)]
        for x in range(5):
            for y in range(5):
                b[y][(2 * x + 3 * y) % 5] = _rol(a[x][y], _ROT[x][y])
        a = [[b[x][y] ^ ((~b[(x + 1) % 5][y]) & b[(x + 2) % 5][y]) for y in range(5)] for x in range(5)]
        a[0][0] ^= rc
    return a


def keccak256(data: bytes) -> bytes:
    rate = 136
    msg = bytearray(data)
    pad = rate - len(msg) % rate
    msg += b"\x01" + b"\x00" * (pad - 1) if pad > 1 else b""
    if pad == 1:
        msg += b"\x81"
    else:
        msg[-1] |= 0x80
    a = [[0] * 5 for _ in range(5)]
    for off in range(0, len(msg), rate):
        block = msg[off:off + rate]
        for i in range(rate // 8):
            a[i % 5][i // 5] ^= int.from_bytes(block[8 * i:8 * i + 8], "little")
        a = _f1600(a)
    out = b"".join(a[i % 5][i // 5].to_bytes(8, "little") for i in range(4))
    return out
