"""BN254 field / curve model in Python big integers.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (the C-ABI library under
``anon-aadhaar-halo2_b200/``) may import or execute this module; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs use it, and only as the checker.

PARITY PIN.  Per-function output vectors: none exist — the reference repository holds no
golden vector, known-answer test or call site for the MSM / NTT / quotient path (SURVEY.md
F1, F2, section 4), so at that level parity is unpinned and the golden files are generated
from the definitions.  End to end the oracle IS pinned by the one fixture the reference holds
for this path, its Solidity verifier: complete SquareCircuit proofs assembled from the oracle's
functions are accepted by the statement-by-statement transliteration of
``solidity_verifier_contract/contract.sol`` (``oracle/sol_verifier.py``,
``tests/test_square_proof_oracle.py``), and tampered ones are rejected.  The
arithmetic lives in un-vendored git dependencies:

  * halo2curves 0.3.1, tag ``0.3.1`` rev 9b67e19b (``Cargo.lock:484-486``):
    ``src/bn256/fr.rs``, ``src/bn256/fq.rs``, ``src/bn256/curve.rs``
  * halo2_proofs 0.2.0, tag ``v2023_01_20`` rev c7e42e41 (``Cargo.lock:469-471``)

What the reference *does* pin, and what this file is checked against in
``tests/test_oracle_constants.py``:

  * q, r                      ``solidity_verifier_contract/contract.sol:210-211``
  * curve  y^2 = x^3 + 3      ``solidity_verifier_contract/contract.sol:82``
  * Fr::DELTA = 7^(2^28)      ``solidity_verifier_contract/contract.sol:440``

All values here are *canonical* integers unless a name says ``mont``.  The wire
layout of the halo2curves types (what crosses the C ABI) is 4 little-endian u64
limbs holding the Montgomery form x*2^256 mod p.
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------- moduli
Q = 21888242871839275222246405745257275088696311157297823662689037894645226208583  # Fq
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617  # Fr
B = 3            # curve constant
G1_GEN = (1, 2)  # halo2curves bn256 generator

MONT_BITS = 256
MONT_R = 1 << MONT_BITS

FR_S = 28                      # two-adicity of r-1
FR_GENERATOR = 7               # multiplicative generator used by halo2curves
FR_T = (R - 1) >> FR_S
FR_ROOT_OF_UNITY = pow(FR_GENERATOR, FR_T, R)          # order exactly 2^28
FR_DELTA = pow(FR_GENERATOR, 1 << FR_S, R)             # contract.sol:440
# halo2curves Fr::ZETA: primitive cube root of unity
FR_ZETA = 0x30644E72E131A029048B6E193FD84104CC37A73FEC2BC5E9B8CA0B2D36636F23


def mont_consts(p: int) -> dict:
    """Montgomery constants for modulus p (R = 2^256)."""
    return {
        "R": MONT_R % p,
        "R2": (MONT_R * MONT_R) % p,
        "INV64": (-pow(p, -1, 1 << 64)) % (1 << 64),
        "INV32": (-pow(p, -1, 1 << 32)) % (1 << 32),
    }


FR_MONT = mont_consts(R)
FQ_MONT = mont_consts(Q)
_R_INV = {R: pow(MONT_R, -1, R), Q: pow(MONT_R, -1, Q)}


def to_mont(x: int, p: int) -> int:
    return (x * MONT_R) % p


def from_mont(x: int, p: int) -> int:
    return (x * _R_INV[p]) % p


# ------------------------------------------------------------------ limb (wire) layout
def int_to_limbs(x: int) -> np.ndarray:
    """256-bit integer -> 4 little-endian u64 limbs."""
    return np.frombuffer(int(x).to_bytes(32, "little"), dtype="<u8").copy()


def limbs_to_int(l) -> int:
    return int.from_bytes(np.asarray(l, dtype="<u8").tobytes(), "little")


def ints_to_array(xs, width: int = 1) -> np.ndarray:
    """List of 256-bit ints -> (len/width, 4*width) u64 array (wire layout)."""
    buf = b"".join(int(x).to_bytes(32, "little") for x in xs)
    a = np.frombuffer(buf, dtype="<u8").copy()
    return a.reshape(-1, 4 * width)


def array_to_ints(a) -> list:
    b = np.ascontiguousarray(a, dtype="<u8").tobytes()
    return [int.from_bytes(b[i:i + 32], "little") for i in range(0, len(b), 32)]


def fr_array_from_canonical(xs) -> np.ndarray:
    """Canonical Fr ints -> (n,4) u64 Montgomery limbs (what halo2curves stores)."""
    return ints_to_array([to_mont(x, R) for x in xs])


def fr_array_to_canonical(a) -> list:
    return [from_mont(x, R) for x in array_to_ints(a)]


def g1_affine_array_from_points(pts) -> np.ndarray:
    """[(x,y) | None] -> (n,8) u64: Montgomery x limbs then y limbs; identity = zeros
    (halo2curves G1Affine identity is (0,0))."""
    flat = []
    for p in pts:
        if p is None:
            flat += [0, 0]
        else:
            flat += [to_mont(p[0], Q), to_mont(p[1], Q)]
    return ints_to_array(flat, width=2)


def g1_affine_array_to_points(a) -> list:
    v = array_to_ints(a)
    out = []
    for i in range(0, len(v), 2):
        x, y = from_mont(v[i], Q), from_mont(v[i + 1], Q)
        out.append(None if (x == 0 and y == 0) else (x, y))
    return out


def g1_jacobian_limbs_to_affine(limbs12):
    """12 u64 limbs (Montgomery X,Y,Z Jacobian, as halo2curves ``G1``) -> affine or None."""
    X, Y, Z = (from_mont(v, Q) for v in array_to_ints(limbs12))
    if Z == 0:
        return None
    zi = pow(Z, -1, Q)
    return (X * zi * zi % Q, Y * zi * zi * zi % Q)


# ----------------------------------------------------------------------------- curve
def g1_is_on_curve(p) -> bool:
    if p is None:
        return True
    x, y = p
    return (y * y - x * x * x - B) % Q == 0


def g1_neg(p):
    return None if p is None else (p[0], (-p[1]) % Q)


def g1_add(p, q):
    """Affine chord-and-tangent addition; None is the identity."""
    if p is None:
        return q
    if q is None:
        return p
    x1, y1 = p
    x2, y2 = q
    if x1 == x2:
        if (y1 + y2) % Q == 0:
            return None
        lam = 3 * x1 * x1 * pow(2 * y1, -1, Q) % Q
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, Q) % Q
    x3 = (lam * lam - x1 - x2) % Q
    return (x3, (lam * (x1 - x3) - y1) % Q)


# Jacobian arithmetic for speed in the Python model (no inversion per add).
def _jac_double(P):
    X, Y, Z = P
    if Z == 0:
        return P
    A = X * X % Q
    Bq = Y * Y % Q
    C = Bq * Bq % Q
    D = 2 * ((X + Bq) * (X + Bq) - A - C) % Q
    E = 3 * A % Q
    F = E * E % Q
    X3 = (F - 2 * D) % Q
    Y3 = (E * (D - X3) - 8 * C) % Q
    Z3 = 2 * Y * Z % Q
    return (X3, Y3, Z3)


def _jac_add_affine(P, q):
    if q is None:
        return P
    X1, Y1, Z1 = P
    if Z1 == 0:
        return (q[0], q[1], 1)
    Z1Z1 = Z1 * Z1 % Q
    U2 = q[0] * Z1Z1 % Q
    S2 = q[1] * Z1 * Z1Z1 % Q
    if U2 == X1:
        if S2 == Y1:
            return _jac_double(P)
        return (0, 1, 0)
    H = (U2 - X1) % Q
    HH = H * H % Q
    I = 4 * HH % Q
    J = H * I % Q
    r = 2 * (S2 - Y1) % Q
    V = X1 * I % Q
    X3 = (r * r - J - 2 * V) % Q
    Y3 = (r * (V - X3) - 2 * Y1 * J) % Q
    Z3 = ((Z1 + H) * (Z1 + H) - Z1Z1 - HH) % Q
    return (X3, Y3, Z3)


def _jac_to_affine(P):
    X, Y, Z = P
    if Z == 0:
        return None
    zi = pow(Z, -1, Q)
    return (X * zi * zi % Q, Y * zi * zi * zi % Q)


def g1_mul(p, k: int):
    """[k]p by left-to-right double-and-add (k taken mod r)."""
    k %= R
    if p is None or k == 0:
        return None
    acc = (0, 1, 0)
    for bit in bin(k)[2:]:
        acc = _jac_double(acc)
        if bit == "1":
            acc = _jac_add_affine(acc, p)
    return _jac_to_affine(acc)


def g1_msm_naive(scalars, points):
    """sum_i scalars[i] * points[i], the *definition* of best_multiexp's result."""
    acc = (0, 1, 0)
    for s, p in zip(scalars, points):
        t = g1_mul(p, s)
        acc = _jac_add_affine(acc, t)
    return _jac_to_affine(acc)


# --------------------------------------------------------------------- seeded inputs
# Counter-based generator shared by oracle (Python + C) and the CUDA input generator so
# that 2^26-sized synthetic inputs can be produced in HBM without a host copy.
_M64 = (1 << 64) - 1


def splitmix64(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & _M64
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return z ^ (z >> 31)


def _splitmix64_np(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15))
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def seeded_fr_mont_limbs(seed: int, n: int, start: int = 0) -> np.ndarray:
    """(n,4) u64 array: element i has limbs splitmix64(seed*2^32 + 4*i + j), top limb
    masked to 62 bits (value < 2^254), one conditional subtraction of r.  The result
    is *defined* to be the Montgomery representation of scalar i (a uniform Montgomery
    representative is a uniform field element)."""
    with np.errstate(over="ignore"):
        idx = (np.arange(start, start + n, dtype=np.uint64)[:, None] * np.uint64(4)
               + np.arange(4, dtype=np.uint64)[None, :])
        base = np.uint64((seed << 32) & _M64)
        limbs = _splitmix64_np(base + idx)
    limbs[:, 3] &= np.uint64((1 << 62) - 1)
    # conditional subtract r (vectorised through Python ints only where needed)
    r_limbs = int_to_limbs(R)
    ge = np.zeros(n, dtype=bool)
    # lexicographic compare from the top limb
    undecided = np.ones(n, dtype=bool)
    for j in (3, 2, 1, 0):
        gt = undecided & (limbs[:, j] > r_limbs[j])
        lt = undecided & (limbs[:, j] < r_limbs[j])
        ge |= gt
        undecided &= ~(gt | lt)
    ge |= undecided  # equal
    for i in np.nonzero(ge)[0]:
        limbs[i] = int_to_limbs(limbs_to_int(limbs[i]) - R)
    return limbs


def seeded_point_scalar(seed: int, i: int) -> int:
    """64-bit multiplier t_i (never 0) such that base point i = [t_i] G."""
    t = splitmix64(((seed << 32) + i) & _M64)
    return t | 1


def seeded_g1_points(seed: int, n: int, start: int = 0) -> list:
    """Affine points [t_i]G, t_i = seeded_point_scalar(seed, i)."""
    # fixed-base table of 2^j G
    tbl = []
    p = (G1_GEN[0], G1_GEN[1], 1)
    for _ in range(64):
        tbl.append(_jac_to_affine(p))
        p = _jac_double(p)
    out = []
    for i in range(start, start + n):
        t = seeded_point_scalar(seed, i)
        acc = (0, 1, 0)
        j = 0
        while t:
            if t & 1:
                acc = _jac_add_affine(acc, tbl[j])
            t >>= 1
            j += 1
        out.append(_jac_to_affine(acc))
    return out


# ----------------------------------------------------------------------- wire formats
# [DEP] halo2curves 0.3.1 src/derive/curve.rs `new_curve_impl!` (reference Cargo.lock:484-486):
# to_bytes = x little-endian, bit 7 of byte 31 = parity of the canonical y, identity = zeros.
def g1_to_bytes(p) -> bytes:
    if p is None:
        return bytes(32)
    b = bytearray(p[0].to_bytes(32, "little"))
    b[31] |= (p[1] & 1) << 7
    return bytes(b)


def g1_from_bytes(b: bytes):
    """Inverse of g1_to_bytes; raises ValueError for a non-canonical x or a non-residue."""
    b = bytearray(b)
    sign = b[31] >> 7
    b[31] &= 0x7F
    x = int.from_bytes(b, "little")
    if x == 0 and sign == 0:
        return None
    if x >= Q:
        raise ValueError("x coordinate is not canonical")
    rhs = (x * x * x + 3) % Q
    y = pow(rhs, (Q + 1) // 4, Q)
    if y * y % Q != rhs:
        raise ValueError("not on the curve")
    if (y & 1) != sign:
        y = Q - y
    return (x, y)


def g1_to_evm_bytes(p) -> bytes:
    """reference solidity_verifier_contract/contract.sol:77-87: x then y, 32-byte big-endian."""
    if p is None:
        return bytes(64)
    return p[0].to_bytes(32, "big") + p[1].to_bytes(32, "big")
