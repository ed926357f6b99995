"""ORACLE (test infrastructure, never on the product path): the alt_bn128 pairing check of the
EVM precompile at address 0x08 (EIP-197), which the reference's Solidity verifier calls at
solidity_verifier_contract/contract.sol:184-202 (`ec_pairing`), and the G2 arithmetic a local KZG
setup needs for `s_g2`.

Construction (the published curve parameters; nothing here comes from the reference's sources):
    Fq2  = Fq[i] / (i^2 + 1)
    Fq12 = Fq2[w] / (w^6 - xi),  xi = 9 + i        (kept as Fq[w] / (w^12 - 18 w^6 + 82))
    E    : y^2 = x^3 + 3            over Fq   (G1, oracle/bn254.py)
    E'   : y^2 = x^3 + 3 / xi       over Fq2  (G2, D-type sextic twist; (x, y) -> (x w^2, y w^3))
    e(P, Q) = f_{6u+2, Q}(P) * l_{[6u+2]Q, pi(Q)}(P) * l_{., -pi^2(Q)}(P)  to the power (q^12 - 1) / r,
    u = 4965661367192848881  (optimal ate).

Plain affine Miller loop, full Fq12 products, final exponentiation by square-and-multiply: slow
(about a second per pairing) and easy to audit.  Pinned in tests/test_square_proof_oracle.py by
the group order of the G2 generator, bilinearity and non-degeneracy.
"""
from __future__ import annotations

from . import bn254 as bn

Q = bn.Q
R = bn.R
U = 4965661367192848881
ATE_LOOP = 6 * U + 2

# EIP-197 generator of G2: x = x_re + x_im * i (the precompile's wire order is (im, re))
G2_GEN = (
    (10857046999023057135944570762232829481370756359578518086990519993285655852781,
     11559732032986387107991004021392285783925812861821192530917403151452391805634),
    (8495653923123431417604973247489272438418190587263600148770280649306958101930,
     4082367875863433681332203403145435568316851327593401208105741076214120093531),
)


# ------------------------------------------------------------------------------ Fq2
def f2_add(a, b): return ((a[0] + b[0]) % Q, (a[1] + b[1]) % Q)
def f2_sub(a, b): return ((a[0] - b[0]) % Q, (a[1] - b[1]) % Q)
def f2_neg(a): return (-a[0] % Q, -a[1] % Q)
def f2_conj(a): return (a[0], -a[1] % Q)
def f2_mul(a, b): return ((a[0] * b[0] - a[1] * b[1]) % Q, (a[0] * b[1] + a[1] * b[0]) % Q)
def f2_scalar(a, k): return (a[0] * k % Q, a[1] * k % Q)


def f2_inv(a):
    d = pow(a[0] * a[0] + a[1] * a[1], -1, Q)
    return (a[0] * d % Q, -a[1] * d % Q)


def f2_pow(a, e):
    r = (1, 0)
    while e:
        if e & 1:
            r = f2_mul(r, a)
        a = f2_mul(a, a)
        e >>= 1
    return r


XI = (9, 1)
TWIST_B = f2_mul((3, 0), f2_inv(XI))
FROB_X = f2_pow(XI, (Q - 1) // 3)     # w^(2(q-1))
FROB_Y = f2_pow(XI, (Q - 1) // 2)     # w^(3(q-1))


# ------------------------------------------------------------------------------ G2 (affine over Fq2, None = identity)
def g2_is_on_curve(p) -> bool:
    if p is None:
        return True
    x, y = p
    return f2_mul(y, y) == f2_add(f2_mul(f2_mul(x, x), x), TWIST_B)


def g2_neg(p):
    return None if p is None else (p[0], f2_neg(p[1]))


def _g2_slope(p, q):
    if p == q:
        return f2_mul(f2_scalar(f2_mul(p[0], p[0]), 3), f2_inv(f2_scalar(p[1], 2)))
    return f2_mul(f2_sub(q[1], p[1]), f2_inv(f2_sub(q[0], p[0])))


def g2_add(p, q):
    if p is None:
        return q
    if q is None:
        return p
    if p[0] == q[0] and p[1] != q[1]:
        return None
    if p == q and p[1] == (0, 0):
        return None
    m = _g2_slope(p, q)
    x = f2_sub(f2_sub(f2_mul(m, m), p[0]), q[0])
    return (x, f2_sub(f2_mul(m, f2_sub(p[0], x)), p[1]))


def g2_mul(p, k: int):
    k %= R
    acc = None
    while k:
        if k & 1:
            acc = g2_add(acc, p)
        p = g2_add(p, p)
        k >>= 1
    return acc


def g2_frobenius(p):
    """(x, y) -> the q-power Frobenius of the untwisted point, pulled back to the twist."""
    return (f2_mul(f2_conj(p[0]), FROB_X), f2_mul(f2_conj(p[1]), FROB_Y))


# ------------------------------------------------------------------------------ Fq12 = Fq[w] / (w^12 - 18 w^6 + 82)
F12_ONE = [1] + [0] * 11


def f12_mul(a, b):
    c = [0] * 23
    for i, ai in enumerate(a):
        if ai:
            for j, bj in enumerate(b):
                c[i + j] += ai * bj
    for i in range(22, 11, -1):
        t = c[i]
        c[i - 6] += 18 * t
        c[i - 12] -= 82 * t
    return [v % Q for v in c[:12]]


def f12_pow(a, e):
    r = F12_ONE
    while e:
        if e & 1:
            r = f12_mul(r, a)
        a = f12_mul(a, a)
        e >>= 1
    return r


def _embed(v, power):
    """(a + b i) * w^power as an Fq12 element: i = w^6 - 9."""
    out = [0] * 12
    out[power] = (v[0] - 9 * v[1]) % Q
    out[power + 6] = v[1]
    return out


def _line(r_pt, slope, p):
    """The line through the (untwisted) point r_pt with (twist) slope `slope`, evaluated at the
    G1 point p:  -y_P + (slope x_P) w + (y_R - slope x_R) w^3  (up to an Fq factor, which the final
    exponentiation kills)."""
    out = [0] * 12
    out[0] = -p[1] % Q
    a = _embed(f2_scalar(slope, p[0]), 1)
    b = _embed(f2_sub(r_pt[1], f2_mul(slope, r_pt[0])), 3)
    for i in range(12):
        out[i] = (out[i] + a[i] + b[i]) % Q
    return out


def miller_loop(p, q):
    """f_{6u+2, Q}(P) with the two Frobenius lines; p in G1 (affine ints), q in G2.  Either
    identity gives 1."""
    if p is None or q is None:
        return F12_ONE
    f = F12_ONE
    r_pt = q
    for bit in bin(ATE_LOOP)[3:]:
        m = _g2_slope(r_pt, r_pt)
        f = f12_mul(f12_mul(f, f), _line(r_pt, m, p))
        r_pt = g2_add(r_pt, r_pt)
        if bit == "1":
            m = _g2_slope(r_pt, q)
            f = f12_mul(f, _line(r_pt, m, p))
            r_pt = g2_add(r_pt, q)
    q1 = g2_frobenius(q)
    nq2 = g2_neg(g2_frobenius(q1))
    m = _g2_slope(r_pt, q1)
    f = f12_mul(f, _line(r_pt, m, p))
    r_pt = g2_add(r_pt, q1)
    m = _g2_slope(r_pt, nq2)
    f = f12_mul(f, _line(r_pt, m, p))
    return f


FINAL_EXP = (Q ** 12 - 1) // R


def pairing(p, q):
    return f12_pow(miller_loop(p, q), FINAL_EXP)


def pairing_check(pairs) -> bool:
    """EIP-197: prod e(P_i, Q_i) == 1 (one shared final exponentiation)."""
    f = F12_ONE
    for p, q in pairs:
        f = f12_mul(f, miller_loop(p, q))
    return f12_pow(f, FINAL_EXP) == F12_ONE
