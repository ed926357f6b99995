"""Generate tests/golden/*.npz from the Python big-integer oracle (mathematical
definitions: naive DFT, Horner evaluation on the coset, naive sum of scalar multiples).

    python -m oracle.make_golden

The reference repository holds no vectors for this path, so these
known-answer files are produced here and committed; every value is derived from the
definitions in oracle/bn254.py + oracle/halo2_cpu.py, not from the CUDA code.
"""
from __future__ import annotations

import random
from pathlib import Path

import numpy as np

from . import bn254 as bn
from . import halo2_cpu as h

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def omega_for(k: int) -> int:
    w = bn.FR_ROOT_OF_UNITY
    for _ in range(k, bn.FR_S):
        w = w * w % bn.R
    return w


def main() -> None:
    OUT.mkdir(parents=True, exist_ok=True)
    rnd = random.Random(0xB200)

    # ---- NTT known answers by the O(n^2) definition
    ntt = {}
    for k in range(1, 9):
        n = 1 << k
        a = [rnd.randrange(bn.R) for _ in range(n)]
        if k == 3:
            a = [1] + [0] * (n - 1)            # impulse -> all ones
        if k == 4:
            a = [1] * n                         # all ones -> n * impulse
        if k == 5:
            a[0], a[1], a[2] = 0, bn.R - 1, 1
        w = omega_for(k)
        ntt[f"in_{k}"] = bn.fr_array_from_canonical(a)
        ntt[f"omega_{k}"] = bn.fr_array_from_canonical([w])[0]
        ntt[f"out_{k}"] = bn.fr_array_from_canonical(h.dft_naive(a, w))
    np.savez_compressed(OUT / "ntt_kat.npz", **ntt)

    # ---- EvaluationDomain transforms by definition (Horner on the coset, inverse by
    # interpolation identity checked through the forward definition)
    dom = {}
    cases = [(3, 3), (4, 4), (5, 4), (3, 6)]
    dom["cases"] = np.array(cases, dtype=np.int64)
    for (j, k) in cases:
        d = h.EvaluationDomain(j, k)
        coeffs = [rnd.randrange(bn.R) for _ in range(d.n)]
        tag = f"{j}_{k}"
        lag = h.dft_naive(coeffs, d.omega)                       # evaluations on H
        ext = d.eval_coeff_on_coset(coeffs)                      # definition of coeff_to_extended
        dom[f"coeff_{tag}"] = bn.fr_array_from_canonical(coeffs)
        dom[f"lagrange_{tag}"] = bn.fr_array_from_canonical(lag)
        dom[f"extended_{tag}"] = bn.fr_array_from_canonical(ext)
        # extended_to_coeff on a random degree < n*(j-1) polynomial's coset evaluations
        big = [rnd.randrange(bn.R) for _ in range(d.n * d.quotient_poly_degree)]
        big_ext = []
        for i in range(d.extended_n):
            x = bn.FR_ZETA * pow(d.extended_omega, i, bn.R) % bn.R
            acc = 0
            for c in reversed(big):
                acc = (acc * x + c) % bn.R
            big_ext.append(acc)
        dom[f"bigcoeff_{tag}"] = bn.fr_array_from_canonical(big)
        dom[f"bigext_{tag}"] = bn.fr_array_from_canonical(big_ext)
        # vanishing division: h[i] / (x_i^n - 1)
        hv = [rnd.randrange(bn.R) for _ in range(d.extended_n)]
        div = []
        for i, v in enumerate(hv):
            x = bn.FR_ZETA * pow(d.extended_omega, i, bn.R) % bn.R
            div.append(v * pow((pow(x, d.n, bn.R) - 1) % bn.R, -1, bn.R) % bn.R)
        dom[f"h_{tag}"] = bn.fr_array_from_canonical(hv)
        dom[f"hdiv_{tag}"] = bn.fr_array_from_canonical(div)
    np.savez_compressed(OUT / "domain_kat.npz", **dom)

    # ---- MSM known answers by the definition sum_i [s_i] P_i
    msm = {}
    names = []

    def add_case(name, scalars, points):
        names.append(name)
        msm[f"s_{name}"] = bn.fr_array_from_canonical(scalars)
        msm[f"p_{name}"] = bn.g1_affine_array_from_points(points)
        res = bn.g1_msm_naive(scalars, points)
        msm[f"r_{name}"] = bn.g1_affine_array_from_points([res])[0]

    pts = bn.seeded_g1_points(0xBA5E0001, 96)
    add_case("single", [rnd.randrange(bn.R)], pts[:1])
    add_case("three", [rnd.randrange(bn.R) for _ in range(3)], pts[:3])
    add_case("rand40", [rnd.randrange(bn.R) for _ in range(40)], pts[:40])
    add_case("rand96", [rnd.randrange(bn.R) for _ in range(96)], pts[:96])
    add_case("edge_scalars", [0, 1, bn.R - 1, 2, bn.R - 2, (bn.R - 1) // 2, 1 << 253, (1 << 128) - 1], pts[:8])
    add_case("all_rm1", [bn.R - 1] * 20, pts[:20])
    add_case("repeated_point", [rnd.randrange(bn.R) for _ in range(16)], [pts[5]] * 16)
    s = rnd.randrange(bn.R)
    add_case("cancel", [s, s, 7], [pts[1], bn.g1_neg(pts[1]), pts[2]])
    add_case("to_identity", [s, s], [pts[1], bn.g1_neg(pts[1])])
    add_case("identity_bases", [rnd.randrange(bn.R) for _ in range(6)], [pts[0], None, pts[2], None, None, pts[3]])
    add_case("small_scalars", [i % 5 for i in range(64)], pts[:64])
    msm["names"] = np.array(names)
    np.savez_compressed(OUT / "msm_kat.npz", **msm)
    prover_steps()
    print("wrote", sorted(p.name for p in OUT.glob("*.npz")))


def prover_steps() -> None:
    """Known answers for the prover steps either side of the hot path (SURVEY.md section 8 f),
    from the DEFINITIONS (modular inverses by pow(x, -1), polynomial identities, products of
    fractions), not from oracle/prover_steps_cpu.py, which the tests then check against these."""
    OUT.mkdir(parents=True, exist_ok=True)
    rnd = random.Random(0xF1F2)
    g = {}
    R = bn.R
    # batch_invert
    a = [rnd.randrange(R) for _ in range(24)]
    a[5] = 0
    g["inv_in"] = bn.fr_array_from_canonical(a)
    g["inv_out"] = bn.fr_array_from_canonical([pow(x, -1, R) if x else 0 for x in a])
    # eval_polynomial / kate_division: q(X) (X - b) + a(b) == a(X)
    poly = [rnd.randrange(R) for _ in range(37)]
    x = rnd.randrange(R)
    g["eval_poly"] = bn.fr_array_from_canonical(poly)
    g["eval_point"] = bn.fr_array_from_canonical([x])[0]
    g["eval_out"] = bn.fr_array_from_canonical([sum(c * pow(x, i, R) for i, c in enumerate(poly)) % R])[0]
    b = rnd.randrange(R)
    q = [0] * (len(poly) - 1)                    # schoolbook division by (X - b), from the top
    rem = list(poly)
    for i in range(len(poly) - 1, 0, -1):
        q[i - 1] = rem[i]
        rem[i - 1] = (rem[i - 1] + b * rem[i]) % R
    g["kate_b"] = bn.fr_array_from_canonical([b])[0]
    g["kate_out"] = bn.fr_array_from_canonical(q)
    # permutation grand products by definition: z_s[i] = z_s[0] * prod_{r < i} num_s(r) / den_s(r)
    k, n_cols, chunk, bf = 4, 5, 2, 3
    n = 1 << k
    omega = omega_for(k)
    vals = [[rnd.randrange(R) for _ in range(n)] for _ in range(n_cols)]
    sig = [[rnd.randrange(R) for _ in range(n)] for _ in range(n_cols)]
    beta, gamma = rnd.randrange(R), rnd.randrange(R)
    zs, first = [], 1
    for lo in range(0, n_cols, chunk):
        cols = range(lo, min(lo + chunk, n_cols))
        z = [first]
        for r in range(n - 1):
            num = den = 1
            for j in cols:
                num = num * (vals[j][r] + pow(bn.FR_DELTA, j, R) * pow(omega, r, R) * beta + gamma) % R
                den = den * (vals[j][r] + beta * sig[j][r] + gamma) % R
            z.append(z[-1] * num * pow(den, -1, R) % R)
        first = z[n - (bf + 1)]
        zs.append(z)
    g["perm_shape"] = np.array([k, n_cols, chunk, bf], dtype=np.int64)
    g["perm_values"] = np.stack([bn.fr_array_from_canonical(v) for v in vals])
    g["perm_sigma"] = np.stack([bn.fr_array_from_canonical(v) for v in sig])
    g["perm_beta_gamma"] = bn.fr_array_from_canonical([beta, gamma])
    g["perm_z"] = np.stack([bn.fr_array_from_canonical(z) for z in zs])
    # lookup grand product by definition
    ci, ct, pi, pt = ([rnd.randrange(R) for _ in range(n)] for _ in range(4))
    z = [1]
    for r in range(n - 1):
        num = (ci[r] + beta) * (ct[r] + gamma) % R
        den = (pi[r] + beta) * (pt[r] + gamma) % R
        z.append(z[-1] * num * pow(den, -1, R) % R)
    g["lookup_cols"] = np.stack([bn.fr_array_from_canonical(v) for v in (ci, ct, pi, pt)])
    g["lookup_z"] = bn.fr_array_from_canonical(z)
    # G1 wire formats: identity, G, -G, 2G (the public EIP-196 value), a random multiple
    pts = [None, (1, 2), (1, bn.Q - 2), bn.g1_mul((1, 2), 2), bn.g1_mul((1, 2), rnd.randrange(R))]
    g["enc_points"] = bn.g1_affine_array_from_points(pts)
    g["enc_bytes"] = np.frombuffer(b"".join(bn.g1_to_bytes(p) for p in pts), dtype=np.uint8).reshape(len(pts), 32)
    g["enc_evm"] = np.frombuffer(b"".join(bn.g1_to_evm_bytes(p) for p in pts), dtype=np.uint8).reshape(len(pts), 64)
    np.savez_compressed(OUT / "prover_steps_kat.npz", **g)


if __name__ == "__main__":
    main()
