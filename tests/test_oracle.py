"""The oracle against what pins it: the reference's constants, the mathematical
definitions (golden files made by oracle/make_golden.py) and itself (Python <-> C)."""
import random

import numpy as np
import pytest

from oracle import bn254 as bn
from oracle import c_oracle as co
from oracle import halo2_cpu as h
from util import GOLDEN, fr1, jac_affine, omega_for


def test_constants_pinned_by_reference_contract():
    # reference solidity_verifier_contract/contract.sol:210-211, :440, :82
    assert bn.Q == 21888242871839275222246405745257275088696311157297823662689037894645226208583
    assert bn.R == 21888242871839275222246405745257275088548364400416034343698204186575808495617
    assert bn.FR_DELTA == 4131629893567559867359510883348571134090853742863529169391034518566172092834
    assert bn.g1_is_on_curve(bn.G1_GEN) and bn.B == 3
    assert pow(bn.FR_ROOT_OF_UNITY, 1 << 28, bn.R) == 1 and pow(bn.FR_ROOT_OF_UNITY, 1 << 27, bn.R) != 1
    assert pow(bn.FR_ZETA, 3, bn.R) == 1 and bn.FR_ZETA != 1
    # SURVEY.md App. A values
    assert bn.FR_ROOT_OF_UNITY == 0x03DDB9F5166D18B798865EA93DD31F743215CF6DD39329C8D34F1ED960C37C9C
    assert omega_for(15) == 0x2B7DDFE4383C8D806530B94D3120CE6FCB511871E4D44A65F0ACD0B96A8A942E
    assert omega_for(4) == 0x21082CA216CBBF4E1C6E4F4594DD508C996DFBE1174EFB98B11509C6E306460B
    assert bn.FR_MONT["INV64"] == 0xC2E1F593EFFFFFFF and bn.FQ_MONT["INV64"] == 0x87D20782E4866389


def test_best_fft_python_matches_definition_golden():
    z = np.load(GOLDEN / "ntt_kat.npz")
    for k in range(1, 9):
        a = bn.fr_array_to_canonical(z[f"in_{k}"])
        w = bn.fr_array_to_canonical(z[f"omega_{k}"][None, :])[0]
        assert w == omega_for(k)
        assert bn.fr_array_from_canonical(h.best_fft(a, w, k)).tolist() == z[f"out_{k}"].tolist()


def test_best_fft_c_matches_golden_and_python():
    z = np.load(GOLDEN / "ntt_kat.npz")
    for k in range(1, 9):
        for threads in (1, 4):
            got = co.best_fft(z[f"in_{k}"], z[f"omega_{k}"], k, threads)
            assert np.array_equal(got, z[f"out_{k}"])
    for k in (10, 13):
        a = co.gen_scalars(40 + k, 1 << k)
        exp = h.best_fft(bn.fr_array_to_canonical(a), omega_for(k), k)
        assert bn.fr_array_to_canonical(co.best_fft(a, fr1(omega_for(k)), k)) == exp


def test_domain_transforms_match_definition_golden():
    z = np.load(GOLDEN / "domain_kat.npz")
    for (j, k) in z["cases"].tolist():
        tag = f"{j}_{k}"
        d = h.EvaluationDomain(j, k)
        coeff = bn.fr_array_to_canonical(z[f"coeff_{tag}"])
        lag = bn.fr_array_to_canonical(z[f"lagrange_{tag}"])
        assert d.lagrange_to_coeff(lag) == coeff
        assert bn.fr_array_from_canonical(d.coeff_to_extended(coeff)).tolist() == z[f"extended_{tag}"].tolist()
        bigext = bn.fr_array_to_canonical(z[f"bigext_{tag}"])
        assert bn.fr_array_from_canonical(d.extended_to_coeff(bigext)).tolist() == z[f"bigcoeff_{tag}"].tolist()
        hv = bn.fr_array_to_canonical(z[f"h_{tag}"])
        assert bn.fr_array_from_canonical(d.divide_by_vanishing_poly(hv)).tolist() == z[f"hdiv_{tag}"].tolist()
        # C restatement
        L = fr1
        assert np.array_equal(co.ifft(z[f"lagrange_{tag}"], L(d.omega_inv), k, L(d.ifft_divisor)), z[f"coeff_{tag}"])
        assert np.array_equal(co.coeff_to_extended(z[f"coeff_{tag}"], k, d.extended_k, L(d.extended_omega),
                                                   L(bn.FR_ZETA)), z[f"extended_{tag}"])
        assert np.array_equal(co.extended_to_coeff(z[f"bigext_{tag}"], d.extended_k, L(d.extended_omega_inv),
                                                   L(d.extended_ifft_divisor), L(bn.FR_ZETA),
                                                   d.n * d.quotient_poly_degree), z[f"bigcoeff_{tag}"])
        tev = bn.fr_array_from_canonical(d.t_evaluations)
        assert np.array_equal(co.divide_by_vanishing(z[f"h_{tag}"], d.extended_k, tev), z[f"hdiv_{tag}"])


def test_best_multiexp_matches_definition_golden():
    z = np.load(GOLDEN / "msm_kat.npz")
    for name in z["names"].tolist():
        s, p, r = z[f"s_{name}"], z[f"p_{name}"], z[f"r_{name}"]
        exp = bn.g1_affine_array_to_points(r[None, :])[0]
        # Python restatement of upstream's windowed algorithm
        got_py = h.best_multiexp(bn.fr_array_to_canonical(s), bn.g1_affine_array_to_points(p), num_threads=3)
        assert got_py == exp, name
        for threads in (1, 2, 8):
            assert jac_affine(co.best_multiexp(s, p, threads)) == exp, (name, threads)


def test_c_oracle_msm_medium_vs_python_definition():
    n = 300
    s = co.gen_scalars(1, n)
    p = co.gen_points(2, n)
    exp = bn.g1_msm_naive(bn.fr_array_to_canonical(s), bn.g1_affine_array_to_points(p))
    assert jac_affine(co.best_multiexp(s, p, 8)) == exp


def test_seeded_generators_agree():
    assert np.array_equal(co.gen_scalars(0xA11CE010, 300, 5), bn.seeded_fr_mont_limbs(0xA11CE010, 300, 5))
    assert bn.g1_affine_array_to_points(co.gen_points(0xBA5E0010, 20, 3)) == bn.seeded_g1_points(0xBA5E0010, 20, 3)
    assert all(bn.g1_is_on_curve(p) for p in bn.seeded_g1_points(7, 10))


def test_fft_round_trip_and_linearity_c_large():
    k = 16
    w, wi = omega_for(k), pow(omega_for(k), -1, bn.R)
    a = co.gen_scalars(11, 1 << k)
    f = co.best_fft(a, fr1(w), k)
    back = co.ifft(f, fr1(wi), k, fr1(pow(1 << k, -1, bn.R)))
    assert np.array_equal(back, a)
