"""Host-side mirror (b200zk.api): domain constants and argument checking, on the CPU."""
import numpy as np
import pytest

from oracle import bn254 as bn
from oracle import halo2_cpu as h


def test_evaluation_domain_constants_match_oracle():
    from b200zk.api import EvaluationDomain, FR_ROOT_OF_UNITY, FR_ZETA

    assert FR_ROOT_OF_UNITY == bn.FR_ROOT_OF_UNITY and FR_ZETA == bn.FR_ZETA
    for (j, k) in [(3, 4), (4, 5), (5, 15), (4, 15), (9, 10)]:
        d, o = EvaluationDomain(j, k), h.EvaluationDomain(j, k)
        assert d.extended_k == o.extended_k and d.quotient_poly_degree == o.quotient_poly_degree
        f = lambda limbs: bn.fr_array_to_canonical(np.asarray(limbs)[None, :])[0]
        assert f(d.omega) == o.omega and f(d.omega_inv) == o.omega_inv
        assert f(d.extended_omega) == o.extended_omega and f(d.extended_omega_inv) == o.extended_omega_inv
        assert f(d.ifft_divisor) == o.ifft_divisor and f(d.extended_ifft_divisor) == o.extended_ifft_divisor
        assert f(d.g_coset) == o.g_coset and f(d.g_coset_inv) == o.g_coset_inv
        assert bn.fr_array_to_canonical(d.t_evaluations) == o.t_evaluations
    # RSA-SHA256 shape of the reference (src/lib.rs:444: k = 15; degree 4 -> ext_k = 17)
    assert EvaluationDomain(4, 15).extended_k == 17


def test_length_assertions_mirror_upstream():
    import b200zk

    a = np.zeros((7, 4), dtype=np.uint64)
    with pytest.raises(AssertionError):
        b200zk.best_fft(a, 1, 3)              # assert_eq!(n, 1 << log_n)
    with pytest.raises(AssertionError):
        b200zk.best_multiexp(np.zeros((4, 4), dtype=np.uint64), np.zeros((5, 8), dtype=np.uint64))
    d = b200zk.EvaluationDomain(3, 3)
    with pytest.raises(AssertionError):
        d.lagrange_to_coeff(np.zeros((4, 4), dtype=np.uint64))
    with pytest.raises(AssertionError):
        d.extended_to_coeff(np.zeros((8, 4), dtype=np.uint64))
