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


@pytest.mark.gpu
def test_page_locked_host_buffers(zk):
    """b200zk_host_register / b200zk_host_alloc: same results from pageable, registered and
    library-allocated host memory; double registration and foreign pointers are errors."""
    import ctypes as C

    from oracle import c_oracle as co

    k = 14
    n = 1 << k
    s = co.gen_scalars(1, n)
    g = co.gen_points(2, n)
    params = zk.ParamsKZG(g, g)
    aff = bn.g1_jacobian_limbs_to_affine          # the Jacobian representative depends on the addition order
    want = aff(params.commit(s))
    d = zk.EvaluationDomain(3, k)
    want_ntt = s.copy()
    zk.best_fft(want_ntt, d.omega, k)
    reg = s.copy()
    with zk.pinned(reg) as buf:
        assert aff(params.commit(buf)) == want
        zk.best_fft(buf, d.omega, k)
        assert np.array_equal(buf, want_ntt)
        with pytest.raises(zk.B200zkError, match="already page-locked"):
            zk.check(zk.load().b200zk_host_register(C.c_void_p(buf.ctypes.data), buf.nbytes))
    with pytest.raises(zk.B200zkError, match="was not registered"):
        zk.check(zk.load().b200zk_host_unregister(C.c_void_p(reg.ctypes.data)))
    own = zk.host_alloc_fr(n)
    own[:] = s
    assert aff(params.commit(own)) == want
    zk.best_fft(own, d.omega, k)
    assert np.array_equal(own, want_ntt)
    with pytest.raises(zk.B200zkError, match="not a b200zk_host_alloc"):
        zk.check(zk.load().b200zk_host_free(C.c_void_p(reg.ctypes.data)))
    zk.host_free(own)
    params.close()


@pytest.mark.gpu
def test_shutdown_and_reinit_with_pipelines(zk):
    """b200zk_shutdown releases the pipeline streams / events together with everything else; the
    library comes back with a fresh state and the same results."""
    from oracle import c_oracle as co

    lib = zk.load()
    k = 12
    n = 1 << k
    s = co.gen_scalars(5, n)
    g = co.gen_points(6, n)
    w = zk.EvaluationDomain(3, k).omega
    exp_fft = co.best_fft(s, w, k)
    exp_pt = bn.g1_jacobian_limbs_to_affine(co.best_multiexp(s, g))

    def run():
        zk.check(lib.b200zk_msm_upload_pipeline(4, 1))
        zk.check(lib.b200zk_ntt_transfer_pipeline(4, 1))
        try:
            params = zk.ParamsKZG(g, g)
            got_pt = bn.g1_jacobian_limbs_to_affine(params.commit(s))
            params.close()
            a = s.copy()
            zk.best_fft(a, w, k)
            return got_pt, a
        finally:
            zk.check(lib.b200zk_msm_upload_pipeline(0, 0))
            zk.check(lib.b200zk_ntt_transfer_pipeline(4, 22))

    for _ in range(2):
        got_pt, a = run()
        assert got_pt == exp_pt and np.array_equal(a, exp_fft)
        zk.shutdown()
        zk.init(0)
