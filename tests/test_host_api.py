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


def test_upload_pipeline_ranges():
    """b200zk_msm_upload_ranges (pure host code, no device): the point ranges a piped commit is fed in cover [0, n)
    exactly once, every inner boundary is a multiple of 256 points, lengths grow by the factor asked for, and small n
    yields fewer ranges instead of empty ones."""
    import ctypes as C

    import b200zk

    lib = b200zk.load()

    def ranges(n, parts, growth):
        begin = (C.c_size_t * (parts + 1))()
        count = C.c_uint32(0)
        b200zk.check(lib.b200zk_msm_upload_ranges(n, parts, growth, begin, C.byref(count)))
        return [int(begin[i]) for i in range(count.value + 1)]

    for n in (0, 1, 255, 256, 257, 300, 4096, 5000, (1 << 14) + 77, 1 << 16, (1 << 22) - 3, 1 << 24, (1 << 26) + 1):
        for parts in (1, 2, 3, 4, 5, 16):
            for growth in (1.0, 1.8, 4.0, 0.5):
                b = ranges(n, parts, growth)
                assert b[0] == 0 and b[-1] == n and len(b) - 1 <= parts
                assert all(x < y for x, y in zip(b, b[1:])) or n == 0
                assert all(x % 256 == 0 for x in b[:-1])
                if n == 0:
                    assert b == [0]
    # the default schedule at the benchmark size: 1/21, 4/21, 16/21 of 2^24 points
    b = ranges(1 << 24, 3, 4.0)
    lens = [y - x for x, y in zip(b, b[1:])]
    assert len(lens) == 3 and abs(lens[0] - (1 << 24) / 21) <= 256 and abs(lens[1] - 4 * (1 << 24) / 21) <= 256
    assert 3.9 < lens[1] / lens[0] < 4.1 and 3.9 < lens[2] / lens[1] < 4.1
    # equal ranges with growth 1
    b = ranges(1 << 22, 4, 1.0)
    lens = [y - x for x, y in zip(b, b[1:])]
    assert len(lens) == 4 and max(lens) - min(lens) <= 1024
    # parts out of range
    begin = (C.c_size_t * 18)()
    count = C.c_uint32(0)
    assert lib.b200zk_msm_upload_ranges(100, 0, 1.0, begin, C.byref(count)) != 0
    assert lib.b200zk_msm_upload_ranges(100, 17, 1.0, begin, C.byref(count)) != 0


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
