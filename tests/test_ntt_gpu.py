"""GPU parity of the NTT path through the C ABI against the oracle (bit-exact)."""
import ctypes as C

import numpy as np
import pytest

from oracle import bn254 as bn
from oracle import c_oracle as co
from oracle import halo2_cpu as h
from util import GOLDEN, fr1, omega_for

pytestmark = pytest.mark.gpu


def test_best_fft_golden(zk):
    z = np.load(GOLDEN / "ntt_kat.npz")
    for k in range(1, 9):
        a = z[f"in_{k}"].copy()
        zk.best_fft(a, z[f"omega_{k}"], k)
        assert np.array_equal(a, z[f"out_{k}"]), k


@pytest.mark.parametrize("k", list(range(1, 21)))
def test_best_fft_vs_oracle(zk, k):
    a = co.gen_scalars(0xA11CE000 + k, 1 << k)
    exp = co.best_fft(a, fr1(omega_for(k)), k)
    got = a.copy()
    zk.best_fft(got, omega_for(k), k)
    assert np.array_equal(got, exp)


@pytest.mark.parametrize("k", [21, 22, 23, 24, 25])
def test_best_fft_vs_oracle_at_benchmark_sizes(zk, k):
    """The headline sizes, bit for bit against the C restatement: three passes (7+7+7 ... 9+8+8), the first
    boundary on the lo / hi two-multiplication twiddle path that only exists above 2^20, and from 2^22 the
    host-buffer transfer pipeline (first and last pass in column ranges around chunked copies).  Both the
    host-pointer call and the device-resident one."""
    import torch

    lib = zk.load()
    n = 1 << k
    a = co.gen_scalars(0xA11CE000 + k, n)
    w = omega_for(k)
    exp = co.best_fft(a, fr1(w), k)
    dev = torch.from_numpy(a.view(np.int64).reshape(-1)).cuda()
    torch.cuda.synchronize()              # torch's stream and the library's are not ordered
    zk.check(lib.b200zk_ntt_dev(C.c_void_p(dev.data_ptr()), n, 1, k, C.c_void_p(fr1(w).ctypes.data), None, None))
    torch.cuda.synchronize()
    assert np.array_equal(dev.cpu().numpy().view(np.uint64).reshape(n, 4), exp)
    del dev
    got = a.copy()
    zk.best_fft(got, w, k)
    assert np.array_equal(got, exp)


@pytest.mark.parametrize("k", [2, 5, 9, 10, 13, 16, 18, 19, 20])
def test_best_fft_two_table_twiddles_forced(zk, k):
    """b200zk_ntt_tune(0): every pass boundary takes the lo / hi table form (two multiplications) that
    the large transforms use for their first boundary; pinned here by the oracle at the small sizes, for
    the forward transform and the fused-divisor inverse."""
    lib = zk.load()
    a = co.gen_scalars(0xA11CE000 + k, 1 << k)
    w = omega_for(k)
    try:
        zk.check(lib.b200zk_ntt_tune(0))
        got = a.copy()
        zk.best_fft(got, w, k)
        assert np.array_equal(got, co.best_fft(a, fr1(w), k))
        d = zk.EvaluationDomain(3, k)
        assert np.array_equal(d.lagrange_to_coeff(got.copy()), a)
    finally:
        zk.check(lib.b200zk_ntt_tune(20))


def test_best_fft_edge_inputs(zk):
    k = 11
    n = 1 << k
    w = omega_for(k)
    zero = np.zeros((n, 4), dtype=np.uint64)
    zk.best_fft(zero, w, k)
    assert not zero.any()
    imp = bn.fr_array_from_canonical([1] + [0] * (n - 1))
    zk.best_fft(imp, w, k)
    assert np.array_equal(imp, bn.fr_array_from_canonical([1] * n))
    top = bn.fr_array_from_canonical([bn.R - 1] * n)          # all r-1 -> (r-1)*n at index 0
    zk.best_fft(top, w, k)
    assert bn.fr_array_to_canonical(top) == [(bn.R - 1) * n % bn.R] + [0] * (n - 1)
    # an arbitrary omega of the right order (upstream accepts any omega)
    w3 = pow(w, 3, bn.R)
    a = co.gen_scalars(5, n)
    got = a.copy()
    zk.best_fft(got, w3, k)
    assert np.array_equal(got, co.best_fft(a, fr1(w3), k))


def test_domain_transforms_golden(zk):
    z = np.load(GOLDEN / "domain_kat.npz")
    for (j, k) in z["cases"].tolist():
        tag = f"{j}_{k}"
        d = zk.EvaluationDomain(j, k)
        assert np.array_equal(d.lagrange_to_coeff(z[f"lagrange_{tag}"]), z[f"coeff_{tag}"])
        assert np.array_equal(d.coeff_to_extended(z[f"coeff_{tag}"]), z[f"extended_{tag}"])
        assert np.array_equal(d.extended_to_coeff(z[f"bigext_{tag}"]), z[f"bigcoeff_{tag}"])
        assert np.array_equal(d.divide_by_vanishing_poly(z[f"h_{tag}"]), z[f"hdiv_{tag}"])


@pytest.mark.parametrize("j,k", [(3, 4), (4, 9), (5, 10), (3, 12), (5, 15), (4, 16), (9, 12)])
def test_domain_transforms_vs_oracle(zk, j, k):
    d = zk.EvaluationDomain(j, k)
    o = h.EvaluationDomain(j, k)
    a = co.gen_scalars(77 + k, 1 << k)
    assert np.array_equal(d.lagrange_to_coeff(a), co.ifft(a, fr1(o.omega_inv), k, fr1(o.ifft_divisor)))
    exp_ext = co.coeff_to_extended(a, k, o.extended_k, fr1(o.extended_omega), fr1(bn.FR_ZETA))
    got_ext = d.coeff_to_extended(a)
    assert np.array_equal(got_ext, exp_ext)
    e = co.gen_scalars(99 + k, 1 << o.extended_k)
    keep = o.n * o.quotient_poly_degree
    assert np.array_equal(d.extended_to_coeff(e), co.extended_to_coeff(e, o.extended_k, fr1(o.extended_omega_inv),
                                                                         fr1(o.extended_ifft_divisor),
                                                                         fr1(bn.FR_ZETA), keep))
    tev = bn.fr_array_from_canonical(o.t_evaluations)
    assert np.array_equal(d.divide_by_vanishing_poly(e), co.divide_by_vanishing(e, o.extended_k, tev))
    # round trip through the coset: extended_to_coeff(coeff_to_extended(a)) == a padded
    back = d.extended_to_coeff(got_ext)
    assert np.array_equal(back[: 1 << k], a) and not back[1 << k:].any()


def test_batched_transforms(zk):
    k, cnt = 10, 7
    d = zk.EvaluationDomain(4, k)
    o = h.EvaluationDomain(4, k)
    cols = co.gen_scalars(3, cnt << k).reshape(cnt, 1 << k, 4)
    got = d.lagrange_to_coeff_many(cols)
    ext = d.coeff_to_extended_many(cols)
    for c in range(cnt):
        assert np.array_equal(got[c], co.ifft(cols[c], fr1(o.omega_inv), k, fr1(o.ifft_divisor)))
        assert np.array_equal(ext[c], co.coeff_to_extended(cols[c], k, o.extended_k, fr1(o.extended_omega),
                                                           fr1(bn.FR_ZETA)))


def test_device_resident_fused_vanishing_and_extended_to_coeff(zk):
    import torch

    lib = zk.load()
    j, k = 4, 12
    d, o = zk.EvaluationDomain(j, k), h.EvaluationDomain(j, k)
    N = 1 << o.extended_k
    keep = o.n * o.quotient_poly_degree
    e = co.gen_scalars(123, N)
    tev = bn.fr_array_from_canonical(o.t_evaluations)
    exp = co.extended_to_coeff(co.divide_by_vanishing(e, o.extended_k, tev), o.extended_k, fr1(o.extended_omega_inv),
                               fr1(o.extended_ifft_divisor), fr1(bn.FR_ZETA), keep)
    de = torch.from_numpy(e.view(np.int64)).cuda()
    dt = torch.from_numpy(tev.view(np.int64)).cuda()
    dout = torch.zeros(keep * 4, dtype=torch.int64, device="cuda")
    p = lambda a: C.c_void_p(a.ctypes.data)
    zk.check(lib.b200zk_extended_to_coeff_dev(C.c_void_p(de.data_ptr()), o.extended_k, p(d.extended_omega_inv),
                                              p(d.extended_ifft_divisor), p(d.g_coset), C.c_void_p(dt.data_ptr()),
                                              tev.shape[0], C.c_void_p(dout.data_ptr()), keep, None))
    torch.cuda.synchronize()
    assert np.array_equal(dout.cpu().numpy().view(np.uint64).reshape(keep, 4), exp)


@pytest.mark.parametrize("k", [22, 24, 26])
def test_full_size_properties(zk, k):
    """BASELINE sizes: size-independent properties instead of an oracle run."""
    import torch

    lib = zk.load()
    n = 1 << k
    w, wi = omega_for(k), pow(omega_for(k), -1, bn.R)
    buf = torch.empty(n * 4, dtype=torch.int64, device="cuda")
    zk.check(lib.b200zk_gen_scalars_dev(C.c_void_p(buf.data_ptr()), n, 0xA11CE000 + k, 0))
    orig = buf.clone()
    p = lambda a: C.c_void_p(a.ctypes.data)
    wl, wil, ninv = fr1(w), fr1(wi), fr1(pow(n, -1, bn.R))
    zk.check(lib.b200zk_ntt_dev(C.c_void_p(buf.data_ptr()), n, 1, k, p(wl), None, None))
    torch.cuda.synchronize()
    fwd = buf.clone()
    # (1) out[0] = sum of inputs ; out[n/2] = alternating sum  -- checked on the CPU in O(n)
    a = orig.cpu().numpy().view(np.uint64).reshape(n, 4)
    ints = None
    if k <= 22:
        vals = bn.array_to_ints(a)
        s0 = sum(vals) % bn.R
        s1 = (sum(vals[0::2]) - sum(vals[1::2])) % bn.R
        f = fwd.cpu().numpy().view(np.uint64).reshape(n, 4)
        assert bn.array_to_ints(f[0:1])[0] == s0 and bn.array_to_ints(f[n // 2:n // 2 + 1])[0] == s1
    # (2) inverse round trip restores the input bit for bit
    zk.check(lib.b200zk_ntt_dev(C.c_void_p(buf.data_ptr()), n, 1, k, p(wil), p(ninv), None))
    torch.cuda.synchronize()
    assert torch.equal(buf, orig)


@pytest.mark.parametrize("k,chunks", [(10, 2), (13, 4), (16, 8), (19, 4), (20, 16)])
def test_host_transfer_pipeline_same_transform(zk, k, chunks):
    """b200zk_ntt / b200zk_intt with the first and last pass launched in column ranges around
    chunked 2-D transfers (the large-transform path, forced at small sizes): bit-identical to the
    oracle, on pageable and on page-locked buffers."""
    lib = zk.load()
    n = 1 << k
    a = co.gen_scalars(0xC0DE + k, n)
    w = omega_for(k)
    exp = co.best_fft(a, fr1(w), k)
    d = zk.EvaluationDomain(3, k)
    try:
        zk.check(lib.b200zk_ntt_transfer_pipeline(chunks, 1))
        got = a.copy()
        zk.best_fft(got, w, k)
        assert np.array_equal(got, exp)
        with zk.pinned(a.copy()) as buf:
            zk.best_fft(buf, w, k)
            assert np.array_equal(buf, exp)
        back = d.lagrange_to_coeff(exp.copy())            # intt: the divisor is fused into the chunked last pass
        assert np.array_equal(back, a)
    finally:
        zk.check(lib.b200zk_ntt_transfer_pipeline(4, 22))
