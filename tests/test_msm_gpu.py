"""GPU parity of the MSM path through the C ABI against the oracle (same group element,
compared in affine as SURVEY.md section 8 c prescribes)."""
import ctypes as C

import numpy as np
import pytest

from oracle import bn254 as bn
from oracle import c_oracle as co
from util import GOLDEN, fr1, jac_affine, omega_for

pytestmark = pytest.mark.gpu


def test_msm_golden(zk):
    z = np.load(GOLDEN / "msm_kat.npz")
    for name in z["names"].tolist():
        got = jac_affine(zk.best_multiexp(z[f"s_{name}"], z[f"p_{name}"]))
        exp = bn.g1_affine_array_to_points(z[f"r_{name}"][None, :])[0]
        assert got == exp, name


def test_msm_empty_is_identity(zk):
    out = zk.best_multiexp(np.zeros((0, 4), dtype=np.uint64), np.zeros((0, 8), dtype=np.uint64))
    assert jac_affine(out) is None
    assert bn.array_to_ints(out)[1] == bn.FQ_MONT["R"]        # (0, 1, 0) like G1::identity()


@pytest.mark.parametrize("n", [1, 2, 3, 31, 32, 33, 255, 1000, 4097, 1 << 14, (1 << 16) + 3, 1 << 18])
def test_msm_vs_oracle(zk, n):
    s = co.gen_scalars(0xA11CE000 + n, n)
    p = co.gen_points(0xBA5E0000 + n, n)
    assert jac_affine(zk.best_multiexp(s, p)) == jac_affine(co.best_multiexp(s, p))


def test_msm_adversarial_scalar_sets(zk):
    n = 6000
    p = co.gen_points(9, n)
    sets = {
        "all r-1": bn.fr_array_from_canonical([bn.R - 1] * n),
        "all one": bn.fr_array_from_canonical([1] * n),
        "all zero": bn.fr_array_from_canonical([0] * n),
        "small": bn.fr_array_from_canonical([(i * 7) % 65536 for i in range(n)]),
        "window boundaries": bn.fr_array_from_canonical([(1 << (i % 254)) - (i % 3) for i in range(n)]),
    }
    sparse = co.gen_scalars(5, n)
    sparse[np.arange(n) % 10 != 0] = 0
    sets["90% zero"] = sparse
    for name, s in sets.items():
        assert jac_affine(zk.best_multiexp(s, p)) == jac_affine(co.best_multiexp(s, p)), name


def test_msm_adversarial_points(zk):
    n = 3000
    s = co.gen_scalars(4, n)
    same = np.repeat(co.gen_points(3, 1), n, axis=0)          # one bucket run per window, P + P cases
    assert jac_affine(zk.best_multiexp(s, same)) == jac_affine(co.best_multiexp(s, same))
    p = co.gen_points(3, n)
    p[::3] = 0                                                # identity bases
    assert jac_affine(zk.best_multiexp(s, p)) == jac_affine(co.best_multiexp(s, p))
    pts = bn.g1_affine_array_to_points(p[1:3])
    p[5] = bn.g1_affine_array_from_points([bn.g1_neg(pts[0])])[0]   # P and -P with equal scalars
    s[5] = s[1]
    assert jac_affine(zk.best_multiexp(s, p)) == jac_affine(co.best_multiexp(s, p))


def test_params_kzg_commit(zk):
    n = 1 << 12
    g, gl = co.gen_points(21, n), co.gen_points(22, n)
    params = zk.ParamsKZG(g, gl)
    poly = co.gen_scalars(23, n)
    assert jac_affine(params.commit(poly)) == jac_affine(co.best_multiexp(poly, g))
    assert jac_affine(params.commit_lagrange(poly)) == jac_affine(co.best_multiexp(poly, gl))
    short = poly[:1000]                                       # poly shorter than the SRS
    assert jac_affine(params.commit(short)) == jac_affine(co.best_multiexp(short, g[:1000]))
    with pytest.raises(AssertionError):
        params.commit(co.gen_scalars(1, n + 1))
    params.close()


def test_g1_sum_matches_fold(zk):
    n = 900
    s, p = co.gen_scalars(31, n), co.gen_points(32, n)
    parts = np.stack([zk.best_multiexp(s[i:i + 300], p[i:i + 300]) for i in range(0, n, 300)])
    ident = zk.best_multiexp(s[:0], p[:0])
    parts = np.concatenate([parts, ident[None, :]])
    assert jac_affine(zk.g1_sum(parts)) == jac_affine(co.best_multiexp(s, p))


@pytest.mark.parametrize("k", [20, 22])
def test_msm_full_size_linearity(zk, k):
    """BASELINE sizes: MSM(s, P) with P_i = [t_i]G equals [sum s_i t_i] G (checked with
    Python integers), and MSM(2s, P) == 2 MSM(s, P)."""
    import torch

    lib = zk.load()
    n = 1 << k
    ds = torch.empty(n * 4, dtype=torch.int64, device="cuda")
    db = torch.empty(n * 8, dtype=torch.int64, device="cuda")
    zk.check(lib.b200zk_gen_scalars_dev(C.c_void_p(ds.data_ptr()), n, 0xA11CE000 + k, 0))
    zk.check(lib.b200zk_gen_points_dev(C.c_void_p(db.data_ptr()), n, 0xBA5E0000 + k, 0))
    out = np.zeros(12, dtype=np.uint64)
    zk.check(lib.b200zk_msm_g1_dev(C.c_void_p(ds.data_ptr()), C.c_void_p(db.data_ptr()), n, C.c_void_p(out.ctypes.data), None))
    s = ds.cpu().numpy().view(np.uint64).reshape(n, 4)
    # t_i stream of oracle/bn254.py seeded_point_scalar, vectorised
    with np.errstate(over="ignore"):
        t = bn._splitmix64_np(np.uint64(((0xBA5E0000 + k) << 32) & ((1 << 64) - 1)) + np.arange(n, dtype=np.uint64)) | np.uint64(1)
    rinv = pow(1 << 256, -1, bn.R)
    acc = 0
    svals = bn.array_to_ints(s)
    for si, ti in zip(svals, t.tolist()):
        acc += si * ti
    scalar = acc % bn.R * rinv % bn.R                          # scalars are Montgomery representatives
    assert jac_affine(out) == bn.g1_mul(bn.G1_GEN, scalar)


def _dot_mod_r(s_limbs: np.ndarray, t: np.ndarray) -> int:
    """sum_i s_i * t_i for s (n, 4) u64 limbs and t (n,) u64, exactly, without a Python loop over
    n: 32-bit pieces of s against 16-bit pieces of t, 65536 rows at a time (48-bit products, so a
    block sum stays below 2^64)."""
    n = s_limbs.shape[0]
    s32 = np.ascontiguousarray(s_limbs).view(np.uint32).reshape(n, 8).astype(np.uint64)
    t16 = np.ascontiguousarray(t).view(np.uint16).reshape(n, 4).astype(np.uint64)
    total = 0
    for lo in range(0, n, 1 << 16):
        m = s32[lo:lo + (1 << 16)].T @ t16[lo:lo + (1 << 16)]          # (8, 4) exact in uint64
        for a in range(8):
            for b in range(4):
                total += int(m[a, b]) << (32 * a + 16 * b)
    return total


def test_commit_at_benchmark_size_matches_known_discrete_logs(zk):
    """The exact configuration bench.py times — 2^24 registered points with the window table
    (c = 22, shared bucket set), device-resident scalars — against [sum s_i t_i] G computed with
    exact integer arithmetic; the plain best_multiexp path on the same inputs must agree."""
    import torch

    lib = zk.load()
    k = 24
    n = 1 << k
    ds = torch.empty(n * 4, dtype=torch.int64, device="cuda")
    db = torch.empty(n * 8, dtype=torch.int64, device="cuda")
    zk.check(lib.b200zk_gen_scalars_dev(C.c_void_p(ds.data_ptr()), n, 0xA11CE000 + k, 0))
    zk.check(lib.b200zk_gen_points_dev(C.c_void_p(db.data_ptr()), n, 0xBA5E0000 + k, 0))
    hb = db.cpu().numpy().view(np.uint64).reshape(n, 8)
    h = C.c_uint64(0)
    zk.check(lib.b200zk_bases_register(C.c_void_p(hb.ctypes.data), n, C.byref(h)))
    d_out = torch.zeros(12, dtype=torch.int64, device="cuda")
    zk.check(lib.b200zk_msm_g1_registered_dev(h.value, C.c_void_p(ds.data_ptr()), n, 1, n, C.c_void_p(d_out.data_ptr()), None))
    torch.cuda.synchronize()
    got = d_out.cpu().numpy().view(np.uint64)
    plain = np.zeros(12, dtype=np.uint64)
    zk.check(lib.b200zk_msm_g1_dev(C.c_void_p(ds.data_ptr()), C.c_void_p(db.data_ptr()), n, C.c_void_p(plain.ctypes.data), None))
    s = ds.cpu().numpy().view(np.uint64).reshape(n, 4)
    # the same commit through the host-pointer call (what bench.py's e2e leg times): the upload pipeline's default
    # schedule (three point ranges growing 4x), then the schedule it derives from the first call's copy / commit times
    piped = [np.zeros(12, dtype=np.uint64) for _ in range(3)]
    with zk.pinned(s):
        for out in piped:
            zk.check(lib.b200zk_msm_g1_registered(h.value, C.c_void_p(s.ctypes.data), n, C.c_void_p(out.ctypes.data)))
    zk.check(lib.b200zk_bases_evict(h.value))
    with np.errstate(over="ignore"):
        t = bn._splitmix64_np(np.uint64(((0xBA5E0000 + k) << 32) & ((1 << 64) - 1)) + np.arange(n, dtype=np.uint64)) | np.uint64(1)
    scalar = _dot_mod_r(s, t) % bn.R * pow(1 << 256, -1, bn.R) % bn.R       # scalars are Montgomery representatives
    want = bn.g1_mul(bn.G1_GEN, scalar)
    assert jac_affine(got) == want
    assert jac_affine(plain) == want
    assert [jac_affine(o) for o in piped] == [want] * 3


@pytest.mark.parametrize("precompute", [True, False])
def test_params_kzg_window_table_and_batch(zk, precompute):
    """Registered bases with / without the per-window table, single and batched commits,
    including scalars that exercise every exceptional addition."""
    n = 1 << 11
    g, gl = co.gen_points(51, n), co.gen_points(52, n)
    g[7] = g[6]                                   # repeated base
    g[9] = 0                                      # identity base
    params = zk.ParamsKZG(g, gl, precompute_windows=precompute)
    cols = co.gen_scalars(53, 5 * n).reshape(5, n, 4).copy()
    cols[1] = bn.fr_array_from_canonical([1] * n)                       # all ones
    cols[2] = bn.fr_array_from_canonical([bn.R - 1] * n)                # all r-1
    cols[3][np.arange(n) % 7 != 0] = 0                                  # sparse
    cols[4] = bn.fr_array_from_canonical([(3 * i) % 4096 for i in range(n)])   # small witnesses
    exp = [jac_affine(co.best_multiexp(cols[j], g)) for j in range(5)]
    assert [jac_affine(params.commit(cols[j])) for j in range(5)] == exp
    assert [jac_affine(p) for p in params.commit_many(cols)] == exp
    expl = [jac_affine(co.best_multiexp(cols[j][:1500], gl[:1500])) for j in range(5)]
    assert [jac_affine(p) for p in params.commit_lagrange_many(np.ascontiguousarray(cols[:, :1500]))] == expl
    params.close()


def test_commit_many_prover_shape(zk):
    """The RSA-SHA256 shape of the reference (k = 15, src/lib.rs:444): a batch of advice
    columns committed against g_lagrange in one pipeline."""
    n, cnt = 1 << 15, 6
    gl = co.gen_points(61, n)
    params = zk.ParamsKZG(gl, gl)
    cols = co.gen_scalars(62, cnt * n).reshape(cnt, n, 4)
    got = [jac_affine(p) for p in params.commit_lagrange_many(cols)]
    assert got == [jac_affine(co.best_multiexp(cols[j], gl)) for j in range(cnt)]
    params.close()


@pytest.mark.parametrize("parts,n", [(2, 1 << 12), (3, 5000), (4, (1 << 14) + 77), (16, 300), (5, 1 << 16)])
def test_commit_upload_pipeline_same_point(zk, parts, n):
    """b200zk_msm_g1_registered fed in point ranges that share the bucket set (the upload
    pipeline of large host-side commits): the same group element as the one-shot path and as
    the oracle, with scalars that force later ranges to add into buckets earlier ranges filled
    (few distinct digits), empty ranges of buckets, repeated and identity bases."""
    lib = zk.load()
    g = co.gen_points(71, n)
    g[5] = g[4]
    g[11] = 0
    s = co.gen_scalars(72 + parts, n)
    m = n // 8
    s[n // 2:n // 2 + m] = bn.fr_array_from_canonical([bn.R - 1] * m)       # one bucket per window, every range
    s[m:2 * m] = 0                                                          # dropped digits
    s[n - m:] = bn.fr_array_from_canonical([(5 * i) % 97 for i in range(m)])
    want = jac_affine(co.best_multiexp(s, g))
    params = zk.ParamsKZG(g, g)
    try:
        zk.check(lib.b200zk_msm_upload_pipeline(1, 1 << 22))
        one_shot = jac_affine(params.commit(s))
        zk.check(lib.b200zk_msm_upload_pipeline(parts, 1))
        piped = jac_affine(params.commit(s))
        piped_short = jac_affine(params.commit(s[: n - 3]))          # ragged last range, fewer scalars than bases
    finally:
        zk.check(lib.b200zk_msm_upload_pipeline(0, 0))
        params.close()
    assert one_shot == want
    assert piped == want
    assert piped_short == jac_affine(co.best_multiexp(s[: n - 3], g[: n - 3]))


def test_two_commits_on_two_streams_run_concurrently_and_agree(zk):
    """`*_dev` calls are asynchronous and stream-safe: two different commitments queued back to back on
    two streams (no host synchronisation in between, so their kernels overlap on the device and would
    trample a shared scratch arena) give the points the same calls give one at a time; the same for a
    commit on one stream under a transform on another."""
    import torch

    lib = zk.load()
    k = 16
    n = 1 << k
    pts = co.gen_points(0xBA5E0000 + k, n)
    h = C.c_uint64(0)
    zk.check(lib.b200zk_bases_register(C.c_void_p(pts.ctypes.data), n, C.byref(h)))
    sa = torch.from_numpy(co.gen_scalars(1, n).view(np.int64).reshape(-1)).cuda()
    sb = torch.from_numpy(co.gen_scalars(2, n).view(np.int64).reshape(-1)).cuda()
    db = torch.from_numpy(pts.view(np.int64).reshape(-1)).cuda()
    ref = torch.zeros(4, 12, dtype=torch.int64, device="cuda")
    vp = lambda t: C.c_void_p(t.data_ptr())
    zk.check(lib.b200zk_msm_g1_registered_dev(h.value, vp(sa), n, 1, n, vp(ref[0]), None))
    zk.check(lib.b200zk_msm_g1_registered_dev(h.value, vp(sb), n, 1, n, vp(ref[1]), None))
    zk.check(lib.b200zk_msm_g1_dev_async(vp(sa), vp(db), n, vp(ref[2]), None))
    zk.check(lib.b200zk_msm_g1_dev_async(vp(sb), vp(db), n, vp(ref[3]), None))
    torch.cuda.synchronize()
    w = fr1(omega_for(k))
    ntt_ref = sa.clone()
    torch.cuda.synchronize()
    zk.check(lib.b200zk_ntt_dev(vp(ntt_ref), n, 1, k, C.c_void_p(w.ctypes.data), None, None))
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    p1, p2 = C.c_void_p(s1.cuda_stream), C.c_void_p(s2.cuda_stream)
    canon = lambda t: zk.g1_to_bytes(np.ascontiguousarray(t.cpu().numpy().view(np.uint64).reshape(-1, 12)))
    for rep in range(4):
        got = torch.zeros(4, 12, dtype=torch.int64, device="cuda")
        t1, t2 = sa.clone(), sb.clone()
        torch.cuda.synchronize()
        zk.check(lib.b200zk_msm_g1_registered_dev(h.value, vp(sa), n, 1, n, vp(got[0]), p1))
        zk.check(lib.b200zk_msm_g1_registered_dev(h.value, vp(sb), n, 1, n, vp(got[1]), p2))
        zk.check(lib.b200zk_msm_g1_dev_async(vp(sa), vp(db), n, vp(got[2]), p1))
        zk.check(lib.b200zk_msm_g1_dev_async(vp(sb), vp(db), n, vp(got[3]), p2))
        zk.check(lib.b200zk_ntt_dev(vp(t1), n, 1, k, C.c_void_p(w.ctypes.data), None, p1))
        zk.check(lib.b200zk_msm_g1_registered_dev(h.value, vp(sa), n, 1, n, vp(got[0]), p2))
        torch.cuda.synchronize()
        # Jacobian coordinates depend on the (atomic) order of the pairs inside a bucket: compare the points
        assert np.array_equal(canon(got), canon(ref)), rep
        assert torch.equal(t1, ntt_ref), rep
    exp = jac_affine(co.best_multiexp(co.gen_scalars(1, n), pts))
    assert jac_affine(ref[0].cpu().numpy().view(np.uint64)) == exp and jac_affine(ref[2].cpu().numpy().view(np.uint64)) == exp
    zk.check(lib.b200zk_stream_release(p1))
    zk.check(lib.b200zk_stream_release(p2))
    zk.check(lib.b200zk_bases_evict(h.value))


@pytest.mark.parametrize("kind", ["zeros", "one_nonzero", "bits"])
def test_sparse_scalars_plan_on_the_device(zk, kind):
    """The pair count never visits the host: the accumulation is launched for the worst case and plans
    itself from the count the sort produced.  Scalars with (almost) no non-zero digits are the far end."""
    k = 12
    n = 1 << k
    pts = co.gen_points(0xBA5E0000 + k, n)
    sc = np.zeros((n, 4), dtype=np.uint64)
    if kind == "one_nonzero":
        sc[n // 3] = bn.fr_array_from_canonical([bn.R - 5])[0]
    elif kind == "bits":
        sc[::2] = bn.fr_array_from_canonical([1])[0]
    exp = jac_affine(co.best_multiexp(sc, pts))
    assert jac_affine(zk.best_multiexp(sc, pts)) == exp
    params = zk.ParamsKZG(pts, pts)
    try:
        assert jac_affine(params.commit_lagrange(sc)) == exp
    finally:
        params.close()


@pytest.mark.parametrize("kind", ["uniform", "all_equal", "all_r_minus_1", "zero_one", "sparse"])
def test_partition_sort_sizes_with_skewed_scalars(zk, kind):
    """2^19 points is past the size from which the (bucket, point) pairs are sorted in two levels through shared
    memory (csrc/msm.cu, partition sort).  Scalars that put every point of a window into one bucket, or almost none
    anywhere, are the far ends of its tile-local counters; both the window-table commit and the plain best_multiexp
    are checked against [sum s_i t_i] G with exact integers."""
    import torch

    lib = zk.load()
    k = 19
    n = 1 << k
    db = torch.empty(n * 8, dtype=torch.int64, device="cuda")
    zk.check(lib.b200zk_gen_points_dev(C.c_void_p(db.data_ptr()), n, 0xBA5E0000 + k, 0))
    s = co.gen_scalars(0xA11CE000 + k, n)
    if kind == "all_equal":
        s[:] = s[7]
    elif kind == "all_r_minus_1":
        s[:] = bn.fr_array_from_canonical([bn.R - 1])[0]
    elif kind == "zero_one":
        s[:] = bn.fr_array_from_canonical([1])[0]
        s[::2] = 0
    elif kind == "sparse":
        s[np.arange(n) % 10 != 0] = 0
    ds = torch.from_numpy(s.view(np.int64).reshape(-1)).cuda()
    torch.cuda.synchronize()
    with np.errstate(over="ignore"):
        t = bn._splitmix64_np(np.uint64(((0xBA5E0000 + k) << 32) & ((1 << 64) - 1)) + np.arange(n, dtype=np.uint64)) | np.uint64(1)
    want = bn.g1_mul(bn.G1_GEN, _dot_mod_r(s, t) % bn.R * pow(1 << 256, -1, bn.R) % bn.R)
    plain = np.zeros(12, dtype=np.uint64)
    zk.check(lib.b200zk_msm_g1_dev(C.c_void_p(ds.data_ptr()), C.c_void_p(db.data_ptr()), n, C.c_void_p(plain.ctypes.data), None))
    assert jac_affine(plain) == want
    hb = db.cpu().numpy().view(np.uint64).reshape(n, 8)
    h = C.c_uint64(0)
    zk.check(lib.b200zk_bases_register(C.c_void_p(hb.ctypes.data), n, C.byref(h)))
    try:
        d_out = torch.zeros(12, dtype=torch.int64, device="cuda")
        zk.check(lib.b200zk_msm_g1_registered_dev(h.value, C.c_void_p(ds.data_ptr()), n, 1, n, C.c_void_p(d_out.data_ptr()), None))
        torch.cuda.synchronize()
        assert jac_affine(d_out.cpu().numpy().view(np.uint64)) == want
    finally:
        zk.check(lib.b200zk_bases_evict(h.value))
