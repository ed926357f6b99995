"""bench.py's own exact check of a commitment against the discrete logs of the synthetic SRS
(`msm_point_verified`): its plain-integer curve arithmetic and dot product, pinned by the oracle."""
import numpy as np

import bench
from oracle import bn254 as bn
from oracle import c_oracle as co


def test_generator_mul_matches_oracle():
    for s in (0, 1, 2, 3, 5, bn.R - 1, bn.R, 0xDEADBEEF << 200):
        want = bn.g1_mul(bn.G1_GEN, s % bn.R)
        got = bench.generator_mul(s)
        assert (got is None and want is None) or tuple(got) == tuple(want)


def test_point_stream_and_dot_product_verify_a_cpu_commitment():
    k, n = 10, 1 << 10
    seed_s, seed_p, start = bench.SEED_S + k, bench.SEED_P + k, 3 * n      # a rank-3 slice
    scalars = co.gen_scalars(seed_s, n, start)
    points = co.gen_points(seed_p, n, start)
    t = bench.point_discrete_logs(seed_p, start, n)
    # the documented stream: P_i = [t_i] G
    for i in (0, 1, n - 1):
        assert tuple(bn.g1_affine_array_to_points(points[i:i + 1])[0]) == tuple(bn.g1_mul(bn.G1_GEN, int(t[i])))
    dot = bench.dot_mod_r(scalars, t)
    assert dot == sum(int.from_bytes(scalars[i].tobytes(), "little") * int(t[i]) for i in range(n)) % bn.R
    point = co.best_multiexp(scalars, points)
    assert bench.commitment_matches_discrete_logs(point, dot)
    assert bench.jacobian_limbs_to_affine(point) == tuple(bn.g1_jacobian_limbs_to_affine(point))
    bad = point.copy()
    bad[0] ^= np.uint64(1)
    assert not bench.commitment_matches_discrete_logs(bad, dot)
