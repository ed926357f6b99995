"""GPU parity of the prover steps either side of the hot path (csrc/poly.cu) through the C ABI
against oracle/prover_steps_cpu.py on the same seeded inputs, bit-exact; sizes cross every
tile boundary of the kernels (2048-element tiles, 8 elements per thread)."""
import random

import numpy as np
import pytest

from oracle import bn254 as bn
from oracle import c_oracle as co
from oracle import prover_steps_cpu as ps

pytestmark = pytest.mark.gpu

R = bn.R
F = bn.fr_array_from_canonical
I = bn.fr_array_to_canonical


@pytest.mark.parametrize("n", [1, 7, 2048, 2049, 5000, 20000, (1 << 17) + 3])
def test_batch_invert(zk, n):
    a = co.gen_scalars(0xB1 + n, n)
    a[n // 2] = 0
    if n > 10:
        a[0] = 0
        a[n - 1] = 0
    got = a.copy()
    zk.batch_invert(got)
    ai, gi = I(a), I(got)
    # a * a^-1 == 1 (the inverse is unique) and zeros stay zero; spot-check against pow()
    for x, y in list(zip(ai, gi))[:: max(1, n // 500)]:
        assert (x == 0 and y == 0) or x * y % R == 1
    assert gi[n // 2] == 0
    if n <= 5000:
        assert gi == ps.batch_invert(ai)


@pytest.mark.parametrize("n", [0, 1, 5, 2048, 2049, 4097, 1 << 15, (1 << 15) + 77])
def test_eval_polynomial(zk, n):
    rnd = random.Random(n)
    poly = co.gen_scalars(0xE0 + n, max(n, 1))[:n]
    x = rnd.randrange(R)
    got = zk.eval_polynomial(poly, x)
    assert I(got[None, :])[0] == ps.eval_polynomial(I(poly), x)


def test_eval_polynomial_batch_of_points(zk):
    rnd = random.Random(9)
    count, n = 37, 3000
    polys = co.gen_scalars(0xE9, count * n).reshape(count, n, 4)
    pts = [rnd.randrange(R) for _ in range(count)]
    pts[0], pts[1] = 0, 1
    got = I(zk.eval_polynomial_many(polys, pts))
    assert got == [ps.eval_polynomial(I(polys[c]), pts[c]) for c in range(count)]


@pytest.mark.parametrize("n", [1, 2, 9, 2048, 2049, 6145, (1 << 15), (1 << 15) + 5])
def test_kate_division(zk, n):
    rnd = random.Random(n + 1)
    a = co.gen_scalars(0xCA + n, n)
    for b in (rnd.randrange(R), 0, 1):
        got = zk.kate_division(a, b)
        assert got.shape == (n - 1, 4)
        assert I(got) == ps.kate_division(I(a), b)


@pytest.mark.parametrize("k,n_cols,chunk,bf", [(4, 5, 2, 3), (6, 7, 3, 5), (12, 9, 2, 5), (13, 4, 1, 5)])
def test_permutation_products(zk, k, n_cols, chunk, bf):
    n = 1 << k
    rnd = random.Random(k)
    values = [co.gen_scalars(0x70 + j, n) for j in range(n_cols)]
    sigma = [co.gen_scalars(0x90 + j, n) for j in range(n_cols)]
    beta, gamma = rnd.randrange(R), rnd.randrange(R)
    n_sets = -(-n_cols // chunk)
    blinds = [[rnd.randrange(R) for _ in range(bf)] for _ in range(n_sets)]
    omega = pow(bn.FR_ROOT_OF_UNITY, 1 << (28 - k), R)
    want = ps.permutation_products([I(v) for v in values], [I(s) for s in sigma], chunk, omega, beta, gamma, bf, blinds)
    got = zk.permutation_products(values, sigma, chunk, k, beta, gamma, bf, np.stack([F(b) for b in blinds]))
    assert got.shape == (n_sets, n, 4)
    for s in range(n_sets):
        assert I(got[s]) == want[s], f"set {s}"
    # without blinds the computed rows are kept
    want = ps.permutation_products([I(v) for v in values], [I(s) for s in sigma], chunk, omega, beta, gamma, bf)
    got = zk.permutation_products(values, sigma, chunk, k, beta, gamma, bf)
    assert [I(g) for g in got] == want


def test_permutation_product_identity_permutation_is_all_ones(zk):
    k, n_cols = 10, 5
    n = 1 << k
    omega = pow(bn.FR_ROOT_OF_UNITY, 1 << (28 - k), R)
    values = [co.gen_scalars(0x50 + j, n) for j in range(n_cols)]
    ident = [F([pow(bn.FR_DELTA, j, R) * pow(omega, i, R) % R for i in range(n)]) for j in range(n_cols)]
    got = zk.permutation_products(values, ident, 2, k, 12345, 67890, 5)
    assert all(I(g) == [1] * n for g in got)


@pytest.mark.parametrize("k,count,bf", [(4, 1, 3), (11, 3, 5), (12, 2, 5)])
def test_lookup_products(zk, k, count, bf):
    n = 1 << k
    rnd = random.Random(k * 7)
    cols = [[co.gen_scalars(0x30 + 16 * g + j, n) for j in range(count)] for g in range(4)]
    beta, gamma = rnd.randrange(R), rnd.randrange(R)
    blinds = [[rnd.randrange(R) for _ in range(bf)] for _ in range(count)]
    got = zk.lookup_products(cols[0], cols[1], cols[2], cols[3], k, beta, gamma, bf, np.stack([F(b) for b in blinds]))
    for j in range(count):
        want = ps.lookup_product(I(cols[0][j]), I(cols[1][j]), I(cols[2][j]), I(cols[3][j]), beta, gamma, bf, blinds[j])
        assert I(got[j]) == want, f"lookup {j}"
    got = zk.lookup_products(cols[0], cols[1], cols[2], cols[3], k, beta, gamma, bf)
    want = ps.lookup_product(I(cols[0][0]), I(cols[1][0]), I(cols[2][0]), I(cols[3][0]), beta, gamma, bf)
    assert I(got[0]) == want


@pytest.mark.parametrize("k,count,bf,distinct", [(4, 1, 3, 5), (8, 3, 5, 40), (11, 2, 5, 4096), (12, 2, 5, 300), (13, 1, 5, 1 << 12)])
def test_permute_expression_pairs(zk, k, count, bf, distinct):
    """Random lookups whose table covers every input value, small and large value sets (runs of
    repeated inputs, tables with duplicates), sizes on both sides of the 1024-key sort tile."""
    n = 1 << k
    usable = n - bf - 1
    rnd = random.Random(k * 31 + count)
    ins, tbs, blinds = [], [], []
    for c in range(count):
        vals = [rnd.randrange(R) for _ in range(min(distinct, usable))]
        if c % 2:
            vals = [v % 4096 for v in vals]                     # small witnesses: keys differ only in the low limb
        tab = vals + [rnd.choice(vals) for _ in range(usable - len(vals))]
        rnd.shuffle(tab)
        inp = [rnd.choice(vals) for _ in range(usable)]
        pad = [rnd.randrange(R) for _ in range(bf + 1)]         # ignored rows
        ins.append(inp + pad); tbs.append(tab + pad)
        blinds.append(([rnd.randrange(R) for _ in range(bf + 1)], [rnd.randrange(R) for _ in range(bf + 1)]))
    got_in, got_tb = zk.permute_expression_pairs(np.stack([F(x) for x in ins]), np.stack([F(x) for x in tbs]), k, bf,
                                                 np.stack([np.stack([F(b[0]), F(b[1])]) for b in blinds]))
    for c in range(count):
        want_in, want_tb = ps.permute_expression_pair(ins[c], tbs[c], bf, blinds[c])
        assert I(got_in[c]) == want_in, f"permuted input {c}"
        assert I(got_tb[c]) == want_tb, f"permuted table {c}"


def test_permute_expression_pair_missing_value_is_an_error(zk):
    k, bf = 6, 5
    n = 1 << k
    tab = list(range(1, n + 1))
    inp = [1] * n
    inp[3] = 10 ** 30                                           # not in the table
    with pytest.raises(zk.B200zkError, match="ConstraintSystemFailure"):
        zk.permute_expression_pairs(F(inp)[None], F(tab)[None], k, bf)


def test_linear_combination(zk):
    rnd = random.Random(12)
    n, count = 5000, 7
    polys = [co.gen_scalars(0x11C + j, n) for j in range(count)]
    coeffs = [rnd.randrange(R) for _ in range(count)]
    coeffs[2], coeffs[3] = 0, 1
    got = I(zk.linear_combination(polys, coeffs))
    cols = [I(p) for p in polys]
    assert got == [sum(c * col[i] for c, col in zip(coeffs, cols)) % R for i in range(n)]


def test_device_matches_the_definition_golden(zk):
    from pathlib import Path
    g = np.load(Path(__file__).parent / "golden" / "prover_steps_kat.npz")
    one = lambda a: I(a[None, :])[0]
    inv = g["inv_in"].copy()
    zk.batch_invert(inv)
    assert np.array_equal(inv, g["inv_out"])
    assert np.array_equal(zk.eval_polynomial(g["eval_poly"], g["eval_point"]), g["eval_out"])
    assert np.array_equal(zk.kate_division(g["eval_poly"], g["kate_b"]), g["kate_out"])
    k, n_cols, chunk, bf = (int(x) for x in g["perm_shape"])
    beta, gamma = g["perm_beta_gamma"]
    got = zk.permutation_products(list(g["perm_values"]), list(g["perm_sigma"]), chunk, k, beta, gamma, bf)
    assert np.array_equal(got, g["perm_z"])
    ci, ct, pi, pt = g["lookup_cols"]
    got = zk.lookup_products([ci], [ct], [pi], [pt], k, beta, gamma, bf)
    assert np.array_equal(got[0], g["lookup_z"])
    assert np.array_equal(zk.g1_affine_to_bytes(g["enc_points"]), g["enc_bytes"])
    assert np.array_equal(zk.g1_affine_from_bytes(g["enc_bytes"]), g["enc_points"])
    jac = np.zeros((g["enc_points"].shape[0], 12), dtype=np.uint64)
    jac[:, :8] = g["enc_points"]
    jac[1:, 8:] = bn.fr_array_from_canonical([0])[0]          # placeholder, overwritten below
    one_q = bn.ints_to_array([bn.to_mont(1, bn.Q)])[0]
    jac[1:, 8:] = one_q                                        # z = 1 for finite points
    jac[0, 4:8] = one_q                                        # identity = (0, 1, 0)
    assert np.array_equal(zk.g1_to_bytes(jac), g["enc_bytes"])
    assert np.array_equal(zk.g1_to_evm_bytes(jac), g["enc_evm"])
