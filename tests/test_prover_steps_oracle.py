"""The oracle's prover-step restatements (oracle/prover_steps_cpu.py) against identities that any
correct implementation satisfies — these pin the oracle before the GPU path is compared to it."""
import random

from oracle import bn254 as bn
from oracle import prover_steps_cpu as ps

R = bn.R


def test_batch_invert_identity_and_zeros():
    rnd = random.Random(1)
    a = [rnd.randrange(R) for _ in range(40)]
    a[3] = a[17] = 0
    inv = ps.batch_invert(a)
    for x, y in zip(a, inv):
        assert (x == 0 and y == 0) or x * y % R == 1


def test_eval_polynomial_matches_definition():
    rnd = random.Random(2)
    poly = [rnd.randrange(R) for _ in range(33)]
    x = rnd.randrange(R)
    assert ps.eval_polynomial(poly, x) == sum(c * pow(x, i, R) for i, c in enumerate(poly)) % R
    assert ps.eval_polynomial([], x) == 0


def test_kate_division_reconstructs_dividend():
    rnd = random.Random(3)
    a = [rnd.randrange(R) for _ in range(50)]
    b = rnd.randrange(R)
    q = ps.kate_division(a, b)
    rem = ps.eval_polynomial(a, b)
    # a(X) = q(X) (X - b) + a(b)
    back = [0] * len(a)
    for i, c in enumerate(q):
        back[i + 1] = (back[i + 1] + c) % R
        back[i] = (back[i] - c * b) % R
    back[0] = (back[0] + rem) % R
    assert back == a


def test_permutation_product_is_one_for_the_identity_permutation():
    """With sigma = the identity permutation (sigma_j(omega^i) = delta^j omega^i) every fraction
    is 1, so every z is the all-ones column; with a genuine permutation that only swaps equal
    values, z returns to 1 at the last usable row (the `l_last (z^2 - z)` constraint)."""
    rnd = random.Random(4)
    k, n_cols, chunk = 4, 5, 2
    n = 1 << k
    omega = pow(bn.FR_ROOT_OF_UNITY, 1 << (28 - k), R)
    values = [[rnd.randrange(R) for _ in range(n)] for _ in range(n_cols)]
    ident = [[pow(bn.FR_DELTA, j, R) * pow(omega, i, R) % R for i in range(n)] for j in range(n_cols)]
    beta, gamma = rnd.randrange(R), rnd.randrange(R)
    zs = ps.permutation_products(values, ident, chunk, omega, beta, gamma, 3)
    assert all(z == [1] * n for z in zs) and len(zs) == 3
    # copy constraint between (col 0, row 1) and (col 3, row 2): equal values, swapped labels.
    # The usable rows are 0 .. n - blinding - 1; with blinding 0 the product closes over all rows.
    values[3][2] = values[0][1]
    sig = [row[:] for row in ident]
    sig[0][1], sig[3][2] = ident[3][2], ident[0][1]
    zs = ps.permutation_products(values, sig, chunk, omega, beta, gamma, 0)
    # z_last[n-1] * modified_last[n-1] == 1: recompute the closing factor from the definition
    total = 1
    for j in range(n_cols):
        for i in range(n):
            num = (values[j][i] + beta * ident[j][i] + gamma) % R
            den = (values[j][i] + beta * sig[j][i] + gamma) % R
            total = total * num % R * pow(den, -1, R) % R
    assert total == 1
    assert zs[0][0] == 1 and zs[1][0] == zs[0][n - 1] and zs[2][0] == zs[1][n - 1]


def test_lookup_product_closes_for_a_permuted_pair():
    """If (permuted_input, permuted_table) is a row permutation of (input, table) pairs the
    product over all usable rows telescopes to 1."""
    rnd = random.Random(5)
    n, bf = 16, 3
    a = [rnd.randrange(R) for _ in range(n)]
    s = [rnd.randrange(R) for _ in range(n)]
    usable = n - bf - 1
    perm = list(range(usable))
    rnd.shuffle(perm)
    a2 = [a[p] for p in perm] + a[usable:]
    s2 = [s[p] for p in perm] + s[usable:]
    beta, gamma = rnd.randrange(R), rnd.randrange(R)
    z = ps.lookup_product(a, s, a2, s2, beta, gamma, bf, blinds=[7, 8, 9])
    assert z[0] == 1 and z[usable] == 1 and z[-3:] == [7, 8, 9] and len(z) == n


def test_permute_expression_pair_satisfies_the_lookup_constraints():
    """The defining properties of the permuted pair (what the lookup argument's constraints
    check): same multisets as the originals, and every row has a' == s' or a' == a'[row - 1],
    with a'[0] == s'[0]."""
    import pytest
    rnd = random.Random(8)
    n, bf = 64, 5
    usable = n - bf - 1
    table = [rnd.randrange(R) for _ in range(12)] + [3, 3, 7]
    tab = [rnd.choice(table) for _ in range(usable)]
    tab[:len(table)] = table                                   # every table value present at least once
    inp = [rnd.choice(table) for _ in range(usable)]
    pad = [0] * (bf + 1)
    a, s = ps.permute_expression_pair(inp + pad, tab + pad, bf, blinds=([1] * (bf + 1), [2] * (bf + 1)))
    assert sorted(a[:usable]) == sorted(inp) and sorted(s[:usable]) == sorted(tab)
    assert a[usable:] == [1] * (bf + 1) and s[usable:] == [2] * (bf + 1)
    assert a[0] == s[0]
    assert all(a[i] == s[i] or a[i] == a[i - 1] for i in range(1, usable))
    with pytest.raises(ValueError, match="ConstraintSystemFailure"):
        ps.permute_expression_pair([R - 5] + inp[1:] + pad, tab + pad, bf)


def test_restatements_match_the_definition_golden():
    """tests/golden/prover_steps_kat.npz is produced from the definitions (oracle/make_golden.py
    prover_steps); the statement-by-statement restatements must reproduce it."""
    from pathlib import Path

    import numpy as np
    g = np.load(Path(__file__).parent / "golden" / "prover_steps_kat.npz")
    I, one = bn.fr_array_to_canonical, lambda a: bn.fr_array_to_canonical(a[None, :])[0]
    assert ps.batch_invert(I(g["inv_in"])) == I(g["inv_out"])
    assert ps.eval_polynomial(I(g["eval_poly"]), one(g["eval_point"])) == one(g["eval_out"])
    assert ps.kate_division(I(g["eval_poly"]), one(g["kate_b"])) == I(g["kate_out"])
    k, n_cols, chunk, bf = (int(x) for x in g["perm_shape"])
    beta, gamma = I(g["perm_beta_gamma"])
    omega = pow(bn.FR_ROOT_OF_UNITY, 1 << (28 - k), R)
    zs = ps.permutation_products([I(v) for v in g["perm_values"]], [I(v) for v in g["perm_sigma"]], chunk, omega, beta,
                                 gamma, bf)
    assert zs == [I(z) for z in g["perm_z"]]
    ci, ct, pi, pt = (I(c) for c in g["lookup_cols"])
    assert ps.lookup_product(ci, ct, pi, pt, beta, gamma, bf) == I(g["lookup_z"])
    pts = bn.g1_affine_array_to_points(g["enc_points"])
    assert [bn.g1_to_bytes(p) for p in pts] == [bytes(r) for r in g["enc_bytes"]]
    assert [bn.g1_to_evm_bytes(p) for p in pts] == [bytes(r) for r in g["enc_evm"]]
    assert [bn.g1_from_bytes(bytes(r)) for r in g["enc_bytes"]] == pts
