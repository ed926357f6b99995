"""CPU restatement (Python big integers) of the prover loops either side of the MSM / NTT hot
path (SURVEY.md section 8 f, ranks 1 and 2).

TEST INFRASTRUCTURE ONLY — see the header of ``oracle/bn254.py``.  PARITY PIN: no upstream output
vectors exist (the reference never runs the prover); permutation_products,
eval_polynomial and kate_division are pinned end to end by the reference's verifier contract
accepting the proofs built from them (tests/test_square_proof_oracle.py), the lookup functions
only by identities.  Each function follows the published algorithm of the pinned
dependency ([DEP] halo2_proofs 0.2.0 @ v2023_01_20, reference ``Cargo.lock:469-471``; ff 0.12,
``Cargo.lock:339-341``) and is checked against identities that hold for any correct
implementation (tests/test_prover_steps_oracle.py).

  * ``ff::BatchInvert::batch_invert``                                  batch_invert
  * ``halo2_proofs/src/arithmetic.rs``  eval_polynomial, kate_division  eval_polynomial, kate_division
  * ``halo2_proofs/src/plonk/permutation/prover.rs``  Argument::commit  permutation_products
  * ``halo2_proofs/src/plonk/lookup/prover.rs``  Permuted::commit_product  lookup_product
  * ``halo2_proofs/src/plonk/lookup/prover.rs``  permute_expression_pair   permute_expression_pair
"""
from __future__ import annotations

from . import bn254 as bn

R = bn.R


def batch_invert(a):
    """ff BatchInvert: every non-zero element replaced by its inverse, zeros left alone."""
    return [pow(x, -1, R) if x % R else 0 for x in a]


def eval_polynomial(poly, point):
    """arithmetic.rs eval_polynomial: Horner from the top coefficient."""
    acc = 0
    for c in reversed(poly):
        acc = (acc * point + c) % R
    return acc


def kate_division(a, b):
    """arithmetic.rs kate_division: q = a / (X - b) (remainder dropped).  Upstream negates b and
    walks the coefficients from the top: q[i-1] = a[i] + b * q[i], with q[len-1] = 0."""
    q = [0] * (len(a) - 1)
    tmp = 0
    for i in range(len(a) - 1, 0, -1):
        tmp = (a[i] + b * tmp) % R
        q[i - 1] = tmp
    return q


def permutation_products(values, sigma, chunk_len, omega, beta, gamma, blinding_factors, blinds=None):
    """permutation/prover.rs Argument::commit, the z polynomials (Lagrange basis) of every set.
    ``values[j]`` / ``sigma[j]``: column j's values and its permutation polynomial's values
    (pkey.permutations[j]).  ``blinds[s]``: the scalars upstream draws for the last
    blinding_factors rows of set s (left as computed when None)."""
    n = len(values[0])
    deltaomega = 1
    last_z = 1
    out = []
    for s, lo in enumerate(range(0, len(values), chunk_len)):
        cols = range(lo, min(lo + chunk_len, len(values)))
        modified = [1] * n
        for j in cols:
            for i in range(n):
                modified[i] = modified[i] * (beta * sigma[j][i] + gamma + values[j][i]) % R
        modified = batch_invert(modified)
        for j in cols:
            cur = deltaomega
            for i in range(n):
                modified[i] = modified[i] * (cur * beta + gamma + values[j][i]) % R
                cur = cur * omega % R
            deltaomega = deltaomega * bn.FR_DELTA % R
        z = [last_z]
        for row in range(1, n):
            z.append(z[row - 1] * modified[row - 1] % R)
        if blinds is not None:
            for t in range(blinding_factors):
                z[n - blinding_factors + t] = blinds[s][t] % R
        last_z = z[n - (blinding_factors + 1)]
        out.append(z)
    return out


def lookup_product(compressed_input, compressed_table, permuted_input, permuted_table, beta, gamma,
                   blinding_factors, blinds=None):
    """lookup/prover.rs Permuted::commit_product: z[0] = 1 and the running product of
    (a + beta)(s + gamma) / ((a' + beta)(s' + gamma)); the first n - blinding_factors rows are
    kept, the rest are the caller's blinding scalars."""
    n = len(compressed_input)
    prod = [(beta + a) * (gamma + s) % R for a, s in zip(permuted_input, permuted_table)]
    prod = batch_invert(prod)
    for i in range(n):
        prod[i] = prod[i] * (compressed_input[i] + beta) % R * (compressed_table[i] + gamma) % R
    z, state = [], 1
    for cur in [1] + prod:
        state = state * cur % R
        z.append(state)
    z = z[: n - blinding_factors]
    tail = [b % R for b in blinds] if blinds is not None else None
    if tail is None:
        # no RNG here: continue the running product so that the device's "computed" rows compare
        state = z[-1]
        tail = []
        for i in range(n - blinding_factors - 1, n - 1):
            state = state * prod[i] % R
            tail.append(state)
    return z + tail


def permute_expression_pair(input_expression, table_expression, blinding_factors, blinds=None):
    """lookup/prover.rs permute_expression_pair, statement by statement: sort the usable rows of
    the input (Fr's Ord = canonical integer order), count the table's values in an ordered map,
    give every first occurrence of an input value its own value in the table column (taking one
    instance out of the map; a missing value is ConstraintSystemFailure), then pop the repeated
    rows from the END while walking the leftover table values in ascending order."""
    n = len(input_expression)
    usable = n - (blinding_factors + 1)
    permuted_input = sorted(v % R for v in input_expression[:usable])
    leftover = {}
    for v in table_expression[:usable]:
        leftover[v % R] = leftover.get(v % R, 0) + 1
    permuted_table = [0] * usable
    repeated_rows = []
    for row, v in enumerate(permuted_input):
        if row == 0 or v != permuted_input[row - 1]:
            permuted_table[row] = v
            if leftover.get(v, 0) <= 0:
                raise ValueError("ConstraintSystemFailure")
            leftover[v] -= 1
        else:
            repeated_rows.append(row)
    for v in sorted(leftover):
        for _ in range(leftover[v]):
            permuted_table[repeated_rows.pop()] = v
    assert not repeated_rows
    tail_in = [b % R for b in blinds[0]] if blinds is not None else [0] * (blinding_factors + 1)
    tail_tb = [b % R for b in blinds[1]] if blinds is not None else [0] * (blinding_factors + 1)
    return permuted_input + tail_in, permuted_table + tail_tb
