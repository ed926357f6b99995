"""A small circuit with two lookup arguments, driven through the prover's lookup chain
(`lookup::prover`: compress -> permute_expression_pair -> commit_product, then
`Evaluator::evaluate_h`), shared by the oracle and GPU tests.

The check is the one the verifier ultimately relies on: for a satisfying witness the y-folded
numerator is divisible by X^n - 1, i.e. after divide_by_vanishing_poly the inverse transform over
the extended domain has no coefficient at or above n * (degree - 1) — and for an inconsistent
witness it has.  It ties permute_expression_pair, the lookup grand product and the lookup terms of
evaluate_h together (none of which the reference's Solidity verifier covers: its circuit has no
lookups)."""
from __future__ import annotations

import random

from oracle import bn254 as bn
from oracle import halo2_cpu as h
from oracle import quotient_cpu as q
from quotient_cases import lagrange_cols

R = bn.R
BF = 5
DEGREE = 4            # lookup::Argument::required_degree = max(4, 2 + input_degree + table_degree)


def build(k=5, seed=3, corrupt=False):
    rnd = random.Random(seed)
    d = h.EvaluationDomain(DEGREE, k)
    n, usable = d.n, d.n - (BF + 1)
    t0 = [(7 * i + 1) % R for i in range(usable)] + [0] * (BF + 1)               # fixed 0: table
    t1 = [v * v % R for v in t0]                                                  # fixed 1: table of squares
    sel = [1 if i < usable and i % 3 != 2 else 0 for i in range(n)]               # fixed 2: gate selector
    a0 = [rnd.choice(t0[:usable]) for _ in range(usable)] + [rnd.randrange(R) for _ in range(BF + 1)]
    a1 = [v * v % R for v in a0[:usable]] + [rnd.randrange(R) for _ in range(BF + 1)]
    fixed, advice = [t0, t1, sel], [a0, a1]
    gates = [q.Fixed(2) * (q.Advice(1) - q.Advice(0) * q.Advice(0))]
    lookups = [([q.Advice(0)], [q.Fixed(0)]),
               ([q.Advice(0), q.Advice(1)], [q.Fixed(0), q.Fixed(1)])]
    ch = {name: rnd.randrange(R) for name in ("theta", "beta", "gamma", "y")}
    blinds = [{"pair": ([rnd.randrange(R) for _ in range(BF + 1)], [rnd.randrange(R) for _ in range(BF + 1)]),
               "z": [rnd.randrange(R) for _ in range(BF)]} for _ in lookups]

    def compress(exprs):
        out = [0] * n
        for e in exprs:
            col = {"advice": advice, "fixed": fixed}[e.op][e.args[0]]
            out = [(o * ch["theta"] + v) % R for o, v in zip(out, col)]
        return out

    compressed = [(compress(i), compress(t)) for i, t in lookups]
    if corrupt:      # the witness changes after the permuted columns were derived from it
        advice[0][1] = next(v for v in t0[:usable] if v != advice[0][1])
        advice[1][1] = advice[0][1] ** 2 % R
    l0, l_last, l_blind, l_active = lagrange_cols(d, BF)
    return {"domain": d, "k": k, "fixed": fixed, "advice": advice, "gates": gates, "lookups": lookups,
            "compressed": compressed, "blinds": blinds, "l0": l0, "l_last": l_last, "l_active": l_active, **ch}


def high_coefficients(case, h_ext):
    """Coefficients n * (degree - 1) .. 2^ext_k - 1 of h / (X^n - 1) on the coset (all zero iff the
    quotient is a polynomial of the degree `extended_to_coeff` keeps)."""
    d = case["domain"]
    qv = d.divide_by_vanishing_poly(h_ext)
    coeffs = h.best_fft(qv, d.extended_omega_inv, d.extended_k)
    return coeffs[d.n * d.quotient_poly_degree:]
