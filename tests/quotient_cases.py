"""Shared builders for the quotient tests: circuits as expression lists + random columns."""
from __future__ import annotations

import random

import numpy as np

from oracle import bn254 as bn
from oracle import halo2_cpu as h
from oracle import quotient_cpu as q

R = bn.R


def square_circuit_gates():
    """reference src/signal.rs:36-42: s * (b - a^2) with a = advice 0, b = advice 1 (the
    selector is fixed column 0 after selector compression)."""
    a0, a1, f0 = q.Advice(0), q.Advice(1), q.Fixed(0)
    return [f0 * (a1 - a0 * a0)]


def rand_cols(rnd, count, size):
    return [[rnd.randrange(R) for _ in range(size)] for _ in range(count)]


def lagrange_cols(domain, blinding_factors):
    """l_0, l_last, l_active_row on the extended coset (keygen.rs): l_last is the Lagrange
    polynomial of row n - blinding_factors - 1, l_blind the sum over the last
    blinding_factors rows, l_active = 1 - (l_last + l_blind)."""
    n = domain.n
    def ext_of_rows(rows):
        lag = [1 if i in rows else 0 for i in range(n)]
        return domain.coeff_to_extended(domain.lagrange_to_coeff(lag))
    l0 = ext_of_rows({0})
    l_last = ext_of_rows({n - blinding_factors - 1})
    l_blind = ext_of_rows(set(range(n - blinding_factors, n)))
    l_active = [(1 - a - b) % R for a, b in zip(l_last, l_blind)]
    return l0, l_last, l_blind, l_active


def square_case(k=4, seed=1):
    """SquareCircuit shape: 1 fixed, 2 advice, 1 instance, permutation over
    [advice0, advice1, instance0], degree 3 (chunk 1, 3 sets), blinding_factors 5."""
    rnd = random.Random(seed)
    domain = h.EvaluationDomain(3, k)      # EvaluationDomain::new(cs.degree(), k): ext_k = k + 1, two quotient pieces (contract.sol:11-12)
    N = domain.extended_n
    bf = 5
    l0, l_last, l_blind, l_active = lagrange_cols(domain, bf)
    case = {
        "domain": domain, "k": k, "degree": 3, "blinding_factors": bf,
        "gates": square_circuit_gates(), "lookups": [],
        "fixed": rand_cols(rnd, 1, N),
        "advice_coeff": rand_cols(rnd, 2, domain.n), "instance_coeff": rand_cols(rnd, 1, domain.n),
        "challenges": [],
        "perm_columns": [("advice", 0), ("advice", 1), ("instance", 0)],
        "sigma": rand_cols(rnd, 3, N), "product_coeff": rand_cols(rnd, 3, domain.n),
        "l0": l0, "l_last": l_last, "l_blind": l_blind, "l_active": l_active,
        "beta": rnd.randrange(R), "gamma": rnd.randrange(R), "theta": rnd.randrange(R), "y": rnd.randrange(R),
    }
    return case


def wide_case(k=6, seed=2, n_advice=6, n_fixed=4, n_lookups=2):
    """A synthetic wide circuit: many gates with rotations, challenges, scaled / negated /
    constant sub-expressions (every add_expression branch), a 7-column permutation with
    chunk_len 3 (degree 5) and two lookups."""
    rnd = random.Random(seed)
    degree = 5
    domain = h.EvaluationDomain(degree, k)
    N = domain.extended_n
    bf = 5
    A = lambda i, r=0: q.Advice(i % n_advice, r)
    F = lambda i, r=0: q.Fixed(i % n_fixed, r)
    I = lambda r=0: q.Instance(0, r)
    gates = [
        F(0) * (A(0) + A(1) * A(2) - A(3)),
        F(1) * (A(0, 1) - A(0) * q.Const(2)),
        F(2) * (A(4, -1) * A(4, -1) - A(5, 2)) + q.Const(0) * A(1),
        (A(1) * q.Const(1)) * (A(2) * 7 - q.Const(5)),
        -(A(3) * F(3, 1)) + q.Challenge(0) * A(2, 3),
        F(0) * (A(0) + A(1) * A(2) - A(3)),            # duplicate gate: dedup path
        (A(5) - I()) * (q.Const(0) - A(1)) + (q.Const(0) + A(2)) * 0,
        q.Const(3) + q.Challenge(1) * q.Challenge(0) - (-q.Const(4)),
    ]
    lookups = []
    for j in range(n_lookups):
        inputs = [F(j) * A(j), A(j + 1, 1) + q.Const(j + 1)]
        tables = [F(j + 1), F(j + 2, -1)]
        lookups.append((inputs, tables))
    l0, l_last, l_blind, l_active = lagrange_cols(domain, bf)
    perm_cols = [("advice", 0), ("advice", 1), ("fixed", 1), ("advice", 3), ("instance", 0), ("advice", 5), ("fixed", 0)]
    n_sets = -(-len(perm_cols) // (degree - 2))
    return {
        "domain": domain, "k": k, "degree": degree, "blinding_factors": bf,
        "gates": gates, "lookups": lookups,
        "fixed": rand_cols(rnd, n_fixed, N),
        "advice_coeff": rand_cols(rnd, n_advice, domain.n), "instance_coeff": rand_cols(rnd, 1, domain.n),
        "challenges": [rnd.randrange(R) for _ in range(2)],
        "perm_columns": perm_cols,
        "sigma": rand_cols(rnd, len(perm_cols), N), "product_coeff": rand_cols(rnd, n_sets, domain.n),
        "lookup_coeff": [tuple(rand_cols(rnd, 3, domain.n)) for _ in range(n_lookups)],
        "l0": l0, "l_last": l_last, "l_blind": l_blind, "l_active": l_active,
        "beta": rnd.randrange(R), "gamma": rnd.randrange(R), "theta": rnd.randrange(R), "y": rnd.randrange(R),
    }


def oracle_evaluate_h(case):
    """Run oracle/quotient_cpu.evaluate_h on a case; returns (values, env, graphs)."""
    d = case["domain"]
    env = {
        "fixed": case["fixed"],
        "advice": [d.coeff_to_extended(c) for c in case["advice_coeff"]],
        "instance": [d.coeff_to_extended(c) for c in case["instance_coeff"]],
        "challenges": case["challenges"],
        "beta": case["beta"], "gamma": case["gamma"], "theta": case["theta"], "y": case["y"],
    }
    gates = q.custom_gates_graph(case["gates"])
    perm = q.PermutationData(case["perm_columns"], case["sigma"],
                             [d.coeff_to_extended(c) for c in case["product_coeff"]],
                             case["degree"] - 2, case["blinding_factors"])
    lookups = []
    lgraphs = []
    for (inputs, tables), coeffs in zip(case["lookups"], case.get("lookup_coeff", [])):
        g = q.lookup_graph(inputs, tables)
        lgraphs.append(g)
        lookups.append(q.LookupData(g, *[d.coeff_to_extended(c) for c in coeffs]))
    values = q.evaluate_h(d, gates, env, case["l0"], case["l_last"], case["l_active"], perm, lookups)
    return values, env, gates, lgraphs
