"""GPU: the prover's lookup chain on the device (permute_expression_pair -> lookup grand product ->
lagrange_to_coeff -> evaluate_h), each stage fed by the previous device output, gives the oracle's
numerator bit for bit and a polynomial quotient — see tests/lookup_chain.py."""
import numpy as np
import pytest

import lookup_chain as lc
from oracle import bn254 as bn
from oracle import quotient_cpu as q
from test_lookup_chain_oracle import oracle_chain

pytestmark = pytest.mark.gpu
F = bn.fr_array_from_canonical


@pytest.mark.parametrize("k,seed", [(5, 3), (8, 9)])
def test_device_lookup_chain(zk, k, seed):
    case = lc.build(k=k, seed=seed)
    d = zk.EvaluationDomain(lc.DEGREE, k)
    one = lambda v: F([v])[0]
    ci = np.stack([F(c[0]) for c in case["compressed"]])
    ct = np.stack([F(c[1]) for c in case["compressed"]])
    pair_blinds = np.stack([np.stack([F(b["pair"][0]), F(b["pair"][1])]) for b in case["blinds"]])
    z_blinds = np.stack([F(b["z"]) for b in case["blinds"]])
    pa, pt = zk.permute_expression_pairs(ci, ct, k, lc.BF, pair_blinds)
    z = zk.lookup_products(list(ci), list(ct), list(pa), list(pt), k, one(case["beta"]), one(case["gamma"]), lc.BF, z_blinds)
    committed = [zk.LookupCommitted(d.lagrange_to_coeff(z[j].copy()), d.lagrange_to_coeff(pa[j].copy()),
                                    d.lagrange_to_coeff(pt[j].copy())) for j in range(len(case["lookups"]))]
    ext_col = lambda lagrange: zk.DeviceColumn.from_host(d.coeff_to_extended(d.lagrange_to_coeff(F(lagrange))))
    col = lambda ints: zk.DeviceColumn.from_host(F(ints))
    pk = zk.ProvingKeyCosets(fixed_cosets=[ext_col(c) for c in case["fixed"]], l0=col(case["l0"]), l_last=col(case["l_last"]),
                             l_active_row=col(case["l_active"]), permutation_cosets=[], permutation_columns=[],
                             degree=lc.DEGREE, blinding_factors=lc.BF)
    flat = lambda g: zk.FlatGraph(**g.to_flat())
    ev = zk.Evaluator(flat(q.custom_gates_graph(case["gates"])), [flat(q.lookup_graph(i, t)) for i, t in case["lookups"]])
    h_dev = ev.evaluate_h(d, pk, [d.lagrange_to_coeff(F(c)) for c in case["advice"]], [], np.zeros((0, 4), np.uint64),
                          one(case["y"]), one(case["beta"]), one(case["gamma"]), one(case["theta"]), committed, [])
    h_ref = oracle_chain(case)
    assert np.array_equal(h_dev, F(h_ref))
    assert not any(lc.high_coefficients(case, bn.fr_array_to_canonical(h_dev)))
