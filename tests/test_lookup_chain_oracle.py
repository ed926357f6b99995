"""CPU: the oracle's lookup chain is self-consistent — see tests/lookup_chain.py."""
import lookup_chain as lc
from oracle import prover_steps_cpu as ps
from oracle import quotient_cpu as q


def oracle_chain(case):
    d = case["domain"]
    ext = lambda lagrange: d.coeff_to_extended(d.lagrange_to_coeff(lagrange))
    env = {"fixed": [ext(c) for c in case["fixed"]], "advice": [ext(c) for c in case["advice"]], "instance": [],
           "challenges": [], "beta": case["beta"], "gamma": case["gamma"], "theta": case["theta"], "y": case["y"]}
    data = []
    for (inputs, tables), (ci, ct), bl in zip(case["lookups"], case["compressed"], case["blinds"]):
        pa, ps_ = ps.permute_expression_pair(ci, ct, lc.BF, blinds=bl["pair"])
        z = ps.lookup_product(ci, ct, pa, ps_, case["beta"], case["gamma"], lc.BF, blinds=bl["z"])
        data.append(q.LookupData(q.lookup_graph(inputs, tables), ext(z), ext(pa), ext(ps_)))
    return q.evaluate_h(d, q.custom_gates_graph(case["gates"]), env, case["l0"], case["l_last"], case["l_active"],
                        None, data)


def test_satisfying_witness_gives_a_polynomial_quotient():
    for k, seed in ((5, 3), (6, 4)):
        case = lc.build(k=k, seed=seed)
        h_ext = oracle_chain(case)
        assert any(h_ext)
        assert not any(lc.high_coefficients(case, h_ext))


def test_inconsistent_witness_does_not():
    case = lc.build(k=5, seed=3, corrupt=True)
    assert any(lc.high_coefficients(case, oracle_chain(case)))
