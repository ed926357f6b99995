"""The quotient restatement against its pins: the expression-tree definition and the
SquareCircuit identity in the reference's Solidity verifier."""
import random

from oracle import bn254 as bn
from oracle import quotient_cpu as q
from quotient_cases import oracle_evaluate_h, square_case, wide_case

R = bn.R


def test_graph_evaluation_equals_expression_definition():
    case = wide_case(k=4)
    values, env, gates, lgraphs = oracle_evaluate_h({**case, "perm_columns": [], "sigma": [], "product_coeff": [],
                                                     "lookups": [], "lookup_coeff": []})
    d = case["domain"]
    size, rs = d.extended_n, 1 << (d.extended_k - d.k)
    for idx in range(0, size, 7):
        def get(kind, col, rot):
            if kind == "challenge":
                return env["challenges"][col]
            return env[kind][col][(idx + rot * rs) % size]
        acc = 0
        for poly in case["gates"]:                 # values = fold over gates of (acc * y + poly)
            acc = (acc * case["y"] + q.eval_expr(poly, get)) % R
        assert values[idx] == acc
    assert gates.constants[:3] == [0, 1, 2]
    assert len(gates.calculations) == gates.num_intermediates


def test_lookup_graph_is_theta_compression():
    case = wide_case(k=4)
    d = case["domain"]
    _, env, _, lgraphs = oracle_evaluate_h(case)
    size, rs = d.extended_n, 1 << (d.extended_k - d.k)
    for (inputs, tables), g in zip(case["lookups"], lgraphs):
        for idx in (0, 5, size - 1):
            get = lambda kind, col, rot: env[kind][col][(idx + rot * rs) % size]
            comp = lambda exprs: sum(q.eval_expr(e, get) * pow(case["theta"], len(exprs) - 1 - i, R)
                                     for i, e in enumerate(exprs)) % R
            exp = (comp(inputs) + case["beta"]) * (comp(tables) + case["gamma"]) % R
            assert g.evaluate(env, idx, rs, size, 0) == exp


def contract_quotient_numer(ev, y, beta, gamma, x, l_0, l_last, l_blind):
    """Transliteration of reference solidity_verifier_contract/contract.sol:439-505.
    `ev` holds the evaluations the contract reads from calldata."""
    delta = 4131629893567559867359510883348571134090853742863529169391034518566172092834  # :440
    num = ev["f_0"] * ((ev["a_1"] - ev["a_0"] * ev["a_0"]) % R) % R                         # :442-450
    num = (num * y + (l_0 - l_0 * ev["z0"])) % R                                            # :452-456
    num = (num * y + l_last * (ev["z2"] * ev["z2"] - ev["z2"])) % R                         # :457-461
    num = (num * y + l_0 * (ev["z1"] - ev["z0_last"])) % R                                  # :462-465
    num = (num * y + l_0 * (ev["z2"] - ev["z1_last"])) % R                                  # :466-469
    cur = beta * x % R
    for (z_next, z, val, sigma) in ((ev["z0_next"], ev["z0"], ev["a_0"], ev["s0"]),
                                    (ev["z1_next"], ev["z1"], ev["a_1"], ev["s1"]),
                                    (ev["z2_next"], ev["z2"], ev["inst"], ev["s2"])):        # :470-505
        lhs = z_next * ((val + beta * sigma + gamma) % R) % R
        rhs = z * ((val + cur + gamma) % R) % R
        cur = cur * delta % R
        lsr = (lhs - rhs) % R
        num = (num * y + (lsr - lsr * ((l_last + l_blind) % R))) % R
    return num


def test_square_circuit_identity_matches_reference_contract():
    case = square_case(k=4)
    values, env, gates, _ = oracle_evaluate_h(case)
    d = case["domain"]
    size, rs = d.extended_n, 1 << (d.extended_k - d.k)
    z = [d.coeff_to_extended(c) for c in case["product_coeff"]]
    last = -(case["blinding_factors"] + 1)      # contract rotation omega^-6 (:544-550)
    assert last == -6
    for idx in range(size):
        nx, ls = (idx + rs) % size, (idx + last * rs) % size
        ev = {
            "f_0": env["fixed"][0][idx], "a_0": env["advice"][0][idx], "a_1": env["advice"][1][idx],
            "inst": env["instance"][0][idx],
            "s0": case["sigma"][0][idx], "s1": case["sigma"][1][idx], "s2": case["sigma"][2][idx],
            "z0": z[0][idx], "z0_next": z[0][nx], "z0_last": z[0][ls],
            "z1": z[1][idx], "z1_next": z[1][nx], "z1_last": z[1][ls],
            "z2": z[2][idx], "z2_next": z[2][nx],
        }
        x = bn.FR_ZETA * pow(d.extended_omega, idx, R) % R
        exp = contract_quotient_numer(ev, case["y"], case["beta"], case["gamma"], x, case["l0"][idx],
                                      case["l_last"][idx], case["l_blind"][idx])
        assert values[idx] == exp, idx
