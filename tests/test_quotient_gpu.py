"""GPU parity of evaluate_h (graph interpreter + permutation + lookup kernels) through
the C ABI against the oracle, bit-exact."""
import numpy as np
import pytest

from oracle import bn254 as bn
from quotient_cases import oracle_evaluate_h, square_case, wide_case

pytestmark = pytest.mark.gpu

F = bn.fr_array_from_canonical


def run_device(zk, case, gates, lgraphs):
    d = zk.EvaluationDomain(case["degree"], case["k"])
    col = lambda ints: zk.DeviceColumn.from_host(F(ints))
    flat = lambda g: zk.FlatGraph(**g.to_flat())
    pk = zk.ProvingKeyCosets(
        fixed_cosets=[col(c) for c in case["fixed"]], l0=col(case["l0"]), l_last=col(case["l_last"]),
        l_active_row=col(case["l_active"]), permutation_cosets=[col(c) for c in case["sigma"]],
        permutation_columns=case["perm_columns"], degree=case["degree"], blinding_factors=case["blinding_factors"])
    ev = zk.Evaluator(flat(gates), [flat(g) for g in lgraphs])
    lookups = [zk.LookupCommitted(*[F(c) for c in coeffs]) for coeffs in case.get("lookup_coeff", [])]
    ch = F(case["challenges"]) if case["challenges"] else np.zeros((0, 4), dtype=np.uint64)
    one = lambda v: F([v])[0]
    out = ev.evaluate_h(d, pk, [F(c) for c in case["advice_coeff"]], [F(c) for c in case["instance_coeff"]], ch,
                        one(case["y"]), one(case["beta"]), one(case["gamma"]), one(case["theta"]), lookups,
                        [F(c) for c in case["product_coeff"]])
    return out


@pytest.mark.parametrize("k", [4, 7])
def test_square_circuit_shape(zk, k):
    case = square_case(k=k)
    values, _, gates, lgraphs = oracle_evaluate_h(case)
    assert np.array_equal(run_device(zk, case, gates, lgraphs), F(values))


@pytest.mark.parametrize("k,seed", [(5, 2), (7, 3)])
def test_wide_circuit_with_lookups_and_multi_set_permutation(zk, k, seed):
    case = wide_case(k=k, seed=seed)
    values, _, gates, lgraphs = oracle_evaluate_h(case)
    assert np.array_equal(run_device(zk, case, gates, lgraphs), F(values))


def test_bad_graph_is_rejected(zk):
    case = square_case(k=4)
    _, _, gates, lgraphs = oracle_evaluate_h(case)
    flat = gates.to_flat()
    flat["calcs"] = flat["calcs"].copy()
    flat["calcs"][0, 3] = 99            # advice column index out of range
    with pytest.raises(zk.B200zkError, match="out of range"):
        run_device(zk, case, type("G", (), {"to_flat": lambda self: flat})(), lgraphs)


def test_sharded_extended_domain_equals_whole(zk):
    """Multi-GPU h(X): per-rank index ranges (here evaluated one after the other on one
    GPU) concatenate to the single-GPU result."""
    from b200zk.sharding import shard_range

    case = wide_case(k=6, seed=5)
    values, _, gates, lgraphs = oracle_evaluate_h(case)
    d = zk.EvaluationDomain(case["degree"], case["k"])
    col = lambda ints: zk.DeviceColumn.from_host(F(ints))
    flat = lambda g: zk.FlatGraph(**g.to_flat())
    pk = zk.ProvingKeyCosets(
        fixed_cosets=[col(c) for c in case["fixed"]], l0=col(case["l0"]), l_last=col(case["l_last"]),
        l_active_row=col(case["l_active"]), permutation_cosets=[col(c) for c in case["sigma"]],
        permutation_columns=case["perm_columns"], degree=case["degree"], blinding_factors=case["blinding_factors"])
    ev = zk.Evaluator(flat(gates), [flat(g) for g in lgraphs])
    lookups = [zk.LookupCommitted(*[F(c) for c in coeffs]) for coeffs in case["lookup_coeff"]]
    one = lambda v: F([v])[0]
    world = 3
    parts = []
    for r in range(world):
        rng = shard_range(d.extended_len(), r, world)
        parts.append(ev.evaluate_h(d, pk, [F(c) for c in case["advice_coeff"]], [F(c) for c in case["instance_coeff"]],
                                   F(case["challenges"]), one(case["y"]), one(case["beta"]), one(case["gamma"]),
                                   one(case["theta"]), lookups, [F(c) for c in case["product_coeff"]], idx_range=rng))
    assert np.array_equal(np.concatenate(parts), F(values))
