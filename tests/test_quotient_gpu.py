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


# ------------------------------------------------------------------------------------------------
# Flat graphs built by hand — not by the oracle's GraphEvaluator, which gives every calculation a fresh slot — against a
# pure-Python interpreter of the flat format: a slot written twice, values that stay live across many calculations,
# values nobody reads, a Horner whose parts are intermediates, a source and the target in the same slot, two Horners
# sharing a parts range.  (A liveness renumbering of the slots before upload was measured with this test in place:
# the RSA-shaped gate graph's 321 slots become 82, the kernel time does not change — 0.97 against 0.99 ms — so the
# interpreter is not bound by its thread-local array; not kept.)
_CONST, _INTER, _FIXED, _ADVICE, _INSTANCE, _CHALL, _BETA, _GAMMA, _THETA, _Y, _PREV = range(11)
_ADD, _SUB, _MUL, _SQUARE, _DOUBLE, _NEGATE, _HORNER, _STORE = range(8)


def _interpret_flat(calcs, parts, rotations, consts, fixed, advice, scal, prev, size, rot_scale):
    """Pure-Python evaluation of a flat graph at every index (the semantics of include/b200zk.h)."""
    p = bn.R
    out = []
    for idx in range(size):
        inter = {}

        def get(kind, a, b):
            if kind == _CONST: return consts[a]
            if kind == _INTER: return inter[a]
            if kind == _FIXED: return fixed[a][(idx + rotations[b] * rot_scale) % size]
            if kind == _ADVICE: return advice[a][(idx + rotations[b] * rot_scale) % size]
            if kind in (_BETA, _GAMMA, _THETA, _Y): return scal[kind]
            if kind == _PREV: return prev[idx]
            raise AssertionError(kind)

        last = 0
        for op, target, xk, xa, xb, yk, ya, yb, poff, plen in calcs:
            x = get(xk, xa, xb)
            if op in (_ADD, _SUB, _MUL, _HORNER):
                y = get(yk, ya, yb)
            if op == _ADD: v = (x + y) % p
            elif op == _SUB: v = (x - y) % p
            elif op == _MUL: v = x * y % p
            elif op == _SQUARE: v = x * x % p
            elif op == _DOUBLE: v = 2 * x % p
            elif op == _NEGATE: v = -x % p
            elif op == _HORNER:
                v = x
                for q in range(plen):
                    v = (v * y + get(*parts[poff + q])) % p
            else: v = x
            inter[target] = v
            last = v
        out.append(last)
    return out


def _hand_graph(shared_parts: bool):
    I = lambda a: (_INTER, a, 0)
    A = lambda c, r=0: (_ADVICE, c, r)
    Fx = lambda c, r=0: (_FIXED, c, r)
    Z = (_CONST, 0, 0)
    calcs, parts = [], []

    def calc(op, target, x, y=Z, ps=()):
        off = len(parts)
        parts.extend(ps)
        calcs.append((op, target, *x, *y, off, len(ps)))

    calc(_MUL, 0, A(0), A(1, 1))                 # slot 0, read much later (long live range)
    calc(_ADD, 1, A(2), Fx(0))
    calc(_SQUARE, 2, I(1))                        # slot 1 dies here
    calc(_MUL, 1, I(2), A(0, 2))                  # slot 1 written a second time
    calc(_NEGATE, 3, A(1))                        # never read
    calc(_SUB, 2, I(2), I(1))                     # source and target in the same original slot
    calc(_DOUBLE, 4, I(2))
    for j in range(40):                           # many short-lived temporaries between the uses of slot 0
        calc(_MUL, 5 + j, A(j % 3, j % 3), Fx(j % 2, 1))
        calc(_ADD, 4, I(4), I(5 + j))
    calc(_STORE, 50, (_BETA, 0, 0))
    calc(_HORNER, 51, (_PREV, 0, 0), (_Y, 0, 0), [I(0), I(4), A(2, 1), I(50), (_GAMMA, 0, 0), I(1)])
    if shared_parts:                              # a second Horner over the same parts range: the pass gives up
        off, ln = calcs[-1][-2], calcs[-1][-1]
        calcs.append((_HORNER, 52, *I(51), *(_THETA, 0, 0), off, ln))
    return np.array(calcs, dtype=np.uint32), np.array(parts, dtype=np.uint32).reshape(-1, 3), 53


@pytest.mark.parametrize("shared_parts", [False, True])
def test_hand_built_graph_with_slot_reuse_and_long_live_ranges(zk, shared_parts):
    import ctypes as C
    import random
    from b200zk.quotient import EnvC, _handles

    lib = zk.load()
    k, ext_k = 5, 7
    size, rot_scale = 1 << ext_k, 1 << (ext_k - k)
    rnd = random.Random(99)
    col = lambda: [rnd.randrange(bn.R) for _ in range(size)]
    fixed, advice, prev = [col() for _ in range(2)], [col() for _ in range(3)], col()
    scal = {_BETA: rnd.randrange(bn.R), _GAMMA: rnd.randrange(bn.R), _THETA: rnd.randrange(bn.R), _Y: rnd.randrange(bn.R)}
    consts = [7]
    rotations = [0, 1, -2]
    calcs, parts, n_inter = _hand_graph(shared_parts)
    want = _interpret_flat([tuple(int(v) for v in r) for r in calcs], [tuple(int(v) for v in r) for r in parts],
                           rotations, consts, fixed, advice, scal, prev, size, rot_scale)
    g = zk.FlatGraph(F(consts), np.array(rotations, dtype=np.int32), calcs, parts, n_inter)
    dfix, dadv = [zk.DeviceColumn.from_host(F(c)) for c in fixed], [zk.DeviceColumn.from_host(F(c)) for c in advice]
    dprev, dout = zk.DeviceColumn.from_host(F(prev)), zk.DeviceColumn(size)
    keep = [_handles(dfix), _handles(dadv), np.zeros(1, np.uint64), np.zeros((1, 4), np.uint64)]
    env = EnvC()
    env.fixed, env.n_fixed = keep[0].ctypes.data, len(dfix)
    env.advice, env.n_advice = keep[1].ctypes.data, len(dadv)
    env.instance, env.n_instance = keep[2].ctypes.data, 0
    env.challenges, env.n_challenges = keep[3].ctypes.data, 0
    for name, kind in (("beta", _BETA), ("gamma", _GAMMA), ("theta", _THETA), ("y", _Y)):
        getattr(env, name)[:] = [int(x) for x in F([scal[kind]])[0]]
    env.k, env.ext_k = k, ext_k
    env.range_begin, env.range_len = 0, 0
    gc = g.as_c()
    zk.check(lib.b200zk_quotient_graph(C.byref(gc), C.byref(env), dprev.handle, dout.handle))
    assert np.array_equal(dout.to_host(), F(want))
