"""CPU restatement (Python big integers) of halo2_proofs' quotient numerator evaluation.

TEST INFRASTRUCTURE ONLY — see the header of ``oracle/bn254.py``.  PARITY PIN: no upstream
output vectors exist; the gate and permutation parts of evaluate_h are pinned end to end by the
reference's verifier contract accepting the proofs whose h(X) this file computes
(tests/test_square_proof_oracle.py), the lookup part only by its definition.  The structural
pin the reference holds is the quotient identity of the SquareCircuit shape in
``solidity_verifier_contract/contract.sol:443-505`` (gate ``f_0 * (a_1 - a_0^2)``, the
``l_0 / l_last / l_blind`` permutation terms and their y-folding order), which
``tests/test_quotient_oracle.py`` checks this restatement against.

Upstream being restated ([DEP] halo2_proofs 0.2.0 @ v2023_01_20, reference
``Cargo.lock:469-471``): ``halo2_proofs/src/plonk/evaluation.rs``
  * ``ValueSource`` / ``Calculation`` / ``GraphEvaluator::{add_expression, evaluate}``
  * ``Evaluator::evaluate_h``  (custom gates, permutation argument, lookup argument)
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import bn254 as bn

R = bn.R

# ValueSource variants in upstream declaration order (the derived ordering is used to
# canonicalise commutative operands) — these numbers are also the C ABI encoding
# (include/b200zk.h `b200zk_src.kind`).
CONSTANT, INTERMEDIATE, FIXED, ADVICE, INSTANCE, CHALLENGE, BETA, GAMMA, THETA, Y, PREVIOUS = range(11)
# Calculation variants (C ABI `b200zk_calc.op`)
ADD, SUB, MUL, SQUARE, DOUBLE, NEGATE, HORNER, STORE = range(8)


def src(kind, a=0, b=0):
    return (kind, a, b)


# ------------------------------------------------------------------ expressions (plonk/circuit.rs)
@dataclass(frozen=True)
class Expr:
    op: str
    args: tuple = ()

    def __add__(self, o): return Expr("sum", (self, o))
    def __sub__(self, o): return Expr("sum", (self, Expr("neg", (o,))))
    def __mul__(self, o):
        if isinstance(o, int):
            return Expr("scaled", (self, o % R))
        return Expr("product", (self, o))
    def __neg__(self): return Expr("neg", (self,))


def Const(v): return Expr("const", (v % R,))
def Fixed(col, rot=0): return Expr("fixed", (col, rot))
def Advice(col, rot=0): return Expr("advice", (col, rot))
def Instance(col, rot=0): return Expr("instance", (col, rot))
def Challenge(i): return Expr("challenge", (i,))


def eval_expr(e: Expr, get) -> int:
    """Direct evaluation of an expression tree (the *definition* the graph must match)."""
    if e.op == "const": return e.args[0]
    if e.op in ("fixed", "advice", "instance"): return get(e.op, e.args[0], e.args[1])
    if e.op == "challenge": return get("challenge", e.args[0], 0)
    if e.op == "neg": return (-eval_expr(e.args[0], get)) % R
    if e.op == "sum": return (eval_expr(e.args[0], get) + eval_expr(e.args[1], get)) % R
    if e.op == "product": return eval_expr(e.args[0], get) * eval_expr(e.args[1], get) % R
    if e.op == "scaled": return eval_expr(e.args[0], get) * e.args[1] % R
    raise ValueError(e.op)


# ------------------------------------------------------------------ GraphEvaluator
class GraphEvaluator:
    """evaluation.rs GraphEvaluator: constants [0, 1, 2, ...], deduplicated rotations and
    calculations, each calculation writing one intermediate."""

    def __init__(self):
        self.constants = [0, 1, 2]
        self.rotations = []
        self.calculations = []     # (op, target, x, y, parts)
        self.num_intermediates = 0

    def add_rotation(self, rot: int) -> int:
        if rot in self.rotations:
            return self.rotations.index(rot)
        self.rotations.append(rot)
        return len(self.rotations) - 1

    def add_constant(self, c: int):
        c %= R
        if c in self.constants:
            return src(CONSTANT, self.constants.index(c))
        self.constants.append(c)
        return src(CONSTANT, len(self.constants) - 1)

    def add_calculation(self, op, x=src(CONSTANT), y=src(CONSTANT), parts=()):
        key = (op, x, y, tuple(parts))
        for c in self.calculations:
            if (c[0], c[2], c[3], c[4]) == key:
                return src(INTERMEDIATE, c[1])
        target = self.num_intermediates
        self.calculations.append((op, target, x, y, tuple(parts)))
        self.num_intermediates += 1
        return src(INTERMEDIATE, target)

    def add_expression(self, e: Expr):
        zero, one, two = src(CONSTANT, 0), src(CONSTANT, 1), src(CONSTANT, 2)
        if e.op == "const":
            return self.add_constant(e.args[0])
        if e.op in ("fixed", "advice", "instance"):
            kind = {"fixed": FIXED, "advice": ADVICE, "instance": INSTANCE}[e.op]
            r = self.add_rotation(e.args[1])
            return self.add_calculation(STORE, src(kind, e.args[0], r))
        if e.op == "challenge":
            return self.add_calculation(STORE, src(CHALLENGE, e.args[0]))
        if e.op == "neg":
            a = e.args[0]
            if a.op == "const":
                return self.add_constant(-a.args[0])
            ra = self.add_expression(a)
            return ra if ra == zero else self.add_calculation(NEGATE, ra)
        if e.op == "sum":
            a, b = e.args
            if b.op == "neg":          # undo a + (-b)
                ra, rb = self.add_expression(a), self.add_expression(b.args[0])
                if ra == zero:
                    return self.add_calculation(NEGATE, rb)
                if rb == zero:
                    return ra
                return self.add_calculation(SUB, ra, rb)
            ra, rb = self.add_expression(a), self.add_expression(b)
            if ra == zero:
                return rb
            if rb == zero:
                return ra
            return self.add_calculation(ADD, *sorted((ra, rb)))
        if e.op == "product":
            ra, rb = self.add_expression(e.args[0]), self.add_expression(e.args[1])
            if ra == zero or rb == zero:
                return zero
            if ra == one:
                return rb
            if rb == one:
                return ra
            if ra == two:
                return self.add_calculation(DOUBLE, rb)
            if rb == two:
                return self.add_calculation(DOUBLE, ra)
            if ra == rb:
                return self.add_calculation(SQUARE, ra)
            return self.add_calculation(MUL, *sorted((ra, rb)))
        if e.op == "scaled":
            a, f = e.args
            if f == 0:
                return zero
            if f == 1:
                return self.add_expression(a)
            cst = self.add_constant(f)
            ra = self.add_expression(a)
            return self.add_calculation(MUL, ra, cst)
        raise ValueError(e.op)

    # -- evaluation.rs GraphEvaluator::evaluate
    def evaluate(self, env, idx: int, rot_scale: int, isize: int, previous: int) -> int:
        rots = [(idx + r * rot_scale) % isize for r in self.rotations]
        inter = [0] * self.num_intermediates

        def get(s):
            kind, a, b = s
            if kind == CONSTANT: return self.constants[a]
            if kind == INTERMEDIATE: return inter[a]
            if kind == FIXED: return env["fixed"][a][rots[b]]
            if kind == ADVICE: return env["advice"][a][rots[b]]
            if kind == INSTANCE: return env["instance"][a][rots[b]]
            if kind == CHALLENGE: return env["challenges"][a]
            if kind == BETA: return env["beta"]
            if kind == GAMMA: return env["gamma"]
            if kind == THETA: return env["theta"]
            if kind == Y: return env["y"]
            if kind == PREVIOUS: return previous
            raise ValueError(kind)

        for (op, target, x, y, parts) in self.calculations:
            if op == ADD: v = (get(x) + get(y)) % R
            elif op == SUB: v = (get(x) - get(y)) % R
            elif op == MUL: v = get(x) * get(y) % R
            elif op == SQUARE: v = get(x) * get(x) % R
            elif op == DOUBLE: v = 2 * get(x) % R
            elif op == NEGATE: v = (-get(x)) % R
            elif op == STORE: v = get(x)
            elif op == HORNER:
                v = get(x)
                f = get(y)
                for p in parts:
                    v = (v * f + get(p)) % R
            else:
                raise ValueError(op)
            inter[target] = v
        return inter[self.calculations[-1][1]] if self.calculations else 0

    # -- flat encoding for the C ABI (include/b200zk.h: b200zk_src / b200zk_calc)
    def to_flat(self):
        parts_flat = []
        calcs = np.zeros((len(self.calculations), 10), dtype=np.uint32)
        for i, (op, target, x, y, parts) in enumerate(self.calculations):
            off = len(parts_flat)
            parts_flat += list(parts)
            calcs[i] = [op, target, x[0], x[1], x[2], y[0], y[1], y[2], off, len(parts)]
        parts_arr = np.array(parts_flat, dtype=np.uint32).reshape(-1, 3) if parts_flat else np.zeros((0, 3), np.uint32)
        return {
            "constants": bn.fr_array_from_canonical(self.constants),
            "rotations": np.array(self.rotations, dtype=np.int32),
            "calcs": calcs,
            "parts": parts_arr,
            "num_intermediates": self.num_intermediates,
        }


def custom_gates_graph(gate_polys) -> GraphEvaluator:
    """Evaluator::new: all gate polynomials folded by Horner in y over PreviousValue."""
    g = GraphEvaluator()
    parts = [g.add_expression(p) for p in gate_polys]
    g.add_calculation(HORNER, src(PREVIOUS), src(Y), parts)
    return g


def lookup_graph(input_exprs, table_exprs) -> GraphEvaluator:
    """Evaluator::new, lookup part: (theta-compressed input + beta) * (compressed table + gamma)."""
    g = GraphEvaluator()

    def compress(exprs):
        parts = [g.add_expression(e) for e in exprs]
        return g.add_calculation(HORNER, src(CONSTANT, 0), src(THETA), parts)

    ci, ct = compress(input_exprs), compress(table_exprs)
    right_gamma = g.add_calculation(ADD, ct, src(GAMMA))
    lc = g.add_calculation(ADD, ci, src(BETA))
    g.add_calculation(MUL, lc, right_gamma)
    return g


# ------------------------------------------------------------------ evaluate_h
@dataclass
class PermutationData:
    columns: list              # [(kind, index)] in cs.permutation.columns order, kind in fixed/advice/instance
    sigma_cosets: list         # pk.permutation.cosets, one extended column per permutation column
    product_cosets: list       # permutation.sets[i].permutation_product_coset
    chunk_len: int             # cs.degree() - 2
    blinding_factors: int


@dataclass
class LookupData:
    graph: GraphEvaluator
    product_coset: list
    permuted_input_coset: list
    permuted_table_coset: list


def evaluate_h(domain, gates: GraphEvaluator, env, l0, l_last, l_active_row,
               permutation: PermutationData | None, lookups=()):
    """Evaluator::evaluate_h for one circuit instance.  ``env`` holds the *extended*
    fixed / advice / instance columns (lists of ints), challenges, beta, gamma, theta, y."""
    size = domain.extended_n
    rot_scale = 1 << (domain.extended_k - domain.k)
    y, beta, gamma = env["y"], env["beta"], env["gamma"]
    values = [0] * size
    for idx in range(size):
        values[idx] = gates.evaluate(env, idx, rot_scale, size, values[idx])

    if permutation is not None and permutation.product_cosets:
        sets = permutation.product_cosets
        last_rot = -(permutation.blinding_factors + 1)
        delta_start = beta * bn.FR_ZETA % R
        colvals = [env[kind][i] for (kind, i) in permutation.columns]
        for idx in range(size):
            r_next = (idx + rot_scale) % size
            r_last = (idx + last_rot * rot_scale) % size
            v = values[idx]
            v = (v * y + (1 - sets[0][idx]) * l0[idx]) % R
            zl = sets[-1][idx]
            v = (v * y + (zl * zl - zl) * l_last[idx]) % R
            for s in range(1, len(sets)):
                v = (v * y + (sets[s][idx] - sets[s - 1][r_last]) * l0[idx]) % R
            current_delta = delta_start * pow(domain.extended_omega, idx, R) % R
            for s, zset in enumerate(sets):
                lo, hi = s * permutation.chunk_len, min((s + 1) * permutation.chunk_len, len(colvals))
                left = zset[r_next]
                for c in range(lo, hi):
                    left = left * (colvals[c][idx] + beta * permutation.sigma_cosets[c][idx] + gamma) % R
                right = zset[idx]
                for c in range(lo, hi):
                    right = right * (colvals[c][idx] + current_delta + gamma) % R
                    current_delta = current_delta * bn.FR_DELTA % R
                v = (v * y + (left - right) * l_active_row[idx]) % R
            values[idx] = v

    for lk in lookups:
        for idx in range(size):
            table_value = lk.graph.evaluate(env, idx, rot_scale, size, 0)
            r_next = (idx + rot_scale) % size
            r_prev = (idx - rot_scale) % size
            z, a, s = lk.product_coset, lk.permuted_input_coset, lk.permuted_table_coset
            a_minus_s = (a[idx] - s[idx]) % R
            v = values[idx]
            v = (v * y + (1 - z[idx]) * l0[idx]) % R
            v = (v * y + (z[idx] * z[idx] - z[idx]) * l_last[idx]) % R
            v = (v * y + (z[r_next] * (a[idx] + beta) % R * (s[idx] + gamma) - z[idx] * table_value) * l_active_row[idx]) % R
            v = (v * y + a_minus_s * l0[idx]) % R
            v = (v * y + a_minus_s * (a[idx] - a[r_prev]) % R * l_active_row[idx]) % R
            values[idx] = v
    return values
