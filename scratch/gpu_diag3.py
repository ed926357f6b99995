import sys
sys.path.insert(0, "."); sys.path.insert(0, "anon-aadhaar-halo2_b200"); sys.path.insert(0, "tests")
import numpy as np, ctypes as C
import b200zk
from b200zk.quotient import *
from b200zk.api import _ptr, fr_limbs
from oracle import bn254 as bn, quotient_cpu as q
from quotient_cases import *
F = bn.fr_array_from_canonical
b200zk.init(0)
lib = b200zk.load()
case = square_case(k=4)
d_o = case["domain"]
full, env_o, gates, _ = oracle_evaluate_h(case)
# gates only oracle
size = d_o.extended_n; rs = 1 << (d_o.extended_k - d_o.k)
gates_only = [gates.evaluate(env_o, i, rs, size, 0) for i in range(size)]
d = b200zk.EvaluationDomain(4, 4)
col = lambda ints: DeviceColumn.from_host(F(ints))
ev = Evaluator(FlatGraph(**gates.to_flat()))
advice = ev._extend(d, [F(c) for c in case["advice_coeff"]])
inst = ev._extend(d, [F(c) for c in case["instance_coeff"]])
for i, a in enumerate(advice):
    print("advice ext", i, np.array_equal(a.to_host(), F(env_o["advice"][i])))
print("inst ext", np.array_equal(inst[0].to_host(), F(env_o["instance"][0])))
fixed = [col(c) for c in case["fixed"]]
one = lambda v: F([v])[0]
env = ev._env(d, fixed, advice, inst, np.zeros((0,4),np.uint64), one(case["beta"]), one(case["gamma"]), one(case["theta"]), one(case["y"]))
values = DeviceColumn.from_host(np.zeros((size,4),np.uint64))
g = ev.custom_gates.as_c()
print("graph:", gates.to_flat())
b200zk.check(lib.b200zk_quotient_graph(C.byref(g), C.byref(env), values.handle, values.handle))
got = values.to_host()
print("gates only:", np.array_equal(got, F(gates_only)))
if not np.array_equal(got, F(gates_only)):
    print(bn.fr_array_to_canonical(got[:2]), gates_only[:2])
    # try each intermediate: evaluate sub-graphs
    for ncalc in range(1, len(gates.calculations)+1):
        sub = q.GraphEvaluator(); sub.constants = gates.constants; sub.rotations = gates.rotations
        sub.calculations = gates.calculations[:ncalc]; sub.num_intermediates = gates.num_intermediates
        exp = [sub.evaluate(env_o, i, rs, size, 0) for i in range(size)]
        fg = FlatGraph(**sub.to_flat()); gc = fg.as_c()
        out = DeviceColumn(size)
        b200zk.check(lib.b200zk_quotient_graph(C.byref(gc), C.byref(env), 0, out.handle))
        print("  prefix", ncalc, gates.calculations[ncalc-1], np.array_equal(out.to_host(), F(exp)))
