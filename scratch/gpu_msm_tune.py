"""MSM stage timings for several k and chunk caps."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "anon-aadhaar-halo2_b200")
import numpy as np, ctypes as C, torch
import b200zk
from b200zk.api import _ptr
from oracle import bn254 as bn, c_oracle as co
b200zk.init(0); lib = b200zk.load()
names = ["hist", "scan", "scatter", "sync", "accum", "combine", "reduce", "red_comb", "fold"]
stream = torch.cuda.Stream(); st = C.c_void_p(stream.cuda_stream)
b200zk.check(lib.b200zk_msm_profile(1))
ks = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [15, 16, 18, 20, 22, 24]
caps = [tuple(int(y) for y in x.split(":")) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [(128, 32, 0), (128, 128, 0)]
for k in ks:
    n = 1 << k
    ds = torch.empty(n * 4, dtype=torch.int64, device="cuda"); db = torch.empty(n * 8, dtype=torch.int64, device="cuda")
    b200zk.check(lib.b200zk_gen_scalars_dev(C.c_void_p(ds.data_ptr()), n, 0xA11CE000 + k, 0))
    b200zk.check(lib.b200zk_gen_points_dev(C.c_void_p(db.data_ptr()), n, 0xBA5E0000 + k, 0))
    out = np.zeros(12, dtype=np.uint64); ref = None
    for cap in caps:
        b200zk.check(lib.b200zk_msm_tune(*cap))
        with torch.cuda.stream(stream):
            def run(): b200zk.check(lib.b200zk_msm_g1_dev(C.c_void_p(ds.data_ptr()), C.c_void_p(db.data_ptr()), n, _ptr(out), st))
            run(); torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream); run(); run(); run(); e1.record(stream); torch.cuda.synchronize()
        ms = (C.c_float * 9)(); info = (C.c_uint64 * 5)()
        b200zk.check(lib.b200zk_msm_last_stages(ms, 9, info))
        aff = bn.g1_jacobian_limbs_to_affine(out)
        if ref is None: ref = aff
        print(f"k={k} cap={cap} total={e0.elapsed_time(e1)/3:.3f}ms c={info[1]} W={info[2]} pairs={info[3]} L={info[4]} same={aff==ref} | " + " ".join(f"{nm}={v:.3f}" for nm, v in zip(names, ms)), flush=True)
    if k <= 16:
        exp = bn.g1_jacobian_limbs_to_affine(co.best_multiexp(ds.cpu().numpy().view(np.uint64).reshape(n, 4), db.cpu().numpy().view(np.uint64).reshape(n, 8)))
        print("   parity:", ref == exp, flush=True)
    del ds, db
