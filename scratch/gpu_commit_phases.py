"""Stage split of the batched commitments of the proof-shaped workload (k=15)."""
import sys, ctypes as C
sys.path.insert(0, "."); sys.path.insert(0, "anon-aadhaar-halo2_b200")
import numpy as np, torch
import b200zk
from b200zk.api import _ptr
b200zk.init(0); lib = b200zk.load()
names = ["hist", "scan", "scatter", "sync", "accum", "combine", "reduce", "red_comb", "fold"]
k = 15; n = 1 << k
db = torch.empty(n * 8, dtype=torch.int64, device="cuda")
b200zk.check(lib.b200zk_gen_points_dev(C.c_void_p(db.data_ptr()), n, 1, 0))
hb = db.cpu().numpy().view(np.uint64).reshape(n, 8)
h = C.c_uint64(0); b200zk.check(lib.b200zk_bases_register(_ptr(hb), n, C.byref(h)))
b200zk.check(lib.b200zk_msm_profile(1))
import os
if os.environ.get("SEGLEN"): b200zk.check(lib.b200zk_msm_tune(128, int(os.environ["SEGLEN"]), 0))
for cnt in (3, 112, 242):
    ds = torch.empty(cnt * n * 4, dtype=torch.int64, device="cuda")
    b200zk.check(lib.b200zk_gen_scalars_dev(C.c_void_p(ds.data_ptr()), cnt * n, 5, 0))
    dout = torch.zeros(cnt * 12, dtype=torch.int64, device="cuda")
    def run(): b200zk.check(lib.b200zk_msm_g1_registered_dev(h.value, C.c_void_p(ds.data_ptr()), n, cnt, n, C.c_void_p(dout.data_ptr()), None))
    run(); torch.cuda.synchronize()
    import time
    t0 = time.perf_counter(); [run() for _ in range(3)]; torch.cuda.synchronize(); tot = (time.perf_counter() - t0) / 3 * 1e3
    ms = (C.c_float * 9)(); info = (C.c_uint64 * 5)()
    b200zk.check(lib.b200zk_msm_last_stages(ms, 9, info))
    pairs = info[3]
    print(f"cnt={cnt} total={tot:.3f}ms per={tot/cnt*1e3:.1f}us c={info[1]} W={info[2]} L={info[4]} pairs={pairs} ideal_accum={pairs*10/65.8e9*1e3:.2f}ms | " + " ".join(f"{nm}={v:.3f}" for nm, v in zip(names, ms)), flush=True)
