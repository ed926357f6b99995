"""Registered-base MSM stage timings at k (default 24); env B200ZK_MSM_SORT_MODE / B200ZK_MSM_BIN_L2_MB select
the sort.  One process per setting (the knobs are read at load)."""
import os, sys
sys.path.insert(0, "."); sys.path.insert(0, "anon-aadhaar-halo2_b200")
import numpy as np, ctypes as C, torch
import b200zk
from oracle import bn254 as bn
b200zk.init(0); lib = b200zk.load()
names = ["hist", "scan", "scatter", "sync", "accum", "combine", "reduce", "red_comb", "fold"]
k = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n = 1 << k
stream = torch.cuda.Stream(); st = C.c_void_p(stream.cuda_stream)
ds = torch.empty(n * 4, dtype=torch.int64, device="cuda"); db = torch.empty(n * 8, dtype=torch.int64, device="cuda")
b200zk.check(lib.b200zk_gen_scalars_dev(C.c_void_p(ds.data_ptr()), n, 0xA11CE000 + k, 0))
b200zk.check(lib.b200zk_gen_points_dev(C.c_void_p(db.data_ptr()), n, 0xBA5E0000 + k, 0))
hb = db.cpu().numpy().view(np.uint64).reshape(n, 8)
h = C.c_uint64(0)
b200zk.check(lib.b200zk_bases_register(C.c_void_p(hb.ctypes.data), n, C.byref(h)))
d_out = torch.zeros(12, dtype=torch.int64, device="cuda")
if os.environ.get('SEGLEN'): b200zk.check(lib.b200zk_msm_tune(128, int(os.environ['SEGLEN']), 0))
b200zk.check(lib.b200zk_msm_profile(1))
with torch.cuda.stream(stream):
    run = lambda: b200zk.check(lib.b200zk_msm_g1_registered_dev(h.value, C.c_void_p(ds.data_ptr()), n, 1, n, C.c_void_p(d_out.data_ptr()), st))
    run(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream); run(); run(); run(); e1.record(stream); torch.cuda.synchronize()
ms = (C.c_float * 9)(); info = (C.c_uint64 * 5)()
b200zk.check(lib.b200zk_msm_last_stages(ms, 9, info))
aff = bn.g1_jacobian_limbs_to_affine(d_out.cpu().numpy().view(np.uint64))
print(f"seglen={os.environ.get('SEGLEN','-')} mode={os.environ.get('B200ZK_MSM_SORT_MODE','0')} l2mb={os.environ.get('B200ZK_MSM_BIN_L2_MB','-')} k={k} total={e0.elapsed_time(e1)/3:.3f}ms "
      f"x={aff[0] % 1000003} | " + " ".join(f"{nm}={v:.3f}" for nm, v in zip(names, ms)), flush=True)
