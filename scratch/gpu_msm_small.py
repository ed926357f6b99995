"""Latency of registered (window-table) commits at prover sizes, single and batched."""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "anon-aadhaar-halo2_b200")
import numpy as np, ctypes as C, torch
import b200zk
from b200zk.api import _ptr
from oracle import bn254 as bn, c_oracle as co
b200zk.init(0); lib = b200zk.load()
names = ["hist", "scan", "scatter", "sync", "accum", "combine", "reduce", "red_comb", "fold"]
stream = torch.cuda.Stream(); st = C.c_void_p(stream.cuda_stream)
b200zk.check(lib.b200zk_msm_profile(1))
for k in [15, 16, 17]:
    n = 1 << k
    db = torch.empty(n * 8, dtype=torch.int64, device="cuda")
    b200zk.check(lib.b200zk_gen_points_dev(C.c_void_p(db.data_ptr()), n, 0xBA5E0000 + k, 0))
    hb = db.cpu().numpy().view(np.uint64).reshape(n, 8)
    for pre in (0, 1):
        h = C.c_uint64(0)
        t0 = time.time(); b200zk.check(lib.b200zk_bases_register_ex(_ptr(hb), n, pre, C.byref(h))); treg = time.time() - t0
        for cnt in (1, 8, 32):
            ds = torch.empty(cnt * n * 4, dtype=torch.int64, device="cuda")
            b200zk.check(lib.b200zk_gen_scalars_dev(C.c_void_p(ds.data_ptr()), cnt * n, 0xA11CE000 + k, 0))
            dout = torch.zeros(cnt * 12, dtype=torch.int64, device="cuda")
            with torch.cuda.stream(stream):
                def run(): b200zk.check(lib.b200zk_msm_g1_registered_dev(h.value, C.c_void_p(ds.data_ptr()), n, cnt, n, C.c_void_p(dout.data_ptr()), st))
                run(); torch.cuda.synchronize()
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record(stream); [run() for _ in range(5)]; e1.record(stream); torch.cuda.synchronize()
            ms = (C.c_float * 9)(); info = (C.c_uint64 * 5)()
            b200zk.check(lib.b200zk_msm_last_stages(ms, 9, info))
            tot = e0.elapsed_time(e1) / 5
            print(f"k={k} pre={pre} cnt={cnt} reg={treg*1e3:.1f}ms total={tot:.3f}ms per-msm={tot/cnt*1e3:.1f}us c={info[1]} W={info[2]} L={info[4]} | " + " ".join(f"{nm}={v:.3f}" for nm, v in zip(names, ms)), flush=True)
            if cnt == 8 and k == 15:
                hs = ds.cpu().numpy().view(np.uint64).reshape(cnt, n, 4)
                got = dout.cpu().numpy().view(np.uint64).reshape(cnt, 12)
                ok = all(bn.g1_jacobian_limbs_to_affine(got[j]) == bn.g1_jacobian_limbs_to_affine(co.best_multiexp(hs[j], hb)) for j in range(cnt))
                print("   parity:", ok, flush=True)
        b200zk.check(lib.b200zk_bases_evict(h.value))
