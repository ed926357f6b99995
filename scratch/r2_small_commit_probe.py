"""Round-2 probe: where a single 2^15 commit (one synchronous call, window table) spends its time,
for several window sizes.  Run on the GPU box: python scratch/r2_small_commit_probe.py"""
import ctypes as C, json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "anon-aadhaar-halo2_b200"))
import numpy as np, torch
import b200zk
from b200zk.api import _ptr

b200zk.init(0)
lib = b200zk.load()
names = ["hist", "scan", "scatter", "sync", "accumulate", "combine", "reduce", "reduce_combine", "fold"]
out = {}
for k in (15, 17):
    n = 1 << k
    ds = torch.empty(n * 4, dtype=torch.int64, device="cuda")
    db = torch.empty(n * 8, dtype=torch.int64, device="cuda")
    b200zk.check(lib.b200zk_gen_scalars_dev(C.c_void_p(ds.data_ptr()), n, 1, 0))
    b200zk.check(lib.b200zk_gen_points_dev(C.c_void_p(db.data_ptr()), n, 2, 0))
    hb = db.cpu().numpy().view(np.uint64).reshape(n, 8)
    hs = b200zk.host_alloc_fr(n)
    hs[:] = ds.cpu().numpy().view(np.uint64).reshape(n, 4)
    dpt = torch.zeros(12, dtype=torch.int64, device="cuda")
    for cbits in (1, 7, 8, 9, 10, 11, 12, 14, 16):
        h = C.c_uint64(0)
        b200zk.check(lib.b200zk_bases_register_ex(_ptr(hb), n, cbits, C.byref(h)))
        o = np.zeros(12, dtype=np.uint64)
        for _ in range(3):
            b200zk.check(lib.b200zk_msm_g1_registered(h.value, _ptr(hs), n, _ptr(o)))
        t0 = time.perf_counter()
        reps = 30
        for _ in range(reps):
            b200zk.check(lib.b200zk_msm_g1_registered(h.value, _ptr(hs), n, _ptr(o)))
        host_ms = 1e3 * (time.perf_counter() - t0) / reps
        b200zk.check(lib.b200zk_msm_profile(1))
        b200zk.check(lib.b200zk_msm_g1_registered_dev(h.value, C.c_void_p(ds.data_ptr()), n, 1, n, C.c_void_p(dpt.data_ptr()), None))
        ms = (C.c_float * 9)(); info = (C.c_uint64 * 5)()
        b200zk.check(lib.b200zk_msm_last_stages(ms, 9, info))
        b200zk.check(lib.b200zk_msm_profile(0))
        out[f"k{k}_c{cbits}"] = {"host_call_ms": host_ms, "c": int(info[1]), "windows": int(info[2]), "pairs": int(info[3]),
                                 "chunk": int(info[4]), "stages_ms": {nm: round(float(ms[i]), 4) for i, nm in enumerate(names)}}
        b200zk.check(lib.b200zk_bases_evict(h.value))
        print(k, cbits, json.dumps(out[f"k{k}_c{cbits}"]), flush=True)
    b200zk.host_free(hs)
Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "r2_small_commit_probe.json").write_text(json.dumps(out, indent=1))
