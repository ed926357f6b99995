"""Device-resident 2^k commit (window table) under b200zk_msm_tune(max_chunk, max_seglen, 0): total and stage times.
Usage: python scratch/r2_msm_tune_sweep.py [k]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "anon-aadhaar-halo2_b200"))
import b200zk  # noqa: E402
from b200zk.api import _ptr  # noqa: E402

k = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n = 1 << k
b200zk.init(0)
lib = b200zk.load()
dev = torch.device("cuda", 0)
vp = lambda t: C.c_void_p(t.data_ptr())
d_scal = torch.empty(n * 4, dtype=torch.int64, device=dev)
d_base = torch.empty(n * 8, dtype=torch.int64, device=dev)
b200zk.check(lib.b200zk_gen_scalars_dev(vp(d_scal), n, 1000 + k, 0))
b200zk.check(lib.b200zk_gen_points_dev(vp(d_base), n, 2000 + k, 0))
h_bases = d_base.cpu().numpy().view(np.uint64).reshape(n, 8)
handle = C.c_uint64(0)
b200zk.check(lib.b200zk_bases_register(_ptr(h_bases), n, C.byref(handle)))
out_d = torch.zeros(12, dtype=torch.int64, device=dev)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
names = ["hist", "scan", "scatter", "sync", "accumulate", "combine", "reduce", "reduce_combine", "fold"]
ref = None
for chunk, seg in [(128, 64), (192, 64), (256, 64), (384, 64), (128, 32), (128, 48), (128, 96), (256, 32), (256, 48)]:
    b200zk.check(lib.b200zk_msm_tune(chunk, seg, 0))
    b200zk.check(lib.b200zk_msm_profile(1))
    acc = np.zeros(9)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 6
    tot = 0.0
    for r in range(reps + 2):
        e0.record()
        b200zk.check(lib.b200zk_msm_g1_registered_dev(handle.value, vp(d_scal), n, 1, n, vp(out_d), st))
        e1.record()
        torch.cuda.synchronize()
        ms = (C.c_float * 9)()
        info = (C.c_uint64 * 5)()
        b200zk.check(lib.b200zk_msm_last_stages(ms, 9, info))
        if r >= 2:
            acc += np.array(list(ms))
            tot += e0.elapsed_time(e1)
    pt = b200zk.g1_to_bytes(out_d.cpu().numpy().view(np.uint64).reshape(1, 12)).tobytes()
    ref = ref or pt
    acc /= reps
    print(f"chunk {chunk:4d} seglen {seg:3d}: {tot / reps:7.3f} ms  " + " ".join(f"{nm} {v:.3f}" for nm, v in zip(names, acc) if v > 0.005)
          + f"  same point: {pt == ref}", flush=True)
