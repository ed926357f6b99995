"""Stage times of a device-resident commit of the first n of 2^24 registered points (one point range of the upload
pipeline, without its copy): where the ~0.8 ms a range costs beyond its share of the work goes."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "anon-aadhaar-halo2_b200"))
import b200zk  # noqa: E402
from b200zk.api import _ptr  # noqa: E402

k = 24
N = 1 << k
b200zk.init(0)
lib = b200zk.load()
dev = torch.device("cuda", 0)
vp = lambda t: C.c_void_p(t.data_ptr())
d_scal = torch.empty(N * 4, dtype=torch.int64, device=dev)
d_base = torch.empty(N * 8, dtype=torch.int64, device=dev)
b200zk.check(lib.b200zk_gen_scalars_dev(vp(d_scal), N, 1000 + k, 0))
b200zk.check(lib.b200zk_gen_points_dev(vp(d_base), N, 2000 + k, 0))
h_bases = d_base.cpu().numpy().view(np.uint64).reshape(N, 8)
handle = C.c_uint64(0)
b200zk.check(lib.b200zk_bases_register(_ptr(h_bases), N, C.byref(handle)))
out_d = torch.zeros(12, dtype=torch.int64, device=dev)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
names = ["hist", "scan", "scatter", "sync", "accumulate", "combine", "reduce", "reduce_combine", "fold"]
b200zk.check(lib.b200zk_msm_profile(1))
for frac in (1 / 21, 4 / 21, 16 / 21, 0.25, 0.5, 1.0):
    n = int(N * frac) // 256 * 256
    acc = np.zeros(9)
    reps = 5
    for r in range(reps + 2):
        b200zk.check(lib.b200zk_msm_g1_registered_dev(handle.value, vp(d_scal), n, 1, n, vp(out_d), st))
        torch.cuda.synchronize()
        ms = (C.c_float * 9)()
        info = (C.c_uint64 * 5)()
        b200zk.check(lib.b200zk_msm_last_stages(ms, 9, info))
        if r >= 2:
            acc += np.array(list(ms))
    acc /= reps
    tot = acc.sum()
    print(f"n = {n:9d} ({frac:.3f} of 2^24): {tot:7.3f} ms = {tot / frac:6.2f} ms per 2^24 points; chunk {int(info[4])}  "
          + " ".join(f"{nm} {v:.3f}" for nm, v in zip(names, acc) if v > 0.004), flush=True)
