"""Cost per column of a batch of `count` commits of 2^15 device-resident columns (the prover's shape), with stage
times: where the small batches of the sharded proof (8-17 columns per coset owner) lose against the big ones."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "anon-aadhaar-halo2_b200"))
import b200zk  # noqa: E402
from b200zk.api import _ptr  # noqa: E402

k = 15
n = 1 << k
b200zk.init(0)
lib = b200zk.load()
dev = torch.device("cuda", 0)
vp = lambda t: C.c_void_p(t.data_ptr())
maxc = 128
d_scal = torch.empty(maxc * n * 4, dtype=torch.int64, device=dev)
d_base = torch.empty(n * 8, dtype=torch.int64, device=dev)
b200zk.check(lib.b200zk_gen_scalars_dev(vp(d_scal), maxc * n, 1000 + k, 0))
b200zk.check(lib.b200zk_gen_points_dev(vp(d_base), n, 2000 + k, 0))
h_bases = d_base.cpu().numpy().view(np.uint64).reshape(n, 8)
handle = C.c_uint64(0)
b200zk.check(lib.b200zk_bases_register(_ptr(h_bases), n, C.byref(handle)))
out_d = torch.zeros(maxc * 12, dtype=torch.int64, device=dev)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
names = ["hist", "scan", "scatter", "sync", "accumulate", "combine", "reduce", "reduce_combine", "fold"]
for count in (1, 2, 4, 8, 12, 16, 24, 32, 64, 128):
    # wall time without the stage timers
    b200zk.check(lib.b200zk_msm_profile(0))
    for _ in range(3):
        b200zk.check(lib.b200zk_msm_g1_registered_dev(handle.value, vp(d_scal), n, count, n, vp(out_d), st))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        b200zk.check(lib.b200zk_msm_g1_registered_dev(handle.value, vp(d_scal), n, count, n, vp(out_d), st))
    e1.record()
    torch.cuda.synchronize()
    wall = e0.elapsed_time(e1) / reps
    b200zk.check(lib.b200zk_msm_profile(1))
    acc = np.zeros(9)
    for r in range(4):
        b200zk.check(lib.b200zk_msm_g1_registered_dev(handle.value, vp(d_scal), n, count, n, vp(out_d), st))
        torch.cuda.synchronize()
        ms = (C.c_float * 9)()
        info = (C.c_uint64 * 5)()
        b200zk.check(lib.b200zk_msm_last_stages(ms, 9, info))
        if r >= 1:
            acc += np.array(list(ms))
    acc /= 3
    print(f"{count:4d} columns: {wall:7.3f} ms = {wall / count:6.3f} per column; c {int(info[1])} chunk {int(info[4])}  "
          + " ".join(f"{nm} {v:.3f}" for nm, v in zip(names, acc) if v > 0.004), flush=True)
