"""One 2^k commit through the host-pointer call (b200zk_msm_g1_registered, page-locked scalars) against the
device-resident call, for the upload-pipeline knobs in the environment (B200ZK_MSM_PIPE_PARTS / _GROWTH).
Usage: python scratch/r2_e2e_commit.py [k] [reps]"""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "anon-aadhaar-halo2_b200"))
import b200zk  # noqa: E402
from b200zk.api import _ptr  # noqa: E402

k = int(sys.argv[1]) if len(sys.argv) > 1 else 24
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
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
h_scal = torch.empty(n * 4, dtype=torch.int64).pin_memory()
h_scal.copy_(d_scal)
out_h = np.zeros(12, dtype=np.uint64)
out_d = torch.zeros(12, dtype=torch.int64, device=dev)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)


def host_call():
    b200zk.check(lib.b200zk_msm_g1_registered(handle.value, C.c_void_p(h_scal.data_ptr()), n, _ptr(out_h)))


def dev_call():
    b200zk.check(lib.b200zk_msm_g1_registered_dev(handle.value, vp(d_scal), n, 1, n, vp(out_d), st))
    torch.cuda.synchronize()


def wall(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        t.append((time.perf_counter() - t0) * 1e3)
    return min(t), sorted(t)[len(t) // 2]


hb, hm = wall(host_call)
db, dm = wall(dev_call)
same = bool((b200zk.g1_to_bytes(out_d.cpu().numpy().view(np.uint64).reshape(1, 12)) == b200zk.g1_to_bytes(out_h.reshape(1, 12))).all())
print(f"k={k} parts={os.environ.get('B200ZK_MSM_PIPE_PARTS', '4')} growth={os.environ.get('B200ZK_MSM_PIPE_GROWTH', '1')} "
      f"chunkdiv={os.environ.get('B200ZK_MSM_CHUNK_DIV', '2048')}: host-pointer {hb:.2f} / {hm:.2f} ms (min / median), "
      f"device-resident {db:.2f} / {dm:.2f} ms, same point (compressed encodings): {same}", flush=True)
