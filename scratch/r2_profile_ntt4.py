"""ncu target for the sharded transform's kernels on ONE GPU: rank 0 of a world of 8 at k = 24 — the fused first pass
(ntt_pass_kernel<8, 2>: 256-point transforms on the rank's [2^8][2^13] slab, twiddle, scatter into the eight row
buffers, here all local) and the two row passes on the rank's 32 rows of 2^16.  Prints CUDA-event times."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "anon-aadhaar-halo2_b200"))
import b200zk  # noqa: E402
from b200zk import sharding  # noqa: E402
from b200zk.api import FR_MODULUS, FR_ROOT_OF_UNITY, _ptr, fr_limbs  # noqa: E402

k, world = 24, 8
b200zk.init(0)
lib = b200zk.load()
dev = torch.device("cuda", 0)
log_n1 = sharding.four_step_split(k, world)
n1, n2 = 1 << log_n1, 1 << (k - log_n1)
m, rows = n2 // world, n1 // world
omega = pow(FR_ROOT_OF_UNITY, 1 << (28 - k), FR_MODULUS)
w, w2 = fr_limbs(omega), fr_limbs(pow(omega, n1, FR_MODULUS))
slab = torch.empty(n1 * m * 4, dtype=torch.int64, device=dev)
b200zk.check(lib.b200zk_gen_scalars_dev(C.c_void_p(slab.data_ptr()), n1 * m, 77, 0))
bufs = [torch.empty(rows * n2 * 4, dtype=torch.int64, device=dev) for _ in range(world)]
bases = (C.c_void_p * world)(*[b.data_ptr() for b in bufs])
st = torch.cuda.Stream(device=dev)
sp = C.c_void_p(st.cuda_stream)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
with torch.cuda.stream(st):
    for rep in range(4):
        ev[0].record(st)
        b200zk.check(lib.b200zk_ntt4_first_pass_scatter_dev(C.c_void_p(slab.data_ptr()), k, log_n1, _ptr(w), world, 0, bases, n2, 0, sp))
        ev[1].record(st)
        b200zk.check(lib.b200zk_ntt_dev(C.c_void_p(bufs[0].data_ptr()), n2, rows, k - log_n1, _ptr(w2), None, sp))
        ev[2].record(st)
torch.cuda.synchronize()
print(f"k={k} world={world} split 2^{log_n1} x 2^{k - log_n1}: first pass + twiddle + scatter {ev[0].elapsed_time(ev[1]):.3f} ms, "
      f"row passes {ev[1].elapsed_time(ev[2]):.3f} ms ({n1 * m} elements per rank)")
