"""Device-resident NTT timings for several k and batch shapes; env B200ZK_NTT_DIRECT_TW_MAX_LOG_N selects
the inter-pass twiddle table (one process per setting)."""
import os, sys
sys.path.insert(0, "."); sys.path.insert(0, "anon-aadhaar-halo2_b200")
import numpy as np, ctypes as C, torch
import b200zk
from b200zk.api import _ptr, fr_limbs, FR_ROOT_OF_UNITY, FR_MODULUS
b200zk.init(0); lib = b200zk.load()
stream = torch.cuda.Stream(); st = C.c_void_p(stream.cuda_stream)
out = []
for k, count in [(15, 242), (17, 64), (20, 8), (22, 1), (24, 1)]:
    n = 1 << k
    d = torch.empty(count * n * 4, dtype=torch.int64, device="cuda")
    b200zk.check(lib.b200zk_gen_scalars_dev(C.c_void_p(d.data_ptr()), count * n, 7 + k, 0))
    omega = fr_limbs(pow(FR_ROOT_OF_UNITY, 1 << (28 - k), FR_MODULUS))
    with torch.cuda.stream(stream):
        run = lambda: b200zk.check(lib.b200zk_ntt_dev(C.c_void_p(d.data_ptr()), n, count, k, _ptr(omega), None, st))
        run(); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(5): run()
        e1.record(stream); torch.cuda.synchronize()
    chk = int(d[:4].cpu().numpy().view(np.uint64)[0] % 1000003)
    out.append(f"k={k}x{count}: {e0.elapsed_time(e1)/5:.3f} ms (chk {chk})")
    del d
print(f"direct_max={os.environ.get('B200ZK_NTT_DIRECT_TW_MAX_LOG_N','0')} | " + " | ".join(out), flush=True)
