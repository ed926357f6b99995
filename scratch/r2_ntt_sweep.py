"""Round-2 probe: NTT times (device-resident) for the prover's and the benchmark's sizes, with a checksum of the
outputs so that builds can be compared.  B200ZK_LIB_PATH selects the build."""
import ctypes as C, hashlib, json, os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "anon-aadhaar-halo2_b200"))
import numpy as np, torch
import b200zk
from b200zk.api import _ptr, fr_limbs, FR_MODULUS, FR_ROOT_OF_UNITY, FR_S

def omega_for(k):
    w = FR_ROOT_OF_UNITY
    for _ in range(k, FR_S):
        w = w * w % FR_MODULUS
    return w

b200zk.init(0)
lib = b200zk.load()
res = {"lib": os.environ.get("B200ZK_LIB_PATH", "default")}
for k, count in ((15, 242), (17, 61), (20, 8), (22, 1), (24, 1), (26, 1)):
    n = 1 << k
    buf = torch.empty(count * n * 4, dtype=torch.int64, device="cuda")
    b200zk.check(lib.b200zk_gen_scalars_dev(C.c_void_p(buf.data_ptr()), count * n, 77 + k, 0))
    w = fr_limbs(omega_for(k))
    torch.cuda.synchronize()
    b200zk.check(lib.b200zk_ntt_dev(C.c_void_p(buf.data_ptr()), n, count, k, _ptr(w), None, None))
    torch.cuda.synchronize()
    digest = hashlib.sha256(buf[: min(buf.numel(), 1 << 22)].cpu().numpy().tobytes()).hexdigest()[:16]
    for _ in range(2):
        b200zk.check(lib.b200zk_ntt_dev(C.c_void_p(buf.data_ptr()), n, count, k, _ptr(w), None, None))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    # the library's stream is not torch's: time with host-side sync around a batch of launches
    import time
    t0 = time.perf_counter()
    for _ in range(reps):
        b200zk.check(lib.b200zk_ntt_dev(C.c_void_p(buf.data_ptr()), n, count, k, _ptr(w), None, None))
    torch.cuda.synchronize()
    ms = 1e3 * (time.perf_counter() - t0) / reps
    res[f"k{k}x{count}"] = {"ms": round(ms, 4), "digest": digest}
    del buf
print(json.dumps(res), flush=True)
