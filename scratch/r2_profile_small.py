"""Round-2 profiling target: three single 2^15 commits (latency regime: direct quad finish, quad reduce) and one
batch of 242 commits + their inverse transforms (throughput regime), as create_proof issues them."""
import ctypes as C, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "anon-aadhaar-halo2_b200"))
import numpy as np, torch
import b200zk
from b200zk.api import _ptr, EvaluationDomain

b200zk.init(0)
lib = b200zk.load()
k, cols = 15, 242
n = 1 << k
ds = torch.empty(cols * n * 4, dtype=torch.int64, device="cuda")
db = torch.empty(n * 8, dtype=torch.int64, device="cuda")
b200zk.check(lib.b200zk_gen_scalars_dev(C.c_void_p(ds.data_ptr()), cols * n, 1, 0))
b200zk.check(lib.b200zk_gen_points_dev(C.c_void_p(db.data_ptr()), n, 2, 0))
hb = db.cpu().numpy().view(np.uint64).reshape(n, 8)
h = C.c_uint64(0)
b200zk.check(lib.b200zk_bases_register(_ptr(hb), n, C.byref(h)))
hs = b200zk.host_alloc_fr(n)
hs[:] = ds[: n * 4].cpu().numpy().view(np.uint64).reshape(n, 4)
out = np.zeros(12, dtype=np.uint64)
for _ in range(3):
    b200zk.check(lib.b200zk_msm_g1_registered(h.value, _ptr(hs), n, _ptr(out)))
dpts = torch.zeros(cols * 12, dtype=torch.int64, device="cuda")
d = EvaluationDomain(4, k)
for _ in range(2):
    b200zk.check(lib.b200zk_msm_g1_registered_dev(h.value, C.c_void_p(ds.data_ptr()), n, cols, n, C.c_void_p(dpts.data_ptr()), None))
    b200zk.check(lib.b200zk_ntt_dev(C.c_void_p(ds.data_ptr()), n, cols, k, _ptr(d.omega_inv), _ptr(d.ifft_divisor), None))
torch.cuda.synchronize()
print("ok", b200zk.kernel_launches())
