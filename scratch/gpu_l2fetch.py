"""DRAM traffic of msm_accum_kernel vs cudaLimitMaxL2FetchGranularity (run under ncu)."""
import sys, ctypes as C
sys.path.insert(0, "."); sys.path.insert(0, "anon-aadhaar-halo2_b200")
import torch, b200zk
gran = int(sys.argv[1]); k = 22; n = 1 << k
torch.cuda.init()
rt = torch.cuda.cudart()
if gran:
    print("set", rt.cudaDeviceSetLimit(5, gran) if hasattr(rt, "cudaDeviceSetLimit") else C.CDLL("libcudart.so.12").cudaDeviceSetLimit(5, C.c_size_t(gran)))
b200zk.init(0); lib = b200zk.load()
vp = lambda t: C.c_void_p(t.data_ptr())
s = torch.empty(n * 4, dtype=torch.int64, device="cuda"); b = torch.empty(n * 8, dtype=torch.int64, device="cuda"); o = torch.zeros(12, dtype=torch.int64, device="cuda")
b200zk.check(lib.b200zk_gen_scalars_dev(vp(s), n, 1, 0)); b200zk.check(lib.b200zk_gen_points_dev(vp(b), n, 2, 0))
for _ in range(2):
    b200zk.check(lib.b200zk_msm_g1_dev_async(vp(s), vp(b), n, vp(o), None))
torch.cuda.synchronize()
