"""MSM parity + timings; NTT timings on a real (non-default) stream."""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "anon-aadhaar-halo2_b200")
import numpy as np, ctypes as C
import b200zk
from b200zk.api import fr_limbs, _ptr
from oracle import bn254 as bn, c_oracle as co
import torch

b200zk.init(0)
lib = b200zk.load()

def aff(j): return bn.g1_jacobian_limbs_to_affine(j)

for n in [0, 1, 2, 3, 7, 33, 100, 257, 1000, 4096, 5000, 1 << 14, 1 << 16]:
    sc = co.gen_scalars(100 + n, n); pts = co.gen_points(200 + n, n)
    if n >= 7:
        scc = bn.fr_array_to_canonical(sc[:7]); 
        scc[0] = 0; scc[1] = 1; scc[2] = bn.R - 1; scc[5] = scc[6]
        sc[:7] = bn.fr_array_from_canonical(scc)
        ptl = bn.g1_affine_array_to_points(pts[:7]); ptl[4] = ptl[3]; ptl[6] = bn.g1_neg(ptl[5])
        pts[:7] = bn.g1_affine_array_from_points(ptl)
    t = time.time()
    got = aff(b200zk.best_multiexp(sc, pts)); t1 = time.time()
    exp = aff(co.best_multiexp(sc, pts)) if n else None
    print(f"msm n={n}: {'ok' if got == exp else 'MISMATCH'}  gpu(host-call) {1e3*(t1-t):.2f} ms", flush=True)

# adversarial
n = 3000
pts = co.gen_points(9, n)
for name, sc in [("all r-1", bn.fr_array_from_canonical([bn.R - 1] * n)),
                 ("all one", bn.fr_array_from_canonical([1] * n)),
                 ("all zero", bn.fr_array_from_canonical([0] * n)),
                 ("90% zero", None), ("small", None)]:
    if name == "90% zero":
        sc = co.gen_scalars(5, n); sc[np.arange(n) % 10 != 0] = 0
    if name == "small":
        sc = bn.fr_array_from_canonical([(i * 7) % 65536 for i in range(n)])
    got = aff(b200zk.best_multiexp(sc, pts)); exp = aff(co.best_multiexp(sc, pts))
    print(f"msm adversarial {name}: {'ok' if got == exp else 'MISMATCH'}", flush=True)
# same point everywhere, identity bases
p1 = np.repeat(co.gen_points(3, 1), 2000, axis=0); sc = co.gen_scalars(4, 2000)
print("msm same-point:", aff(b200zk.best_multiexp(sc, p1)) == aff(co.best_multiexp(sc, p1)), flush=True)
pz = co.gen_points(3, 500); pz[::3] = 0
sc = co.gen_scalars(4, 500)
print("msm identity-bases:", aff(b200zk.best_multiexp(sc, pz)) == aff(co.best_multiexp(sc, pz)), flush=True)
# registered bases
g = co.gen_points(21, 1 << 10); gl = co.gen_points(22, 1 << 10)
P = b200zk.ParamsKZG(g, gl); poly = co.gen_scalars(23, 1 << 10)
print("commit:", aff(P.commit(poly)) == aff(co.best_multiexp(poly, g)), "commit_lagrange:", aff(P.commit_lagrange(poly[:700])) == aff(co.best_multiexp(poly[:700], gl[:700])), flush=True)
P.close()

stream = torch.cuda.Stream()
st = C.c_void_p(stream.cuda_stream)
def omega_for(k):
    w = bn.FR_ROOT_OF_UNITY
    for _ in range(k, 28): w = w * w % bn.R
    return w
with torch.cuda.stream(stream):
    for k in [15, 16, 17, 18, 20, 22, 24, 26]:
        nn = 1 << k
        buf = torch.empty(nn * 4, dtype=torch.int64, device="cuda")
        b200zk.check(lib.b200zk_gen_scalars_dev(C.c_void_p(buf.data_ptr()), nn, 0xA11CE000 + k, 0))
        w = fr_limbs(omega_for(k))
        def fwd(): b200zk.check(lib.b200zk_ntt_dev(C.c_void_p(buf.data_ptr()), nn, 1, k, _ptr(w), None, st))
        fwd(); fwd(); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record(stream)
        for _ in range(reps): fwd()
        e1.record(stream); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        passes = -(-k // 9)
        print(f"ntt_dev k={k}: {ms:.3f} ms  alg {64*nn*max(1,-(-k//12))/ms/1e6:.1f} GB/s  actual {64*nn*passes/ms/1e6:.1f} GB/s  {nn*k/2/ms/1e6:.2f} Gbfly/s", flush=True)
        del buf
    # batched prover-shape NTTs: 64 columns of 2^15 -> coeff_to_extended 2^17
    for (k, cnt) in [(15, 64), (17, 64)]:
        nn = 1 << k
        buf = torch.empty(cnt * nn * 4, dtype=torch.int64, device="cuda")
        b200zk.check(lib.b200zk_gen_scalars_dev(C.c_void_p(buf.data_ptr()), cnt * nn, 5, 0))
        w = fr_limbs(omega_for(k))
        def fwdb(): b200zk.check(lib.b200zk_ntt_dev(C.c_void_p(buf.data_ptr()), nn, cnt, k, _ptr(w), None, st))
        fwdb(); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(5): fwdb()
        e1.record(stream); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"ntt_dev batch {cnt} x 2^{k}: {ms:.3f} ms  ({ms/cnt*1e3:.1f} us/col)  {cnt*nn*k/2/ms/1e6:.2f} Gbfly/s", flush=True)
    for k in [12, 15, 16, 18, 20, 22, 24]:
        nn = 1 << k
        ds = torch.empty(nn * 4, dtype=torch.int64, device="cuda")
        db = torch.empty(nn * 8, dtype=torch.int64, device="cuda")
        b200zk.check(lib.b200zk_gen_scalars_dev(C.c_void_p(ds.data_ptr()), nn, 0xA11CE000 + k, 0))
        b200zk.check(lib.b200zk_gen_points_dev(C.c_void_p(db.data_ptr()), nn, 0xBA5E0000 + k, 0))
        out = np.zeros(12, dtype=np.uint64)
        def run(): b200zk.check(lib.b200zk_msm_g1_dev(C.c_void_p(ds.data_ptr()), C.c_void_p(db.data_ptr()), nn, _ptr(out), st))
        run(); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        reps = 3
        e0.record(stream)
        for _ in range(reps): run()
        e1.record(stream); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"msm_dev k={k}: {ms:.3f} ms  {nn/ms/1e3:.2f} Mpts/s", flush=True)
        if k <= 18:
            exp = aff(co.best_multiexp(ds.cpu().numpy().view(np.uint64).reshape(nn, 4), db.cpu().numpy().view(np.uint64).reshape(nn, 8)))
            print("    parity vs C oracle:", aff(out) == exp, flush=True)
        del ds, db
print("launches", b200zk.kernel_launches())
