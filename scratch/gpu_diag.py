"""First-contact GPU diagnostic: field peak, generators, NTT parity and timings."""
import sys, time, json
sys.path.insert(0, "."); sys.path.insert(0, "anon-aadhaar-halo2_b200")
import numpy as np, ctypes as C
import b200zk
from b200zk.api import fr_limbs, _ptr
from oracle import bn254 as bn, halo2_cpu as h
import torch

b200zk.init(0)
lib = b200zk.load()
print("modmul peak (G/s):", b200zk.modmul_peak(2048) / 1e9, flush=True)
print("modmul peak (G/s):", b200zk.modmul_peak(8192) / 1e9, flush=True)

# generators
n = 1000
d = torch.empty(n * 4, dtype=torch.int64, device="cuda")
b200zk.check(lib.b200zk_gen_scalars_dev(C.c_void_p(d.data_ptr()), n, 0xA11CE010, 0))
got = d.cpu().numpy().view(np.uint64).reshape(n, 4)
exp = bn.seeded_fr_mont_limbs(0xA11CE010, n)
print("gen_scalars ok:", np.array_equal(got, exp), flush=True)
npt = 64
dp = torch.empty(npt * 8, dtype=torch.int64, device="cuda")
b200zk.check(lib.b200zk_gen_points_dev(C.c_void_p(dp.data_ptr()), npt, 0xBA5E0010, 0))
gotp = bn.g1_affine_array_to_points(dp.cpu().numpy().view(np.uint64).reshape(npt, 8))
expp = bn.seeded_g1_points(0xBA5E0010, npt)
print("gen_points ok:", gotp == expp, flush=True)

def omega_for(k):
    w = bn.FR_ROOT_OF_UNITY
    for _ in range(k, 28): w = w * w % bn.R
    return w

for k in range(1, 15):
    nn = 1 << k
    a = bn.seeded_fr_mont_limbs(77 + k, nn)
    w = omega_for(k)
    exp = h.best_fft(bn.fr_array_to_canonical(a), w, k)
    got = a.copy()
    b200zk.best_fft(got, w, k)
    gotc = bn.fr_array_to_canonical(got)
    ok = gotc == exp
    print(f"ntt k={k}: {'ok' if ok else 'MISMATCH'}", flush=True)
    if not ok:
        bad = [i for i in range(nn) if gotc[i] != exp[i]]
        print("   first bad idx", bad[:16], "count", len(bad))

for (j, k) in [(3, 4), (4, 5), (5, 6), (4, 10), (3, 11)]:
    dom_o = h.EvaluationDomain(j, k); dom = b200zk.EvaluationDomain(j, k)
    a = bn.seeded_fr_mont_limbs(5 + k, 1 << k)
    ac = bn.fr_array_to_canonical(a)
    r1 = bn.fr_array_to_canonical(dom.lagrange_to_coeff(a)) == dom_o.lagrange_to_coeff(ac)
    ext = dom.coeff_to_extended(a)
    r2 = bn.fr_array_to_canonical(ext) == dom_o.coeff_to_extended(ac)
    e = bn.seeded_fr_mont_limbs(9 + k, 1 << dom.extended_k)
    ec = bn.fr_array_to_canonical(e)
    r3 = bn.fr_array_to_canonical(dom.extended_to_coeff(e)) == dom_o.extended_to_coeff(ec)
    r4 = bn.fr_array_to_canonical(dom.divide_by_vanishing_poly(e)) == dom_o.divide_by_vanishing_poly(ec)
    print(f"domain j={j} k={k} ext_k={dom.extended_k}: l2c={r1} c2e={r2} e2c={r3} div={r4}", flush=True)

# timings on device
st = torch.cuda.current_stream().cuda_stream
for k in [15, 16, 17, 18, 20, 22, 24, 26]:
    nn = 1 << k
    buf = torch.empty(nn * 4, dtype=torch.int64, device="cuda")
    b200zk.check(lib.b200zk_gen_scalars_dev(C.c_void_p(buf.data_ptr()), nn, 0xA11CE000 + k, 0))
    orig = buf.clone()
    w = fr_limbs(omega_for(k)); wi = fr_limbs(pow(omega_for(k), -1, bn.R)); ninv = fr_limbs(pow(nn, -1, bn.R))
    def fwd(): b200zk.check(lib.b200zk_ntt_dev(C.c_void_p(buf.data_ptr()), nn, 1, k, _ptr(w), None, C.c_void_p(st)))
    def inv(): b200zk.check(lib.b200zk_ntt_dev(C.c_void_p(buf.data_ptr()), nn, 1, k, _ptr(wi), _ptr(ninv), C.c_void_p(st)))
    fwd(); inv(); torch.cuda.synchronize()
    rt = torch.equal(buf, orig)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps): fwd()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    passes = -(-k // 9)
    print(f"ntt_dev k={k}: roundtrip={rt} {ms:.3f} ms  alg {64*nn*max(1,-(-k//12))/ms/1e6:.1f} GB/s  actual-traffic {64*nn*passes/ms/1e6:.1f} GB/s  {nn*k/2/ms/1e6:.1f} Gbutterfly/s", flush=True)
print("launches", b200zk.kernel_launches())
