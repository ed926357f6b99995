// Probe: Karatsuba + separated Montgomery reduction (csrc/fieldmul_wide.cuh) against the CIOS
// product of field.cuh: bit-exact agreement on edge/random operands, and throughput.
#include <cstdio>
#include <cstdint>
#include <vector>
#include "field.cuh"
#include "fieldmul_wide.cuh"  // scratch/ (probe only)
using namespace zk;

template <class F> __device__ __forceinline__ F mulK(const F& a, const F& b) {
    uint32_t T[16]; F r;
    wide::mul_wide(T, a.l, b.l);
    wide::mont_reduce<typename F::Params>(r.l, T);
    return r;
}
template <class F> __device__ __forceinline__ F sqrK(const F& a) {
    uint32_t T[16]; F r;
    wide::sqr_wide(T, a.l);
    wide::mont_reduce<typename F::Params>(r.l, T);
    return r;
}
template <class F> __global__ void check_kernel(const F* a, const F* b, int n, int* bad) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int j = 0; j < n; ++j) {
        F x = a[i], y = b[j];
        F ref = x * y, k = mulK(x, y);
        bool ok = true;
        for (int t = 0; t < 8; ++t) ok &= (ref.l[t] == k.l[t]);      // same representative, not just same class
        F rs = x * x, ks = sqrK(x);
        for (int t = 0; t < 8; ++t) ok &= (rs.l[t] == ks.l[t]);
        if (!ok) atomicAdd(bad, 1);
    }
}
template <int MODE> __global__ void __launch_bounds__(256) tput(Fq* out, uint32_t iters) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    Fq a = Fq::one(), b = Fq::r2(), c = Fq::one().dbl(), d = Fq::r2().dbl();
    a.l[0] ^= tid & 0xffu; b.l[1] ^= tid & 0x3fu; c.l[2] ^= tid & 0x1fu; d.l[3] ^= tid & 0x7u;
    Fq m = Fq::r2(); m.l[1] ^= tid & 0xfu; m.reduce_once(); a.reduce_once();
    for (uint32_t i = 0; i < iters; ++i) {
        if (MODE == 0) { a = a * m; b = b * m; c = c * m; d = d * m; }
        if (MODE == 1) { a = mulK(a, m); b = mulK(b, m); c = mulK(c, m); d = mulK(d, m); }
        if (MODE == 2) { a = sqrK(a); b = sqrK(b); c = sqrK(c); d = sqrK(d); }
    }
    Fq s = (a + b) + (c + d);
    if (s.l[0] == 0xdeadbeefu && s.l[7] == 0x12345678u) out[tid] = s;
}
template <class F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
template <class F> int run_check(const char* name) {
    const int n = 192;
    std::vector<F> ha(n), hb(n);
    uint64_t s = 88172645463325252ull;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint32_t)(s >> 16); };
    uint32_t p[8], p2[8]; F::modulus(p); F::modulus2(p2);
    for (int i = 0; i < n; ++i) for (int t = 0; t < 8; ++t) { ha[i].l[t] = rnd(); hb[i].l[t] = rnd(); }
    for (int i = 0; i < n; ++i) { ha[i].l[7] &= 0x3fffffffu; hb[i].l[7] &= 0x3fffffffu; }   // < 2^254 < 2p
    auto setv = [&](F& f, const uint32_t* v, int delta) { for (int t = 0; t < 8; ++t) f.l[t] = v[t]; f.l[0] += delta; };
    uint32_t zero[8] = {0}, one[8] = {1}, ff[8];
    for (int t = 0; t < 8; ++t) ff[t] = 0xffffffffu; ff[7] = 0x3fffffffu;
    // edge operands: 0, 1, p-1, p, p+1, 2p-1, 2^254-1, limbs of all ones in each half
    const uint32_t* ev[] = {zero, one, p, p, p, p2, ff};
    const int ed[] = {0, 0, -1, 0, 1, -1, 0};
    for (int e = 0; e < 7; ++e) { setv(ha[e], ev[e], ed[e]); setv(hb[e], ev[e], ed[e]); }
    for (int t = 0; t < 8; ++t) { ha[7].l[t] = t < 4 ? 0xffffffffu : 0; hb[7].l[t] = t < 4 ? 0 : (t == 7 ? 0x3fffffffu : 0xffffffffu); }
    for (int t = 0; t < 8; ++t) { ha[8].l[t] = t < 4 ? 0xffffffffu : (t == 7 ? 0x3fffffffu : 0xffffffffu); hb[8] = ha[8]; }
    F *da, *db; int* dbad; int bad = 0;
    cudaMalloc(&da, n * sizeof(F)); cudaMalloc(&db, n * sizeof(F)); cudaMalloc(&dbad, 4);
    cudaMemcpy(da, ha.data(), n * sizeof(F), cudaMemcpyHostToDevice);
    cudaMemcpy(db, hb.data(), n * sizeof(F), cudaMemcpyHostToDevice);
    cudaMemset(dbad, 0, 4);
    check_kernel<F><<<(n + 63) / 64, 64>>>(da, db, n, dbad);
    cudaMemcpy(&bad, dbad, 4, cudaMemcpyDeviceToHost);
    printf("%s: %d x %d products + squares compared, mismatches %d (%s)\n", name, n, n, bad, cudaGetErrorString(cudaGetLastError()));
    return bad;
}
int main() {
    int bad = run_check<Fq>("Fq") + run_check<Fr>("Fr");
    const int blocks = 148 * 8, threads = 256; const uint32_t iters = 2048;
    void* sink; cudaMalloc(&sink, (size_t)blocks * threads * 64);
    const double n = 4.0 * blocks * threads * iters;
    float t0 = timeit([&] { tput<0><<<blocks, threads>>>((Fq*)sink, iters); });
    float t1 = timeit([&] { tput<1><<<blocks, threads>>>((Fq*)sink, iters); });
    float t2 = timeit([&] { tput<2><<<blocks, threads>>>((Fq*)sink, iters); });
    printf("CIOS mul       : %.3f ms  %.1f G modmul/s\n", t0, n / t0 / 1e6);
    printf("Karatsuba mul  : %.3f ms  %.1f G modmul/s\n", t1, n / t1 / 1e6);
    printf("Karatsuba sqr  : %.3f ms  %.1f G modsqr/s\n", t2, n / t2 / 1e6);
    printf("status %s\n", cudaGetErrorString(cudaGetLastError()));
    return bad != 0;
}
