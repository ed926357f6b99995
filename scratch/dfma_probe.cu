// Probe: 254-bit Montgomery multiplication on the FP64 pipe (5 x 52-bit limbs, R'' = 2^260).
// hi/lo halves of each 52x52 product come from two fma.rz with magic constants; raw bit
// patterns are accumulated in 64-bit integers, the exponent constants removed once per column.
#include <cstdio>
#include <cstdint>
#include <vector>
#include "field.cuh"
using namespace zk;
typedef unsigned long long u64;

__host__ __device__ constexpr u64 p52(int i) {   // Fq modulus in 52-bit limbs
    const u64 w[5] = {0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull, 0};
    const int bit = 52 * i, word = bit >> 6, sh = bit & 63;
    u64 v = w[word] >> sh;
    if (sh > 12 && word + 1 < 5) v |= w[word + 1] << (64 - sh);
    return v & ((1ull << 52) - 1);
}
// -p^-1 mod 2^52 : low 52 bits of 0x87d20782e4866389
constexpr u64 PINV52 = 0x87d20782e4866389ull & ((1ull << 52) - 1);
constexpr u64 HI_C = 0x4670000000000000ull, LO_C = 0x4330000000000000ull;  // bits of 2^104, 2^52
constexpr u64 M52 = (1ull << 52) - 1;

struct FqD { u64 l[5]; };   // limbs < 2^52, value < 2p

__device__ __forceinline__ double to_d(u64 x) { return __longlong_as_double((long long)(x | LO_C)) - 4503599627370496.0; }

__device__ __forceinline__ FqD mulD(const FqD& a, const FqD& b) {
    const double C1 = 20282409603651670423947251286016.0;            // 2^104
    const double C2 = 20282409603651670423947251286016.0 + 4503599627370496.0;  // 2^104 + 2^52
    double ad[5], bd[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) { ad[i] = to_d(a.l[i]); bd[i] = to_d(b.l[i]); }
    // column k receives lo terms with i+j = k and hi terms with i+j = k-1, from a*b and from q*p
    u64 col[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) {
        int nlo = 0, nhi = 0;
        for (int i = 0; i < 5; ++i) for (int j = 0; j < 5; ++j) { if (i + j == k) nlo++; if (i + j + 1 == k) nhi++; }
        col[k] = 0ull - 2ull * ((u64)nlo * LO_C + (u64)nhi * HI_C);   // a*b and q*p contribute the same pattern
    }
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            const double ph = __fma_rz(ad[i], bd[j], C1);
            const double pl = __fma_rz(ad[i], bd[j], C2 - ph);
            col[i + j] += (u64)__double_as_longlong(pl);
            col[i + j + 1] += (u64)__double_as_longlong(ph);
        }
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const u64 q = (col[i] * PINV52) & M52;
        const double qd = to_d(q);
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            const double pj = (double)p52(j);
            const double ph = __fma_rz(qd, pj, C1);
            const double pl = __fma_rz(qd, pj, C2 - ph);
            col[i + j] += (u64)__double_as_longlong(pl);
            col[i + j + 1] += (u64)__double_as_longlong(ph);
        }
        col[i + 1] += col[i] >> 52;
    }
    FqD r;
#pragma unroll
    for (int k = 0; k < 4; ++k) { r.l[k] = col[5 + k] & M52; col[6 + k] += col[5 + k] >> 52; }
    r.l[4] = col[9];
    return r;
}

__host__ __device__ inline FqD from_words(const uint32_t* w32) {
    u64 w[5] = {0, 0, 0, 0, 0};
    for (int i = 0; i < 4; ++i) w[i] = (u64)w32[2 * i] | ((u64)w32[2 * i + 1] << 32);
    FqD r;
    for (int i = 0; i < 5; ++i) {
        const int bit = 52 * i, word = bit >> 6, sh = bit & 63;
        u64 v = w[word] >> sh;
        if (sh > 12 && word + 1 < 5) v |= w[word + 1] << (64 - sh);
        r.l[i] = v & M52;
    }
    return r;
}

__global__ void check_kernel(const uint32_t* a, const uint32_t* b, u64* out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    FqD r = mulD(from_words(a + 8 * i), from_words(b + 8 * i));
    for (int k = 0; k < 5; ++k) out[5 * i + k] = r.l[k];
}

template <int CH> __global__ void __launch_bounds__(256) kD(u64* out, uint32_t iters) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    FqD m; for (int k = 0; k < 5; ++k) m.l[k] = (0x123456789abcdull * (k + 1) + tid) & M52; m.l[4] &= (1ull << 44) - 1;
    FqD v[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) { for (int k = 0; k < 5; ++k) v[c].l[k] = (0xfedcba987654ull * (k + c + 1) + tid) & M52; v[c].l[4] &= (1ull << 44) - 1; }
    for (uint32_t i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CH; ++c) v[c] = mulD(v[c], m);
    }
    u64 s = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) for (int k = 0; k < 5; ++k) s += v[c].l[k];
    if (s == 0x0eadbeef12345ull) out[tid] = s;
}
__global__ void __launch_bounds__(256) k32(Fq* out, uint32_t iters) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    Fq a = Fq::one(), b = Fq::r2(), c = Fq::one().dbl(), d = Fq::r2().dbl();
    a.l[0] ^= tid & 0xffu;
    Fq m = Fq::r2(); m.l[1] ^= tid & 0xfu; m.reduce_once(); a.reduce_once();
    for (uint32_t i = 0; i < iters; ++i) { a = a * m; b = b * m; c = c * m; d = d * m; }
    Fq s = (a + b) + (c + d);
    if (s.l[0] == 0xdeadbeefu && s.l[7] == 0x12345678u) out[tid] = s;
}
// mixed: even warps run the FP64 multiplier, odd warps the integer one
__global__ void __launch_bounds__(256) kMix(u64* out, uint32_t itersD, uint32_t iters32) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    if ((threadIdx.x >> 5) & 1) {
        Fq a = Fq::one(), b = Fq::r2(), c = Fq::one().dbl(), d = Fq::r2().dbl();
        a.l[0] ^= tid & 0xffu;
        Fq m = Fq::r2(); m.l[1] ^= tid & 0xfu; m.reduce_once(); a.reduce_once();
        for (uint32_t i = 0; i < iters32; ++i) { a = a * m; b = b * m; c = c * m; d = d * m; }
        Fq s = (a + b) + (c + d);
        if (s.l[0] == 0xdeadbeefu && s.l[7] == 0x12345678u) out[tid] = s.l[1];
    } else {
        FqD m; for (int k = 0; k < 5; ++k) m.l[k] = (0x123456789abcdull * (k + 1) + tid) & M52; m.l[4] &= (1ull << 44) - 1;
        FqD v[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) { for (int k = 0; k < 5; ++k) v[c].l[k] = (0xfedcba987654ull * (k + c + 1) + tid) & M52; v[c].l[4] &= (1ull << 44) - 1; }
        for (uint32_t i = 0; i < itersD; ++i) {
#pragma unroll
            for (int c = 0; c < 4; ++c) v[c] = mulD(v[c], m);
        }
        u64 s = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) for (int k = 0; k < 5; ++k) s += v[c].l[k];
        if (s == 0x0eadbeef12345ull) out[tid] = s;
    }
}
template <class F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    // ---- correctness vectors: written to gpurun_out/dfma_check.txt for a Python check
    const int n = 256;
    std::vector<uint32_t> ha(8 * n), hb(8 * n);
    uint64_t s = 88172645463325252ull;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
    for (int i = 0; i < 8 * n; ++i) { ha[i] = (uint32_t)rnd(); hb[i] = (uint32_t)rnd(); }
    for (int i = 0; i < n; ++i) { ha[8 * i + 7] &= 0x3fffffffu; hb[8 * i + 7] &= 0x3fffffffu; }   // < 2^254 < 2p
    for (int k = 0; k < 8; ++k) { ha[k] = 0; hb[8 + k] = 0; }   // zero cases
    ha[16] = 1; for (int k = 1; k < 8; ++k) ha[16 + k] = 0;
    uint32_t *da, *db; u64* dout;
    cudaMalloc(&da, 32 * n); cudaMalloc(&db, 32 * n); cudaMalloc(&dout, 40 * n);
    cudaMemcpy(da, ha.data(), 32 * n, cudaMemcpyHostToDevice); cudaMemcpy(db, hb.data(), 32 * n, cudaMemcpyHostToDevice);
    check_kernel<<<1, n>>>(da, db, dout, n);
    std::vector<u64> hout(5 * n);
    cudaMemcpy(hout.data(), dout, 40 * n, cudaMemcpyDeviceToHost);
    FILE* f = fopen("gpurun_out/dfma_check.txt", "w");
    for (int i = 0; i < n; ++i) {
        for (int k = 7; k >= 0; --k) fprintf(f, "%08x", ha[8 * i + k]);
        fprintf(f, " ");
        for (int k = 7; k >= 0; --k) fprintf(f, "%08x", hb[8 * i + k]);
        for (int k = 0; k < 5; ++k) fprintf(f, " %llx", hout[5 * i + k]);
        fprintf(f, "\n");
    }
    fclose(f);
    // ---- throughput
    const int blocks = 148 * 8, threads = 256; const uint32_t iters = 2048;
    void* sink; cudaMalloc(&sink, (size_t)blocks * threads * 64);
    double nthr = (double)blocks * threads;
    float t32 = timeit([&] { k32<<<blocks, threads>>>((Fq*)sink, iters); });
    float tD4 = timeit([&] { kD<4><<<blocks, threads>>>((u64*)sink, iters); });
    float tD2 = timeit([&] { kD<2><<<blocks, threads>>>((u64*)sink, iters); });
    printf("int32 carry-chain x4 : %.3f ms  %.1f G modmul/s\n", t32, 4 * nthr * iters / t32 / 1e6);
    printf("fp64  DFMA x4        : %.3f ms  %.1f G modmul/s\n", tD4, 4 * nthr * iters / tD4 / 1e6);
    printf("fp64  DFMA x2        : %.3f ms  %.1f G modmul/s\n", tD2, 2 * nthr * iters / tD2 / 1e6);
    // mixed: choose iteration counts so both halves take about equally long alone
    const double rD = 4 * nthr * iters / tD4, r32 = 4 * nthr * iters / t32;   // modmul per ms
    const uint32_t itD = iters, it32 = (uint32_t)(iters * (r32 / rD));
    float tm = timeit([&] { kMix<<<blocks, threads>>>((u64*)sink, itD, it32); });
    printf("mixed warps          : %.3f ms  %.1f G modmul/s (D iters %u, int iters %u)\n", tm,
           (4 * (nthr / 2) * itD + 4 * (nthr / 2) * it32) / tm / 1e6, itD, it32);
    printf("status %s\n", cudaGetErrorString(cudaGetLastError()));
}
