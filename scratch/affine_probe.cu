// Probe: is a batched-affine bucket accumulation (6 products per addition + one Fermat inversion
// per thread per batch of B additions, operands streamed through HBM twice) faster than the XYZZ
// mixed addition (10 products) the MSM uses?  Round-0 access pattern: operands gathered at random
// from a large affine table through an index list.  Timing probe only (points are random field
// elements, not curve points; the arithmetic performed is the real affine addition law).
#include <cstdio>
#include <cstdint>
#include <vector>
#include "field.cuh"
#include "ec.cuh"
using namespace zk;

__device__ __forceinline__ Fq ldq(const Fq* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a, b;
    asm volatile("ld.global.nc.L2::64B.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "l"(q));
    asm volatile("ld.global.nc.L2::64B.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(q + 1));
    Fq r; r.l[0]=a.x; r.l[1]=a.y; r.l[2]=a.z; r.l[3]=a.w; r.l[4]=b.x; r.l[5]=b.y; r.l[6]=b.z; r.l[7]=b.w; return r;
}
__device__ __forceinline__ Fq ldp(const Fq* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    Fq r; r.l[0]=a.x; r.l[1]=a.y; r.l[2]=a.z; r.l[3]=a.w; r.l[4]=b.x; r.l[5]=b.y; r.l[6]=b.z; r.l[7]=b.w; return r;
}
__device__ __forceinline__ void stq(Fq* p, const Fq& v) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]); q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}

__global__ void fill_kernel(uint32_t* p, size_t nwords, uint32_t seed) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < nwords; i += stride) {
        uint64_t x = (i + 1) * 0x9E3779B97F4A7C15ull + seed; x ^= x >> 31; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 29;
        uint32_t v = (uint32_t)x;
        if ((i & 7) == 7) v &= 0x1fffffffu;      // keep every element below p
        p[i] = v;
    }
}
__global__ void idx_kernel(uint32_t* idx, size_t n, uint32_t mask) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t x = (i + 7) * 0xD6E8FEB86659FD93ull; x ^= x >> 32; x *= 0xD6E8FEB86659FD93ull; x ^= x >> 32;
    idx[i] = (uint32_t)x & mask;
}

// thread t owns additions t*B .. t*B+B-1; prefix products are stored transposed (step-major) so that
// a warp's stores / loads are contiguous
template <int B>
__global__ void __launch_bounds__(128) affine_round_kernel(const G1Affine* __restrict__ table, const uint32_t* __restrict__ idx,
                                                           Fq* __restrict__ prefix, G1Affine* __restrict__ out, uint32_t nthreads) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nthreads) return;
    const uint32_t* my = idx + (size_t)t * B * 2;
    Fq run = Fq::one();
#pragma unroll 1
    for (int j = 0; j < B; ++j) {
        const uint32_t a = my[2 * j], b = my[2 * j + 1];
        Fq d = ldq(&table[b].x) - ldq(&table[a].x);
        if (d.is_zero()) d = Fq::one();
        run = run * d;
        stq(prefix + (size_t)j * nthreads + t, run);
    }
    Fq inv = run.inverse();
#pragma unroll 1
    for (int j = B - 1; j >= 0; --j) {
        const uint32_t a = my[2 * j], b = my[2 * j + 1];
        const Fq x1 = ldq(&table[a].x), y1 = ldq(&table[a].y), x2 = ldq(&table[b].x), y2 = ldq(&table[b].y);
        Fq d = x2 - x1;
        if (d.is_zero()) d = Fq::one();
        const Fq pre = j ? ldp(prefix + (size_t)(j - 1) * nthreads + t) : Fq::one();
        const Fq dinv = inv * pre;
        inv = inv * d;
        const Fq lam = (y2 - y1) * dinv;
        const Fq x3 = lam.sqr() - x1 - x2;
        const Fq y3 = lam * (x1 - x3) - y1;
        G1Affine* o = out + (size_t)t * B + j;
        stq(&o->x, x3); stq(&o->y, y3);
    }
}

// the kernel it would replace, same access pattern: sequential XYZZ mixed additions
template <int B>
__global__ void __launch_bounds__(128) xyzz_kernel(const G1Affine* __restrict__ table, const uint32_t* __restrict__ idx,
                                                   G1Xyzz* __restrict__ out, uint32_t nthreads) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nthreads) return;
    const uint32_t* my = idx + (size_t)t * B * 2;
    G1Xyzz acc = G1Xyzz::identity();
#pragma unroll 1
    for (int j = 0; j < 2 * B; ++j) {
        G1Affine p; p.x = ldq(&table[my[j]].x); p.y = ldq(&table[my[j]].y);
        acc.add_affine(p);
    }
    out[t] = acc;
}

template <class F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize(); cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
template <int B> void run(const G1Affine* table, const uint32_t* idx, Fq* prefix, G1Affine* out, size_t nadds) {
    const uint32_t nthreads = (uint32_t)(nadds / B);
    float ms = timeit([&] { affine_round_kernel<B><<<(nthreads + 127) / 128, 128>>>(table, idx, prefix, out, nthreads); });
    printf("batched affine B=%-4d: %8.3f ms  %6.2f G additions/s  (%s)\n", B, ms, nadds / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    const size_t tab_n = (size_t)1 << 27;          // 8 GiB of affine points
    const size_t nadds = (size_t)100 << 20;        // round 0 of a 2^24-point MSM: ~100 M additions
    G1Affine* table; uint32_t* idx; Fq* prefix; G1Affine* out;
    cudaMalloc(&table, tab_n * sizeof(G1Affine)); cudaMalloc(&idx, nadds * 2 * 4);
    cudaMalloc(&prefix, nadds * sizeof(Fq)); cudaMalloc(&out, nadds * sizeof(G1Affine));
    fill_kernel<<<148 * 16, 256>>>((uint32_t*)table, tab_n * 16, 12345u);
    idx_kernel<<<(unsigned)((nadds * 2 + 255) / 256), 256>>>(idx, nadds * 2, (uint32_t)(tab_n - 1));
    cudaDeviceSynchronize();
    printf("setup: %s\n", cudaGetErrorString(cudaGetLastError()));
    {
        const uint32_t nthreads = (uint32_t)(nadds / 64);
        float ms = timeit([&] { xyzz_kernel<64><<<(nthreads + 127) / 128, 128>>>(table, idx, (G1Xyzz*)out, nthreads); });
        printf("XYZZ mixed additions : %8.3f ms  %6.2f G additions/s  (%s)\n", ms, 2.0 * nadds / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
    run<64>(table, idx, prefix, out, nadds);
    run<128>(table, idx, prefix, out, nadds);
    run<256>(table, idx, prefix, out, nadds);
    run<512>(table, idx, prefix, out, nadds);
    return 0;
}
