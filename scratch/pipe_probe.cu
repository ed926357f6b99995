// Probe: do IMAD.WIDE (fmaheavy), DFMA (fp64) and IADD3 (alu) pipes overlap on B200?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) probe(uint64_t* out, int iters) {
    uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t a0 = tid, a1 = tid + 1, a2 = tid + 2, a3 = tid + 3;
    uint32_t m0 = tid | 1, m1 = tid * 3 + 7;
    double d0 = tid, d1 = tid + 0.5, d2 = tid + 1.5, d3 = 1.25 * tid;
    double f = 1.0000001, g = 0.5;
    uint32_t i0 = tid, i1 = tid + 5, i2 = 77, i3 = 13;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (MODE & 1) {  // 4 IMAD.WIDE
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a0) : "r"(m0), "r"(m1));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a1) : "r"(m0), "r"(m1));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a2) : "r"(m0), "r"(m1));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a3) : "r"(m0), "r"(m1));
            }
            if (MODE & 2) {  // 4 DFMA
                asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(d0) : "d"(f), "d"(g));
                asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(d1) : "d"(f), "d"(g));
                asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(d2) : "d"(f), "d"(g));
                asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(d3) : "d"(f), "d"(g));
            }
            if (MODE & 4) {  // 4 IADD3-class
                asm volatile("add.u32 %0, %0, %1;" : "+r"(i0) : "r"(i2));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(i1) : "r"(i3));
                asm volatile("xor.b32 %0, %0, %1;" : "+r"(i2) : "r"(i0));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(i3) : "r"(i1));
            }
        }
    }
    out[tid] = a0 + a1 + a2 + a3 + (uint64_t)(d0 + d1 + d2 + d3) + i0 + i1 + i2 + i3;
}
template <int MODE> float run(uint64_t* out, int blocks, int iters) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<MODE><<<blocks, 256>>>(out, iters / 4);
    cudaEventRecord(e0); probe<MODE><<<blocks, 256>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    int sms = 148, blocks = sms * 8, iters = 4096;
    uint64_t* out; cudaMalloc(&out, (size_t)blocks * 256 * 8);
    double per = (double)blocks * 256 * iters * 8 * 4;  // ops of each kind
    float t1 = run<1>(out, blocks, iters), t2 = run<2>(out, blocks, iters), t4 = run<4>(out, blocks, iters);
    float t3 = run<3>(out, blocks, iters), t5 = run<5>(out, blocks, iters), t6 = run<6>(out, blocks, iters), t7 = run<7>(out, blocks, iters);
    printf("IMAD.WIDE only : %.3f ms  %.1f Gops/s\n", t1, per / t1 / 1e6);
    printf("DFMA only      : %.3f ms  %.1f Gops/s\n", t2, per / t2 / 1e6);
    printf("IADD only      : %.3f ms  %.1f Gops/s\n", t4, per / t4 / 1e6);
    printf("WIDE+DFMA      : %.3f ms  (sum %.3f, max %.3f)\n", t3, t1 + t2, t1 > t2 ? t1 : t2);
    printf("WIDE+IADD      : %.3f ms  (sum %.3f)\n", t5, t1 + t4);
    printf("DFMA+IADD      : %.3f ms  (sum %.3f)\n", t6, t2 + t4);
    printf("all three      : %.3f ms  (sum %.3f)\n", t7, t1 + t2 + t4);
    cudaError_t e = cudaGetLastError(); printf("status %s\n", cudaGetErrorString(e));
    return 0;
}
