// Microbenchmark: issue / pipe throughput of FFMA2 (fma.rn.f32x2) against FFMA on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float *out, int iters) {
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const float m = 1.0000001f, c = 1e-7f;
    unsigned long long p0, p1, p2, p3, mm, cc;
    asm("mov.b64 %0, {%1,%2};" : "=l"(p0) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1,%2};" : "=l"(p1) : "f"(a2), "f"(a3));
    asm("mov.b64 %0, {%1,%2};" : "=l"(p2) : "f"(a4), "f"(a5));
    asm("mov.b64 %0, {%1,%2};" : "=l"(p3) : "f"(a6), "f"(a7));
    asm("mov.b64 %0, {%1,%1};" : "=l"(mm) : "f"(m));
    asm("mov.b64 %0, {%1,%1};" : "=l"(cc) : "f"(c));
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (MODE == 0) {  // 8 independent scalar FFMA chains
                a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
                a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
            } else {          // 4 independent FFMA2 chains (same 8 lanes of work)
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p0) : "l"(mm), "l"(cc));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p1) : "l"(mm), "l"(cc));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p2) : "l"(mm), "l"(cc));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p3) : "l"(mm), "l"(cc));
            }
        }
    }
    long long t1 = clock64();
    float r = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    float x, y;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(x), "=f"(y) : "l"(p0)); r += x + y;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(x), "=f"(y) : "l"(p1)); r += x + y;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(x), "=f"(y) : "l"(p2)); r += x + y;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(x), "=f"(y) : "l"(p3)); r += x + y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0)
        printf("mode %d: %d warps/SM-block: %.3f cycles per (8 fp32 FMA lanes-ops per thread) group\n", MODE, blockDim.x / 32,
               (double)(t1 - t0) / (iters * 8.0));
}
int main() {
    float *out; cudaMalloc(&out, 148 * 1024 * 4);
    for (int warps : {4, 8, 16, 32}) {
        k<0><<<148, warps * 32>>>(out, 4096); cudaDeviceSynchronize();
        k<1><<<148, warps * 32>>>(out, 4096); cudaDeviceSynchronize();
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
