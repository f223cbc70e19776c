// Row-gather microbenchmark: how fast can an SM pull uniformly random 80-byte rows into shared memory?
//   mode 0: one cp.async.bulk (TMA, UBLKCP) of 80 bytes per thread, mbarrier complete_tx
//   mode 1: five cp.async.cg 16-byte copies (LDGSTS) per thread, cp.async.mbarrier.arrive.noinc
//   mode 2: five LDG.128 into registers, then STS.128 (the synchronous baseline)
// Array size 64 MB (L2-resident) or 2 GB (HBM); 256 threads per CTA, 1..3 CTAs per SM, 200 rounds per CTA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o row_gather_bench row_gather_bench.cu && ./row_gather_bench
#include <cstdio>
#include <cuda_runtime.h>

constexpr int T = 256, D = 10, ROWB = D * 8, ROUNDS = 200;

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void __launch_bounds__(T, 3) gather_kernel(const double *__restrict__ x, unsigned nrows, double *sink) {
    extern __shared__ __align__(128) unsigned char sm[];
    double *rows = reinterpret_cast<double *>(sm);                      // [2][T][D]
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(rows + 2 * T * D);
    const unsigned tid = threadIdx.x;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(MODE == 1 ? T + 1 : 1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    unsigned state = (blockIdx.x * T + tid) * 2654435761u + 12345u, phase = 0;
    double acc = 0.0;
    for (int r = 0; r < ROUNDS; ++r) {
        state = state * 1664525u + 1013904223u;
        const unsigned row = (unsigned)(((unsigned long long)state * nrows) >> 32);
        const double *src = x + (size_t)row * D;
        double *dst = rows + ((size_t)(r & 1) * T + tid) * D;
        if (MODE == 0) {
            if (tid == 0)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(T * ROWB) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_u32(dst)), "l"(src), "r"(ROWB), "r"(smem_u32(bar)) : "memory");
        } else if (MODE == 1) {
#pragma unroll
            for (int c = 0; c < D; c += 2)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst + c)), "l"(src + c) : "memory");
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
            if (tid == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
        } else {
            double2 v[D / 2];
#pragma unroll
            for (int c = 0; c < D; c += 2) v[c / 2] = __ldcg(reinterpret_cast<const double2 *>(src + c));
#pragma unroll
            for (int c = 0; c < D; c += 2) *reinterpret_cast<double2 *>(dst + c) = v[c / 2];
        }
        if (MODE != 2) {
            unsigned done = 0;
            while (!done)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(smem_u32(bar)), "r"(phase) : "memory");
            phase ^= 1;
        } else {
            __syncthreads();
        }
        acc += dst[(tid * 7 + r) % D];  // consume
    }
    if (acc == 12345.678) sink[0] = acc;
}

template <int MODE>
void run(const char *name, const double *x, unsigned nrows, int ctas_per_sm, double *sink) {
    const size_t smem = 2 * T * ROWB + 16;
    cudaFuncSetAttribute(gather_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int grid = 148 * ctas_per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    gather_kernel<MODE><<<grid, T, smem>>>(x, nrows, sink);
    cudaEventRecord(e0);
    gather_kernel<MODE><<<grid, T, smem>>>(x, nrows, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double rows = (double)grid * T * ROUNDS;
    printf("%-28s rows of %u  ctas/SM %d: %8.3f ms  %7.2f Grows/s  %6.1f GB/s useful  %5.1f cycles/row/SM (1.9 GHz)  err=%s\n", name,
           nrows, ctas_per_sm, ms, rows / ms * 1e-6, rows * ROWB / ms * 1e-6, ms * 1e-3 * 1.9e9 / (rows / 148),
           cudaGetErrorString(cudaGetLastError()));
}

int main() {
    double *x, *sink;
    const size_t big = (size_t)1 << 31;
    cudaMalloc(&x, big);
    cudaMalloc(&sink, 8);
    cudaMemset(x, 0, big);
    for (unsigned bytes_log : {26u, 31u}) {
        const unsigned nrows = (unsigned)(((size_t)1 << bytes_log) / ROWB);
        for (int c : {1, 2, 3}) {
            run<0>("cp.async.bulk 80 B (TMA)", x, nrows, c, sink);
            run<1>("5 x cp.async.cg 16 B (LDGSTS)", x, nrows, c, sink);
            run<2>("5 x LDG.128 + STS.128", x, nrows, c, sink);
        }
    }
    return 0;
}
