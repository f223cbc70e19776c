// kmc_kernels.cuh -- the fused stretch-move kernels (thread-per-walker family, FP64).
//
// One walker-step fuses everything inside the reference's threaded loop body,
// src/samplers.jl:249-272: partner draw, z draw, proposal, log-density, accept test,
// in-place update, accept counter, thinned chain store.
//
// emcee_run_kernel advances a RANGE of half-steps in ONE launch: the ensemble state
// (x, logp, accept counters) stays in L2/HBM, every thread owns the same walkers for the
// whole launch, and consecutive half-steps are separated by a grid-wide barrier (the
// reference's thread join at the end of each `Threads.@threads` sweep, :248/:273).  With a
// range of one half-step it degenerates to the "one launch per half-step" form.
#pragma once
#include <math_constants.h>

#include "kmc_device.cuh"

namespace kmc {

struct RunParams {
    double *x;          // [nw][D] row-major walker positions (theta0s, :198)
    double *lp;         // [nw] current log-density (p0s, :209)
    unsigned *nacc;     // [nw] accept counters (naccept, :242)
    double *chain_x;    // [ns][nw][D] sample-major thinned chain (coalesced stores)
    double *chain_lp;   // [ns][nw]
    const long long *rp_partner;  // replay draws, indexed ((t-rp_t0)*2+batch)*nhalf + i
    const double *rp_z;
    const double *rp_u;
    long long rp_t0;
    long long nw, nhalf;
    long long h0, h1;   // half-step range, h = 2*t + batch, t = 0-based outer iteration
    long long n0;       // the reference's loop variable n (:245) at t = h0/2
    long long phase0;   // n0 mod nthin (floored)
    long long sidx0;    // samples already stored before n0
    long long nthin, ns;
    double sia, span, nm1;  // sqrt(1/a), sqrt(a)-sqrt(1/a), (N-1)
    unsigned long long seed;
    long long id_base;      // Philox walker id = id_base + batch*id_half_stride + i
    long long id_half_stride;
    unsigned lemire_t;      // (2^32 - nhalf) mod nhalf
    unsigned long long *barrier;  // grid barrier arrival counter (monotonic)
    unsigned long long bar_base;  // its value when this launch starts
};

__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Grid-wide barrier on a monotonic arrival counter (cooperative launch guarantees that all
// CTAs are resident).  Release: every thread's stores -> bar.sync -> fence.gpu by thread 0 ->
// atomic arrive.  Acquire: thread 0 spins with ld.acquire.gpu -> bar.sync -> everyone reads
// partner rows with ld.global.cg.
__device__ __forceinline__ void grid_barrier(unsigned long long *ctr, unsigned long long target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(ctr, 1ULL);
        while (ld_acquire_gpu(ctr) < target) {
        }
    }
    __syncthreads();
}

template <template <int> class Dn, int D, bool REPLAY>
__device__ __forceinline__ void walker_step(const RunParams &p, const Dn<D> &dn, long long t,
                                            int batch, long long i, bool store, long long sidx) {
    // :247 batch 0: active = first half, passive = second half; batch 1: swapped
    const long long a0 = batch ? p.nhalf : 0;
    const long long p0 = batch ? 0 : p.nhalf;
    const long long k = a0 + i;
    long long j;
    double z, u;
    if constexpr (REPLAY) {
        const long long slot = ((t - p.rp_t0) * 2 + batch) * p.nhalf + i;
        j = __ldcs(p.rp_partner + slot);
        z = __ldcs(p.rp_z + slot);
        u = __ldcs(p.rp_u + slot);
    } else {
        unsigned pl;
        double uz;
        draw(p.seed, (unsigned long long)(p.id_base + batch * p.id_half_stride + i),
             (unsigned long long)t, (unsigned)batch, (unsigned)p.nhalf, p.lemire_t, pl, uz, u);
        j = p0 + pl;
        const double s = dadd(dmul(uz, p.span), p.sia);  // :227
        z = dmul(s, s);
    }
    double xk[D], xj[D], y[D];
    load_row_cg<D>(p.x + j * D, xj);
    load_row<D>(p.x + k * D, xk);
    const double lpk = p.lp[k];
#pragma unroll
    for (int c = 0; c < D; ++c) y[c] = dadd(xj[c], dmul(z, dsub(xk[c], xj[c])));  // :255
    const double p1 = dn.logpdf(y);                                                // :257
    double lhs;
    if constexpr (D == 1 && !REPLAY) {
        lhs = dsub(p1, lpk);  // (N-1)*log(z) == 0 exactly: z is finite and positive here
    } else {
        lhs = dsub(dadd(dmul(p.nm1, log(z)), p1), lpk);  // :260
    }
    const bool acc = lhs >= log(u);
    double lpn = lpk;
    if (acc) {  // :261-265
        store_row<D>(p.x + k * D, y);
        p.lp[k] = p1;
        p.nacc[k] += 1u;
        lpn = p1;
    }
    if (store) {  // :268-272 -- the walker's current state, right after its own update
        const long long o = sidx * p.nw + k;
        if (acc) store_row_cs<D>(p.chain_x + o * D, y);
        else store_row_cs<D>(p.chain_x + o * D, xk);
        __stcs(p.chain_lp + o, lpn);
    }
}

template <template <int> class Dn, int D, bool REPLAY>
__global__ void __launch_bounds__(256) emcee_run_kernel(const RunParams p, const Dn<D> dn) {
    const long long nthreads = (long long)gridDim.x * blockDim.x;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long n = p.n0, phase = p.phase0, sidx = p.sidx0;
    unsigned long long target = p.bar_base;
    for (long long h = p.h0; h < p.h1; ++h) {
        const long long t = h >> 1;
        const int batch = (int)(h & 1);
        const bool store = (n > 0) && (phase == 0);  // :268  n>0 && rem(n,nthin)==0
        for (long long i = gid; i < p.nhalf; i += nthreads)
            walker_step<Dn, D, REPLAY>(p, dn, t, batch, i, store, sidx);
        if (batch == 1) {
            if (n == 0) {  // :285-288 burn-in counters are discarded
                for (long long i = gid; i < p.nhalf; i += nthreads) {
                    p.nacc[i] = 0u;
                    p.nacc[p.nhalf + i] = 0u;
                }
            }
            if (store) ++sidx;
            ++n;
            if (++phase == p.nthin) phase = 0;
        }
        if (h + 1 < p.h1) {
            target += gridDim.x;
            if (gridDim.x > 1) grid_barrier(p.barrier, target);
            else __syncthreads();
        }
    }
}

// K4: batched log-density (initial p0s, src/samplers.jl:209; make_theta0s, :334-338).
template <template <int> class Dn, int D>
__global__ void __launch_bounds__(256) density_eval_kernel(const double *__restrict__ x, double *__restrict__ out,
                                                           long long nw, const Dn<D> dn) {
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= nw) return;
    double v[D];
    load_row<D>(x + w * D, v);
    out[w] = dn.logpdf(v);
}

// Chain store is sample-major on the device ([ns][nw][D]); the ABI wants walker-major
// ([nw][ns][D]).  Transposes walkers [w0, w0+wc) into a staging buffer.
__global__ void chain_transpose_kernel(const double *__restrict__ in, double *__restrict__ out, long long ns,
                                       long long nw, long long w0, long long wc, int d) {
    __shared__ double tile[32][33];
    // element (s, w) is a d-vector; handle one component per blockIdx.z
    const int c = blockIdx.z;
    const long long wb = (long long)blockIdx.x * 32, sb = (long long)blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const long long s = sb + r, w = wb + threadIdx.x;
        if (s < ns && w < wc) tile[r][threadIdx.x] = in[(s * nw + (w0 + w)) * d + c];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const long long w = wb + r, s = sb + threadIdx.x;
        if (s < ns && w < wc) out[(w * ns + s) * d + c] = tile[threadIdx.x][r];
    }
}

// K5: accept-counter statistics of the progress display (src/samplers.jl:276-278).
__global__ void nacc_sum_kernel(const unsigned *__restrict__ nacc, long long nw, unsigned long long *sum) {
    unsigned long long s = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nw; i += (long long)gridDim.x * blockDim.x)
        s += nacc[i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(sum, s);
}

__global__ void nacc_moment_kernel(const unsigned *__restrict__ nacc, long long nw, double mean, double thresh,
                                   double *ssq, unsigned long long *outl) {
    double s = 0.0;
    unsigned long long o = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nw; i += (long long)gridDim.x * blockDim.x) {
        const double dv = (double)nacc[i] - mean;
        s += dv * dv;
        o += (thresh >= 0.0 && fabs(dv) > thresh) ? 1 : 0;
    }
    for (int k = 16; k > 0; k >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, k);
        o += __shfl_xor_sync(0xffffffffu, o, k);
    }
    if ((threadIdx.x & 31) == 0) {
        if (ssq) atomicAdd(ssq, s);
        if (outl && o) atomicAdd(outl, o);
    }
}

}  // namespace kmc
