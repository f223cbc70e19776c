// kmc_kernels.cuh -- the fused stretch-move kernels (thread-per-walker family, FP64).
//
// One walker-step fuses everything inside the reference's threaded loop body,
// src/samplers.jl:249-272: partner draw, z draw, proposal, log-density, accept test,
// in-place update, accept counter, thinned chain store.
//
// emcee_run_kernel advances a RANGE of half-steps in ONE launch.  Every CTA owns a fixed,
// contiguous slice of walker positions of BOTH halves for the whole launch, and consecutive
// half-steps are separated by a grid-wide barrier (the reference's thread join at the end of
// each `Threads.@threads` sweep, :248/:273).  Two residency modes for the owned state:
//   SMEM=true   x / logp / accept counters of the CTA's walkers live in shared memory for the
//               whole launch (component-major, conflict-free); global x is only written on
//               accept (so partners can gather it) and logp / counters are written back once
//               at the end.  Used when the ensemble fits the SMs' aggregate shared memory
//               (2^20 2-D walkers: 28 B x 7086 walkers = 194 KB per SM).
//   SMEM=false  state stays in L2/HBM; any ensemble size.
// With a range of one half-step the kernel degenerates to "one launch per half-step".
#pragma once
#include <math_constants.h>

#include "kmc_device.cuh"

namespace kmc {

// threads per CTA of emcee_run_kernel: small rows run 2 x 512 threads per SM in <= 64 registers;
// wider rows (more live FP64 values per walker) get up to 255 registers at 256 threads
template <int D>
__host__ __device__ constexpr int block_threads() { return D <= 4 ? 512 : 256; }

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
    long long nw;
    unsigned nhalf;
    unsigned per_cta;   // walker positions (of each half) owned by one CTA
    long long h0, h1;   // half-step range, h = 2*t + batch, t = 0-based outer iteration
    long long n0;       // the reference's loop variable n (:245) at t = h0/2
    long long phase0;   // n0 mod nthin (floored)
    long long sidx0;    // samples already stored before n0
    long long nthin, ns;
    double sia, span, nm1;  // sqrt(1/a), sqrt(a)-sqrt(1/a), (N-1)
    float nm1f, margin;     // fast accept filter: (N-1) and its rigorous error margin
    unsigned long long seed;
    long long id_base;      // Philox walker id = id_base + batch*id_half_stride + i
    long long id_half_stride;
    unsigned lemire_t;      // (2^32 - nhalf) mod nhalf
    unsigned long long *barrier;  // grid barrier arrival counter (monotonic)
    unsigned long long bar_base;  // its value when this launch starts
};

// ------------------------------------------------------------------ grid barrier
// Monotonic arrival counter; cooperative launch guarantees all CTAs are resident.
// arrive: all threads' stores -> bar.sync -> thread 0: red.release.gpu (+1)
// wait:   thread 0 polls (relaxed), then fence.acq_rel.gpu once -> bar.sync -> everyone
//         gathers partner rows with ld.global.cg (L2), never from L1.
__device__ __forceinline__ void barrier_arrive(unsigned long long *ctr) {
    asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(ctr) : "memory");
}
__device__ __forceinline__ void barrier_wait(const unsigned long long *ctr, unsigned long long target) {
    unsigned long long v;
    do {
        asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(ctr) : "memory");
    } while (v < target);
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
}

// ------------------------------------------------------------------ accept test
// Decides  ((N-1)*log(z) + p1) - p0 >= log(u)   (src/samplers.jl:260).
// Fast filter: t = (p1-p0) + ln2*((N-1)*lg2f(z) - lg2f(u)) with FP32 MUFU logs; its error is
// bounded by `margin` (set by the host: (N-1+64)*2e-6, >3x the worst case of two MUFU.LG2
// errors (2^-22.6 absolute on the mantissa part + one FP32 rounding of the result), two
// FP64->FP32 input roundings and one FP32 fma, for |lg2 z| <= 4 and |lg2 u| <= 100), so
// |t| > margin decides exactly like the FP64 expression.  Anything closer, non-finite or out of the safe
// range takes the exact FP64 path below -- the decisions are those of the exact expression.
template <bool SKIP_LOGZ>
__device__ __forceinline__ bool accept_exact(double nm1, double z, double p1, double p0, double u) {
    double lhs;
    if constexpr (SKIP_LOGZ) lhs = dsub(p1, p0);  // (N-1)*log(z) == 0 exactly: N == 1 and z finite positive
    else lhs = dsub(dadd(dmul(nm1, log(z)), p1), p0);
    return lhs >= log(u);
}

template <bool SKIP_LOGZ>
__device__ __forceinline__ bool accept_test(const RunParams &p, double z, double p1, double p0, double u) {
    const float zf = (float)z, uf = (float)u;
    if (zf > 0.0625f && zf < 16.0f && uf > 1e-30f && uf < 1e30f) {
        const float q = fmaf(p.nm1f, __log2f(zf), -__log2f(uf));
        const double t = (p1 - p0) + (double)q * 0.6931471805599453;
        if (t > (double)p.margin) return true;
        if (t < -(double)p.margin) return false;
    }
    return accept_exact<SKIP_LOGZ>(p.nm1, z, p1, p0, u);
}

// ------------------------------------------------------------------ draws for one walker-step
template <bool REPLAY>
__device__ __forceinline__ void step_draws(const RunParams &p, long long t, int batch, unsigned i, unsigned &j,
                                           double &z, double &u) {
    if constexpr (REPLAY) {
        const long long slot = ((t - p.rp_t0) * 2 + batch) * (long long)p.nhalf + i;
        j = (unsigned)__ldcs(p.rp_partner + slot);  // global 0-based index
        z = __ldcs(p.rp_z + slot);
        u = __ldcs(p.rp_u + slot);
    } else {
        unsigned pl;
        double uz;
        draw(p.seed, (unsigned long long)(p.id_base + batch * p.id_half_stride + i), (unsigned long long)t,
             (unsigned)batch, p.nhalf, p.lemire_t, pl, uz, u);
        j = (batch ? 0u : p.nhalf) + pl;               // :247 passive half
        const double s = dadd(dmul(uz, p.span), p.sia);  // :227
        z = dmul(s, s);
    }
}

// Owned-state accessors -------------------------------------------------------------------
// Shared-memory layout of one CTA (SMEM=true), L = 2*per_cta local slots (half 0 then half 1):
//   double xs[D][L]; double lps[L]; unsigned naccs[L];
template <int D>
struct SmemState {
    double *xs, *lps;
    unsigned *naccs;
    unsigned L;
    __device__ __forceinline__ SmemState(unsigned char *base, unsigned L_) : L(L_) {
        xs = reinterpret_cast<double *>(base);
        lps = xs + (size_t)D * L;
        naccs = reinterpret_cast<unsigned *>(lps + L);
    }
    static __host__ __device__ size_t bytes(unsigned per_cta) { return (size_t)2 * per_cta * (8 * D + 8 + 4); }
};

template <template <int> class Dn, int D, bool REPLAY, bool SMEM>
__global__ void __launch_bounds__(block_threads<D>(), (D <= 4 ? 2 : 1)) emcee_run_kernel(const RunParams p, const Dn<D> dn) {
    constexpr unsigned kBlock = block_threads<D>();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const unsigned tid = threadIdx.x;
    const unsigned base = blockIdx.x * p.per_cta;  // first owned position (in each half)
    const unsigned cnt = base >= p.nhalf ? 0u : min(p.per_cta, p.nhalf - base);
    SmemState<D> sm(smem_raw, 2 * p.per_cta);

    if constexpr (SMEM) {  // stage the CTA's walkers of both halves
        for (unsigned b = 0; b < 2; ++b)
            for (unsigned l = tid; l < cnt; l += kBlock) {
                const size_t k = (size_t)b * p.nhalf + base + l;
                double v[D];
                load_row<D>(p.x + k * D, v);
#pragma unroll
                for (int c = 0; c < D; ++c) sm.xs[c * sm.L + b * p.per_cta + l] = v[c];
                sm.lps[b * p.per_cta + l] = p.lp[k];
                sm.naccs[b * p.per_cta + l] = p.nacc[k];
            }
        // each thread only ever touches the slots it staged itself (same l stride): no sync needed
    }

    long long n = p.n0, phase = p.phase0, sidx = p.sidx0;
    unsigned long long target = p.bar_base;
    for (long long h = p.h0; h < p.h1; ++h) {
        const long long t = h >> 1;
        const int batch = (int)(h & 1);
        const bool store = (n > 0) && (phase == 0);  // :268  n>0 && rem(n,nthin)==0
        const unsigned a0 = batch ? p.nhalf : 0u;    // :247 active half
        const unsigned sl0 = batch ? p.per_cta : 0u;

        for (unsigned l = tid; l < cnt; l += kBlock) {
            const unsigned i = base + l;
            const size_t k = (size_t)a0 + i;
            unsigned j;
            double z, u;
            step_draws<REPLAY>(p, t, batch, i, j, z, u);  // :250, :252
            double xj[D], xk[D], y[D];
            load_row_cg<D>(p.x + (size_t)j * D, xj);
            double lpk;
            if constexpr (SMEM) {
#pragma unroll
                for (int c = 0; c < D; ++c) xk[c] = sm.xs[c * sm.L + sl0 + l];
                lpk = sm.lps[sl0 + l];
            } else {
                load_row<D>(p.x + k * D, xk);
                lpk = p.lp[k];
            }
#pragma unroll
            for (int c = 0; c < D; ++c) y[c] = dadd(xj[c], dmul(z, dsub(xk[c], xj[c])));  // :255
            const double p1 = dn.logpdf(y);                                                // :257
            const bool acc = accept_test<(D == 1 && !REPLAY)>(p, z, p1, lpk, u);           // :260
            if (acc) {  // :261-265
                store_row<D>(p.x + k * D, y);
                if constexpr (SMEM) {
#pragma unroll
                    for (int c = 0; c < D; ++c) sm.xs[c * sm.L + sl0 + l] = y[c];
                    sm.lps[sl0 + l] = p1;
                    sm.naccs[sl0 + l] += 1u;
                } else {
                    p.lp[k] = p1;
                    p.nacc[k] += 1u;
                }
            }
            if (store) {  // :268-272 -- the walker's current state, right after its own update
                const size_t o = (size_t)sidx * p.nw + k;
                if (acc) store_row_cs<D>(p.chain_x + o * D, y);
                else store_row_cs<D>(p.chain_x + o * D, xk);
                __stcs(p.chain_lp + o, acc ? p1 : lpk);
            }
        }
        if (batch == 1) {
            if (n == 0) {  // :285-288 burn-in counters are discarded
                for (unsigned l = tid; l < cnt; l += kBlock) {
                    if constexpr (SMEM) {
                        sm.naccs[l] = 0u;
                        sm.naccs[p.per_cta + l] = 0u;
                    } else {
                        p.nacc[base + l] = 0u;
                        p.nacc[(size_t)p.nhalf + base + l] = 0u;
                    }
                }
            }
            if (store) ++sidx;
            ++n;
            if (++phase == p.nthin) phase = 0;
        }
        if (h + 1 < p.h1) {
            target += gridDim.x;
            __syncthreads();
            if (gridDim.x > 1) {
                if (tid == 0) {
                    barrier_arrive(p.barrier);
                    barrier_wait(p.barrier, target);
                }
                __syncthreads();
            }
        }
    }

    if constexpr (SMEM) {  // write back what only lived in shared memory
        for (unsigned b = 0; b < 2; ++b)
            for (unsigned l = tid; l < cnt; l += kBlock) {
                const size_t k = (size_t)b * p.nhalf + base + l;
                p.lp[k] = sm.lps[b * p.per_cta + l];
                p.nacc[k] = sm.naccs[b * p.per_cta + l];
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
