// kmc_device.cuh -- device-side building blocks of the emcee stretch-move kernels (sm_100a).
//
// Arithmetic contract (shared with oracle/kmc_oracle.c): IEEE binary64, round-to-nearest and
// NO fused multiply-add on anything that decides accept/reject or produces chain state.  The
// reference is Julia, which never contracts a*b+c (src/samplers.jl:255,:260), so every such
// operation goes through the __d*_rn intrinsics, which nvcc never fuses.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace kmc {

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }

// ------------------------------------------------------------------ Philox4x32-10
struct Philox4 {
    uint32_t r0, r1, r2, r3;
};

__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return Philox4{c0, c1, c2, c3};
}

// One walker-step's draws, in the reference's order: partner (src/samplers.jl:250), the
// uniform behind z (:252 -> :230), the accept uniform (:260).  counter = (walker id,
// iteration lo, iteration hi, batch | attempt<<8), key = seed.  The partner uses Lemire's
// multiply-shift with rejection, so it is exactly uniform on [0, nhalf) like rand(range).
__device__ __forceinline__ void draw(uint64_t seed, uint64_t walker, uint64_t iter, uint32_t batch,
                                     uint32_t nhalf, uint32_t lemire_t, uint32_t &partner_local,
                                     double &uz, double &uacc) {
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    const uint32_t c0 = (uint32_t)walker, c1 = (uint32_t)iter, c2 = (uint32_t)(iter >> 32);
    const Philox4 r = philox4x32_10(c0, c1, c2, batch, k0, k1);
    uint64_t m = (uint64_t)r.r0 * nhalf;
    uint32_t attempt = 0;
    while ((uint32_t)m < lemire_t) {  // probability nhalf / 2^32 per draw
        ++attempt;
        const Philox4 rr = philox4x32_10(c0, c1, c2, batch | (attempt << 8), k0, k1);
        m = (uint64_t)rr.r0 * nhalf;
    }
    partner_local = (uint32_t)(m >> 32);
    const uint64_t bz = ((uint64_t)r.r1 << 16) | (r.r2 >> 16);
    const uint64_t ba = ((uint64_t)(r.r2 & 0xFFFFu) << 32) | r.r3;
    uz = (double)bz * 0x1p-48;    // exact: 48-bit integer times a power of two
    uacc = (double)ba * 0x1p-48;
}

// ------------------------------------------------------------------ log-density plugins
// Each plugin is a POD passed BY VALUE in the kernel arguments, so its parameters sit in the
// constant bank and feed the FP64 pipe directly (no loads).  logpdf() follows the operation
// order written in oracle/kmc_oracle.c::kmo_logpdf exactly.

constexpr int KIND_EXPONENTIAL = 0;
constexpr int KIND_ROSENBROCK = 1;
constexpr int KIND_GAUSSIAN = 2;
constexpr int KIND_LOGNORMAL = 3;
constexpr int KIND_LOGISTIC = 4;

// README.md:15  logpdf(x) = x<0 ? -Inf : -x   (d>1: independent Exp(1) components)
template <int D>
struct Exponential {
    static constexpr int kind = KIND_EXPONENTIAL;
    static constexpr int nparams = 0;
    double unused;
    __device__ __forceinline__ double logpdf(const double (&x)[D]) const {
        double s = x[0];
        bool neg = x[0] < 0.0;
#pragma unroll
        for (int c = 1; c < D; ++c) {
            neg = neg || (x[c] < 0.0);
            s = dadd(s, x[c]);
        }
        return neg ? -CUDART_INF : -s;
    }
};

// test/runtests.jl:68  -(100*(x2 - x1^2)^2 + (1 - x1)^2)/20, params [a, b, T]
template <int D>
struct Rosenbrock {
    static_assert(D == 2, "rosenbrock is 2-D");
    static constexpr int kind = KIND_ROSENBROCK;
    static constexpr int nparams = 3;
    double p[3];
    __device__ __forceinline__ double logpdf(const double (&x)[D]) const {
        const double t = dsub(x[1], dmul(x[0], x[0]));
        const double q = dmul(p[1], dmul(t, t));
        const double m = dsub(p[0], x[0]);
        const double r = dadd(q, dmul(m, m));
        return ddiv(-r, p[2]);
    }
};

// test/runtests.jl:53,61  params [mu(D), A(D*D row-major), lognorm]
template <int D>
struct Gaussian {
    static constexpr int kind = KIND_GAUSSIAN;
    static constexpr int nparams = D + D * D + 1;
    double p[D + D * D + 1];
    __device__ __forceinline__ double logpdf(const double (&x)[D]) const {
        double c[D];
#pragma unroll
        for (int j = 0; j < D; ++j) c[j] = dsub(x[j], p[j]);
        double ss = 0.0;
#pragma unroll
        for (int i = 0; i < D; ++i) {
            double y = 0.0;
#pragma unroll
            for (int j = 0; j < D; ++j) y = dadd(y, dmul(p[D + i * D + j], c[j]));
            ss = dadd(ss, dmul(y, y));
        }
        return dsub(p[D + D * D], dmul(0.5, ss));
    }
};

// test/runtests.jl:56  LogNormal(mu, sigma); params [mu, sigma, log(sigma)+0.5*log(2pi)]
template <int D>
struct LogNormal {
    static_assert(D == 1, "lognormal is 1-D");
    static constexpr int kind = KIND_LOGNORMAL;
    static constexpr int nparams = 3;
    double p[3];
    __device__ __forceinline__ double logpdf(const double (&x)[D]) const {
        if (!(x[0] > 0.0)) return (x[0] != x[0]) ? x[0] : -CUDART_INF;
        const double lx = log(x[0]);
        const double t = ddiv(dsub(lx, p[0]), p[1]);
        return dsub(dsub(-lx, dmul(0.5, dmul(t, t))), p[2]);
    }
};

// ------------------------------------------------------------------ row access
// Walker rows are contiguous ([nw][D] row-major): a partner gather touches ceil(8D/32)
// sectors instead of D sectors with a component-major layout, and for even D the active
// walker's own row is a coalesced 128-bit access.
template <int D>
__device__ __forceinline__ void load_row(const double *__restrict__ p, double (&v)[D]) {
    if constexpr (D % 2 == 0) {
#pragma unroll
        for (int c = 0; c < D; c += 2) {
            const double2 t = *reinterpret_cast<const double2 *>(p + c);
            v[c] = t.x;
            v[c + 1] = t.y;
        }
    } else {
#pragma unroll
        for (int c = 0; c < D; ++c) v[c] = p[c];
    }
}

// L2-coherent gather (ld.global.cg): partner rows are written by other SMs between
// half-steps of the persistent kernel, so they must not be served from a stale L1 line.
template <int D>
__device__ __forceinline__ void load_row_cg(const double *p, double (&v)[D]) {
    if constexpr (D % 2 == 0) {
#pragma unroll
        for (int c = 0; c < D; c += 2) {
            const double2 t = __ldcg(reinterpret_cast<const double2 *>(p + c));
            v[c] = t.x;
            v[c + 1] = t.y;
        }
    } else {
#pragma unroll
        for (int c = 0; c < D; ++c) v[c] = __ldcg(p + c);
    }
}

template <int D>
__device__ __forceinline__ void store_row(double *p, const double (&v)[D]) {
    if constexpr (D % 2 == 0) {
#pragma unroll
        for (int c = 0; c < D; c += 2) *reinterpret_cast<double2 *>(p + c) = make_double2(v[c], v[c + 1]);
    } else {
#pragma unroll
        for (int c = 0; c < D; ++c) p[c] = v[c];
    }
}

// Streaming store for the chain (written once, never re-read by the sampler).
template <int D>
__device__ __forceinline__ void store_row_cs(double *p, const double (&v)[D]) {
    if constexpr (D % 2 == 0) {
#pragma unroll
        for (int c = 0; c < D; c += 2) __stcs(reinterpret_cast<double2 *>(p + c), make_double2(v[c], v[c + 1]));
    } else {
#pragma unroll
        for (int c = 0; c < D; ++c) __stcs(p + c, v[c]);
    }
}

}  // namespace kmc
