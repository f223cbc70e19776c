/*
 * kmc_oracle.c -- CPU ORACLE for the emcee stretch-move hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the checker, never the product: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (libkissmcmc_cuda.so) never links, loads or calls anything in oracle/.
 *
 * What it restates (all file:line into the reference, /root/reference, KissMCMC.jl v0.2.2):
 *   g_pdf              src/samplers.jl:224
 *   cdf_g_inv          src/samplers.jl:227
 *   sample_g           src/samplers.jl:230
 *   _emcee             src/samplers.jl:232-293   (loop bounds :245, half split :247, draw order
 *                      partner -> z -> accept uniform :250,:252,:260, proposal :255,
 *                      accept test :260, update :261-265, thinned store :268-272,
 *                      burn-in counter reset :285-288, accept ratio :291)
 *   emcee (p0s)        src/samplers.jl:209-210   (initial log-density of every walker)
 *
 * PARITY STATUS: the reference cannot run here (no Julia in the image, no network) and its
 * tests hold no golden vectors and no seeds (test/emcee.jl, test/runtests.jl are purely
 * statistical).  This oracle is therefore pinned only against the reference's own
 * known-answer tests: g-distribution support / end points / moments (test/emcee.jl:2-14),
 * output shapes and sample counts (test/emcee.jl:29-41), accept_ratio > 0.1 (:43) and
 * posterior moments within tol*std (test/runtests.jl:36-43, 52-78; README.md:15).
 * Bitwise parity with Julia's RNG stream and Base.log is UNPINNED ("parity unpinned" for the
 * RNG stream; replay mode sidesteps the stream by taking partner/z/u as inputs).
 *
 * Arithmetic contract (what the CUDA kernels must reproduce bit-for-bit on the scalar paths):
 *   - IEEE-754 binary64, round-to-nearest, NO fused multiply-add (compile with
 *     -ffp-contract=off; Julia does not contract a*b+c either).
 *   - proposal (samplers.jl:255)   y[c] = xj[c] + z*(xk[c] - xj[c])      sub, mul, add
 *   - accept   (samplers.jl:260)   ((N-1)*log(z) + p1) - p0 >= log(u)    left to right
 *   - z        (samplers.jl:227)   s = u*(sqrt(a)-sqrt(1/a)) + sqrt(1/a);  z = s*s
 *   - densities: see kmo_logpdf below; each states its operation order.
 *
 * Counter-based draws (free-running mode): Philox4x32-10, key = 64-bit seed, counter =
 * (walker id, iteration lo, iteration hi, batch | attempt<<8).  One block yields the partner
 * (Lemire multiply-shift with rejection => exactly uniform), a 48-bit uniform for z and a
 * 48-bit uniform for the accept test, both in [0,1) like Julia's rand().
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define KMO_EXPONENTIAL 0
#define KMO_ROSENBROCK 1
#define KMO_GAUSSIAN 2
#define KMO_LOGNORMAL 3
#define KMO_LOGISTIC 4

typedef struct {
    int32_t kind;
    int32_t d;
    const double *params;
    int64_t nparams;
    const float *data; /* logistic: X[N][d] row-major then y[N], float32 */
    int64_t ndata;     /* logistic: N */
} kmo_density;

/* ---------------------------------------------------------------- g distribution */

/* src/samplers.jl:224 */
double kmo_g_pdf(double z, double a) {
    if (1.0 / a <= z && z <= a) return 1.0 / sqrt(z) * 1.0 / (2.0 * (sqrt(a) - sqrt(1.0 / a)));
    return 0.0;
}

/* src/samplers.jl:227 -- (u*(sqrt(a)-sqrt(1/a)) + sqrt(1/a))^2, x^2 lowers to x*x */
double kmo_cdf_g_inv(double u, double a) {
    double sa = sqrt(a), sia = sqrt(1.0 / a);
    double s = u * (sa - sia) + sia;
    return s * s;
}

/* ---------------------------------------------------------------- Philox4x32-10 */

void kmo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* One walker-step's draws.  partner_local in [0, nhalf) indexes the passive half. */
void kmo_draw(uint64_t seed, uint64_t walker, uint64_t iter, uint32_t batch, uint64_t nhalf,
              int64_t *partner_local, double *uz, double *uacc) {
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t ctr[4] = {(uint32_t)walker, (uint32_t)iter, (uint32_t)(iter >> 32), batch};
    uint32_t r[4];
    kmo_philox4x32_10(ctr, key, r);
    uint32_t n = (uint32_t)nhalf;
    uint64_t m = (uint64_t)r[0] * n;
    uint32_t lo = (uint32_t)m;
    if (lo < n) {
        uint32_t t = (uint32_t)(0u - n) % n;
        uint32_t attempt = 0;
        while (lo < t) {
            uint32_t rr[4];
            ++attempt;
            ctr[3] = batch | (attempt << 8);
            kmo_philox4x32_10(ctr, key, rr);
            m = (uint64_t)rr[0] * n;
            lo = (uint32_t)m;
        }
    }
    *partner_local = (int64_t)(m >> 32);
    uint64_t bz = ((uint64_t)r[1] << 16) | (r[2] >> 16);
    uint64_t ba = ((uint64_t)(r[2] & 0xFFFFu) << 32) | r[3];
    *uz = (double)bz * 0x1p-48;
    *uacc = (double)ba * 0x1p-48;
}

/* ---------------------------------------------------------------- densities */

static double softplus(double s) { return fmax(s, 0.0) + log1p(exp(-fabs(s))); }

/*
 * Log-density plugins.  These replace the user closure pdf(theta) of src/samplers.jl:257,:209.
 *  exponential  README.md:15            x<0 ? -Inf : -x   (d>1: product of Exp(1), sum left->right)
 *  rosenbrock   test/runtests.jl:68     -(b*(x2-x1^2)^2 + (a-x1)^2)/T, params [a,b,T]=[1,100,20]
 *  gaussian     test/runtests.jl:53,61  params [mu(d), A(d*d row-major), lognorm];
 *               y = A (x-mu) (row i: j ascending, mul then add), logp = lognorm - 0.5*sum(y_i^2)
 *  lognormal    test/runtests.jl:56     params [mu, sigma, c=log(sigma)+0.5*log(2pi)];
 *               x<=0 ? -Inf : ((-lx) - 0.5*(t*t)) - c, lx=log(x), t=(lx-mu)/sigma
 *  logistic     (not in the reference; BASELINE.json config 4) params [prior_sigma];
 *               sum_n (y_n*s_n - softplus(s_n)) - 0.5*|theta|^2/sigma^2, s_n = x_n . theta
 */
double kmo_logpdf(const kmo_density *dn, const double *x) {
    const int d = dn->d;
    const double *p = dn->params;
    switch (dn->kind) {
    case KMO_EXPONENTIAL: {
        double s = 0.0;
        for (int c = 0; c < d; ++c) {
            if (x[c] < 0.0) return -INFINITY;
            s = (c == 0) ? x[0] : s + x[c];
        }
        return -s;
    }
    case KMO_ROSENBROCK: {
        double t = x[1] - x[0] * x[0];
        double q = p[1] * (t * t);
        double m = p[0] - x[0];
        double r = q + m * m;
        return (-r) / p[2];
    }
    case KMO_GAUSSIAN: {
        const double *mu = p, *A = p + d;
        double lognorm = p[d + (int64_t)d * d];
        double ss = 0.0;
        for (int i = 0; i < d; ++i) {
            double y = 0.0;
            for (int j = 0; j < d; ++j) y = y + A[(int64_t)i * d + j] * (x[j] - mu[j]);
            ss = ss + y * y;
        }
        return lognorm - 0.5 * ss;
    }
    case KMO_LOGNORMAL: {
        if (!(x[0] > 0.0)) return (x[0] != x[0]) ? x[0] : -INFINITY;
        double lx = log(x[0]);
        double t = (lx - p[0]) / p[1];
        return ((-lx) - 0.5 * (t * t)) - p[2];
    }
    case KMO_LOGISTIC: {
        const int64_t N = dn->ndata;
        const float *X = dn->data, *y = dn->data + N * d;
        double acc = 0.0;
        for (int64_t n = 0; n < N; ++n) {
            double s = 0.0;
            for (int c = 0; c < d; ++c) s += (double)X[n * d + c] * x[c];
            acc += (double)y[n] * s - softplus(s);
        }
        double nn = 0.0;
        for (int c = 0; c < d; ++c) nn += x[c] * x[c];
        return acc - 0.5 * nn / (p[0] * p[0]);
    }
    }
    return NAN;
}

/* thetas row-major [nw][d] */
void kmo_density_eval(const kmo_density *dn, const double *thetas, int64_t nw, double *out,
                      int32_t nthreads) {
#ifdef _OPENMP
    if (nthreads < 1) nthreads = omp_get_max_threads();
#endif
#pragma omp parallel for schedule(static) num_threads(nthreads) if (nthreads > 1)
    for (int64_t w = 0; w < nw; ++w) out[w] = kmo_logpdf(dn, thetas + w * dn->d);
}

/* ---------------------------------------------------------------- _emcee */

#define KMO_MODE_PHILOX 0
#define KMO_MODE_REPLAY 1

/*
 * The stretch-move ensemble loop, src/samplers.jl:232-293.
 *
 *  x   [nw][d] row-major, in/out (theta0s; the caller passes a copy, :198)
 *  lp  [nw]    in/out (p0s, :209-210)
 *  replay arrays (mode REPLAY) and trace arrays (optional, any mode) are indexed
 *      ((t*2 + batch)*nhalf + i),  t = 0-based outer iteration, batch in {0,1},
 *      i = position of the active walker inside its half.  Partners are GLOBAL 0-based
 *      walker indices (Julia's `no` minus 1).
 *  chain_x [nw][ns][d], chain_lp [nw][ns]  (= Julia's d x ns x nw column-major), may be NULL
 *  naccept [nw] out, accept_ratio [nw] out
 *  min_margin: min over finite decisions of |lhs - log(u)| (how far any decision was from a
 *      tie; a 1-ulp difference between log implementations cannot flip a decision whose
 *      margin is far above 1e-15*|lhs|).
 */
int kmo_emcee(const kmo_density *dn, double *x, double *lp, int64_t nw, int64_t niter_walker,
              int64_t nburnin_walker, int64_t nthin, double a_scale, int32_t mode, uint64_t seed,
              const int64_t *rp_partner, const double *rp_z, const double *rp_u, double *chain_x,
              double *chain_lp, int64_t *naccept, double *accept_ratio, int64_t *tr_partner,
              double *tr_z, double *tr_u, uint8_t *tr_accept, double *min_margin,
              int32_t nthreads) {
    const int d = dn->d;
    if (nw < 2 || (nw & 1) || !(a_scale > 1.0) || nthin < 1 || d < 1 || d > 4096) return 1;
    const int64_t nhalf = nw / 2;
    const int64_t ns = (niter_walker - nburnin_walker) / nthin; /* :234 */
    const double sa = sqrt(a_scale), sia = sqrt(1.0 / a_scale);
    const double span = sa - sia;
    const double nm1 = (double)(d - 1);
    double margin = INFINITY;
#ifdef _OPENMP
    if (nthreads < 1) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
    for (int64_t w = 0; w < nw; ++w) naccept[w] = 0; /* :242 */

    int64_t t = 0;
    for (int64_t n = 1 - nburnin_walker; n <= niter_walker - nburnin_walker; ++n, ++t) { /* :245 */
        for (int batch = 0; batch < 2; ++batch) { /* :246 */
            /* :247 batch 0: active = first half, passive = second half; batch 1: swapped */
            const int64_t a0 = batch == 0 ? 0 : nhalf;
            const int64_t p0 = batch == 0 ? nhalf : 0;
            const int store = (n > 0 && n % nthin == 0);
            const int64_t sidx = store ? n / nthin - 1 : 0;
#pragma omp parallel for schedule(static) num_threads(nthreads) reduction(min : margin) if (nthreads > 1)
            for (int64_t i = 0; i < nhalf; ++i) { /* :248 */
                const int64_t k = a0 + i;
                const int64_t slot = (t * 2 + batch) * nhalf + i;
                int64_t j;
                double z, u;
                if (mode == KMO_MODE_REPLAY) {
                    j = rp_partner[slot];
                    z = rp_z[slot];
                    u = rp_u[slot];
                } else {
                    int64_t pl;
                    double uz;
                    kmo_draw(seed, (uint64_t)k, (uint64_t)t, (uint32_t)batch, (uint64_t)nhalf, &pl,
                             &uz, &u);             /* :250 partner, then :252 z, then :260 u */
                    j = p0 + pl;
                    double s = uz * span + sia;    /* :227 */
                    z = s * s;
                }
                double y[d];
                const double *xk = x + k * d, *xj = x + j * d;
                for (int c = 0; c < d; ++c) y[c] = xj[c] + z * (xk[c] - xj[c]); /* :255 */
                const double p1 = kmo_logpdf(dn, y);                              /* :257 */
                const double lhs = (nm1 * log(z) + p1) - lp[k];                    /* :260 */
                const double lu = log(u);
                const int acc = lhs >= lu;
                if (acc) { /* :261-265 */
                    for (int c = 0; c < d; ++c) x[k * d + c] = y[c];
                    lp[k] = p1;
                    naccept[k] += 1;
                }
                if (isfinite(lhs) && isfinite(lu)) {
                    double mg = fabs(lhs - lu);
                    if (mg < margin) margin = mg;
                }
                if (tr_partner) tr_partner[slot] = j;
                if (tr_z) tr_z[slot] = z;
                if (tr_u) tr_u[slot] = u;
                if (tr_accept) tr_accept[slot] = (uint8_t)acc;
                if (store && chain_x) { /* :268-272 */
                    for (int c = 0; c < d; ++c) chain_x[(k * ns + sidx) * d + c] = x[k * d + c];
                    chain_lp[k * ns + sidx] = lp[k];
                }
            }
        }
        if (n == 0) /* :285-288 */
            for (int64_t w = 0; w < nw; ++w) naccept[w] = 0;
    }
    for (int64_t w = 0; w < nw; ++w) /* :291 */
        accept_ratio[w] = (double)naccept[w] / (double)(niter_walker - nburnin_walker);
    if (min_margin) *min_margin = margin;
    return 0;
}

int32_t kmo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
