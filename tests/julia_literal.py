"""A second, independently written restatement of the reference's `emcee` / `_emcee`
(/root/reference/src/samplers.jl:188-293): a statement-by-statement transliteration of the
Julia text into plain Python -- 1-based walker indices, `push!`-grown per-walker vectors, the
`circshift` of the two half ranges, the `n` counter running from `1-nburnin_walker` -- with no
shared code with oracle/kmc_oracle.c.  TEST INFRASTRUCTURE ONLY (pure-Python loops: small cases).

It exists to pin the C oracle's control flow (store / reset order, loop bounds, `÷` semantics,
half split) against the Julia source text, since Julia itself is not in the image.  Python
floats are IEEE binary64 without contraction and `math.log` / `math.sqrt` are libm's, so the
two restatements must agree bit for bit when fed the same draws.

Draws come from a `source` object in the reference's order per walker-step:
    source.rand_range(lo, hi)   -> Int in lo:hi (1-based, inclusive)      samplers.jl:250
    source.rand()               -> Float64 in [0,1) for sample_g          samplers.jl:230,:252
    source.rand()               -> Float64 in [0,1) for the accept test   samplers.jl:260
"""
from __future__ import annotations

import math


def julia_div(a: int, b: int) -> int:
    """Julia `÷` truncates toward zero (Python `//` floors)."""
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


def julia_rem(a: int, b: int) -> int:
    """Julia `rem` has the sign of the dividend."""
    return a - b * julia_div(a, b)


def g_pdf(z, a):                                        # :224
    return 1 / math.sqrt(z) * 1 / (2 * (math.sqrt(a) - math.sqrt(1 / a))) if 1 / a <= z <= a else 0.0


def cdf_g_inv(u, a):                                    # :227
    s = u * (math.sqrt(a) - math.sqrt(1 / a)) + math.sqrt(1 / a)
    return s * s                                        # x^2 lowers to x*x (Base.literal_pow)


def log_julia(x):
    """Base.log on Float64 for the values the loop can see: log(0.0) = -Inf, NaN propagates."""
    if x != x:
        return x
    if x == 0.0:
        return -math.inf
    return math.log(x)


def emcee(pdf, theta0s, source, niter=10 ** 5, nburnin=None, nthin=1, a_scale=2.0, sample_z=None):
    """`emcee` (:188-216) + `_emcee` (:232-293), hasblob=false, no progress meter.

    theta0s: list of walkers, each a list of floats (Vector theta).  Returns
    (thetas [nw][ns][d], accept_ratio [nw], logdensities [nw][ns], None, naccept, final theta0s, final p0s).
    sample_z: optional callable replacing sample_g (replay of recorded z values, which the
    hot-path ABI uploads instead of the uniform behind them).
    """
    if nburnin is None:
        nburnin = julia_div(niter, 2)                   # :190
    theta0s = [list(t) for t in theta0s]                # :198 deepcopy
    assert a_scale > 1                                  # :200
    nwalkers = len(theta0s)                             # :201
    assert nwalkers % 2 == 0                            # :202
    niter_walker = julia_div(niter, nwalkers)           # :203
    nburnin_walker = julia_div(nburnin, nwalkers)       # :204
    assert nwalkers >= len(theta0s[0]) + 2              # :205
    p0s = [pdf(t) for t in theta0s]                     # :208-210

    nsamples_walker = julia_div(niter_walker - nburnin_walker, nthin)   # :234
    thetas = [[] for _ in range(nwalkers)]              # :237
    logdensities = [[] for _ in range(nwalkers)]        # :239
    naccept = [0] * nwalkers                            # :242
    N = len(theta0s[0])                                 # :243
    half = julia_div(nwalkers, 2)
    ranges = [(1, half), (half + 1, nwalkers)]          # SVector(1:nw÷2, nw÷2+1:nw)
    for n in range(1 - nburnin_walker, niter_walker - nburnin_walker + 1):   # :245
        for batch in (1, 2):                            # :246
            # circshift(v, 1) of a 2-vector swaps the entries; circshift(v, 0) is the identity   :247
            ncs, ncos = (ranges[0], ranges[1]) if batch == 1 else (ranges[1], ranges[0])
            for nc in range(ncs[0], ncs[1] + 1):        # :248
                no = source.rand_range(ncos[0], ncos[1])                     # :250
                z = sample_z() if sample_z is not None else cdf_g_inv(source.rand(), a_scale)   # :252
                xo, xc = theta0s[no - 1], theta0s[nc - 1]
                theta1 = [xo[c] + z * (xc[c] - xo[c]) for c in range(N)]     # :255
                p1 = pdf(theta1)                                             # :257
                if (N - 1) * log_julia(z) + p1 - p0s[nc - 1] >= log_julia(source.rand()):   # :260
                    theta0s[nc - 1] = theta1                                 # :261
                    p0s[nc - 1] = p1                                         # :262
                    naccept[nc - 1] += 1                                     # :265
                if n > 0 and julia_rem(n, nthin) == 0:                       # :268
                    thetas[nc - 1].append(list(theta0s[nc - 1]))             # :269
                    logdensities[nc - 1].append(p0s[nc - 1])                 # :271
        if n == 0:                                      # :285
            naccept = [0] * nwalkers                    # :286
    denom = niter_walker - nburnin_walker
    accept_ratio = [na / denom if denom != 0 else (math.nan if na == 0 else math.inf) for na in naccept]   # :291
    assert all(len(t) == nsamples_walker for t in thetas) or nsamples_walker < 0
    return thetas, accept_ratio, logdensities, None, naccept, theta0s, p0s


class ReplaySource:
    """Feeds recorded draws (0-based global partner index, z, u) back in call order."""

    def __init__(self, partner, z, u):
        self.partner, self.z, self.u = list(partner), list(z), list(u)
        self.i = 0

    def rand_range(self, lo, hi):
        no = int(self.partner[self.i]) + 1
        assert lo <= no <= hi, "recorded partner lies outside the passive half"
        return no

    def sample_z(self):
        return float(self.z[self.i])

    def rand(self):                                     # the accept uniform closes the walker-step
        u = float(self.u[self.i])
        self.i += 1
        return u


# The reference's test densities (test/runtests.jl:52-78, README.md:15) as plain closures, in the
# operation order oracle/kmc_oracle.c documents for its plugins.

def exponential(theta):
    s = 0.0
    for c, v in enumerate(theta):
        if v < 0.0:
            return -math.inf
        s = v if c == 0 else s + v
    return -s


def rosenbrock(theta, a=1.0, b=100.0, temp=20.0):
    t = theta[1] - theta[0] * theta[0]
    r = b * (t * t) + (a - theta[0]) * (a - theta[0])
    return (-r) / temp


def gaussian(params, d):
    mu, A, lognorm = params[:d], params[d:d + d * d], params[d + d * d]

    def f(theta):
        ss = 0.0
        for i in range(d):
            y = 0.0
            for j in range(d):
                y = y + A[i * d + j] * (theta[j] - mu[j])
            ss = ss + y * y
        return lognorm - 0.5 * ss
    return f
