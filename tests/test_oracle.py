"""CPU tests (no GPU): the oracle against the reference's own known-answer tests and against
the committed golden fixtures.  Reference citations are into /root/reference."""
import math
from pathlib import Path

import numpy as np
import pytest

from tests import cases

GOLDEN = Path(__file__).resolve().parent / "golden"


def test_philox_known_answers(orc):
    # Random123 kat_vectors for philox4x32-10
    assert orc.philox4x32_10([0] * 4, [0] * 2) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert orc.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert orc.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344],
                             [0xa4093822, 0x299f31d0]) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_draw_ranges_and_partner_uniformity(orc):
    nhalf = 7
    counts = np.zeros(nhalf)
    for w in range(4000):
        pl, uz, ua = orc.draw(5, w, 3, 1, nhalf)
        assert 0 <= pl < nhalf and 0.0 <= uz < 1.0 and 0.0 <= ua < 1.0
        counts[pl] += 1
    assert np.all(np.abs(counts - 4000 / nhalf) < 5 * math.sqrt(4000 / nhalf))


def test_g_dist(orc):
    """test/emcee.jl:2-14, with the oracle's cdf_g_inv / g_pdf."""
    a = 3.5
    u = np.array([orc.draw(1, w, 0, 0, 2)[1] for w in range(50000)])
    samples = np.array([orc.cdf_g_inv(v, a) for v in u])
    assert np.all((1 / a <= samples) & (samples <= a))
    assert orc.cdf_g_inv(1, a) == pytest.approx(a)
    assert orc.cdf_g_inv(0, a) == pytest.approx(1 / a)
    z = np.arange(1 / a, a, 0.01)
    g = np.array([orc.g_pdf(v, a) for v in z])
    meang = np.sum(z * g) * 0.01
    assert abs(samples.mean() - meang) < 1e-2
    stdg = math.sqrt(np.sum((meang - z) ** 2 * g) * 0.01)
    assert abs(samples.std(ddof=1) - stdg) < 1e-2
    # analytic moments at a=2: E[z]=7/6, std=0.43461
    s2 = np.array([orc.cdf_g_inv(v, 2.0) for v in u])
    assert abs(s2.mean() - 7 / 6) < 1e-2 and abs(s2.std() - 0.43461) < 1e-2
    assert orc.g_pdf(0.1, a) == 0.0 and orc.g_pdf(4.0, a) == 0.0


def _moments_ok(thetas, mean, std, tol, median=None, skew=None):
    """test/runtests.jl:36-43 test_mean_std."""
    thetas = np.asarray(thetas)
    std = np.asarray(std, dtype=float)
    assert np.all(np.abs(thetas.mean(axis=0) - mean) < np.abs(std * tol))
    assert np.all(np.abs(thetas.std(axis=0, ddof=1) - std) < np.abs(std * tol))
    if median is not None:
        assert abs(np.median(thetas) - median) < abs(std * tol)
    if skew is not None:
        c = thetas - thetas.mean()
        sk = np.mean(c ** 3) / np.mean(c ** 2) ** 1.5
        assert abs(sk - skew) < abs(std * 2 * tol)


# the reference's testcases, test/runtests.jl:52-78 (blob cases are out of scope)
REFERENCE_CASES = [
    # case, niter, tol, mean, std, median, skewness
    ("normal", 10**4, 0.3, -5.0, 3.0, -5.0, 0.0),
    ("lognormal", 10**7, 0.3, math.exp(0.5), math.sqrt((math.e - 1) * math.e), 1.0,
     (math.e + 2) * math.sqrt(math.e - 1)),
    ("mvn2", 10**5, 0.3, [0.5, -0.25], [math.sqrt(0.47), math.sqrt(7.0)], None, None),
    ("rosenbrock", 10**7, 0.6, [0.98, 10.3], [3.1, 13.8], None, None),
]


@pytest.mark.parametrize("case,niter,tol,mean,std,median,skew", REFERENCE_CASES)
def test_oracle_reference_statistics(orc, case, niter, tol, mean, std, median, skew):
    """test/emcee.jl:17-48 run through the oracle: shapes, accept ratio, moments."""
    name, d, params, th0, rad = cases.plugin_specs()[case]
    nw = 100
    dens = orc.Density(name, d, params)
    x0 = cases.ball(th0, rad, nw, 3)
    nitw = niter // nw
    r = orc.emcee(dens, x0, nitw, nitw // 2, 1, 2.0, seed=42, nthreads=4)
    assert r["chain_x"].shape == (nw, niter // nw // 2, d)           # test/emcee.jl:29,35
    t, ar, l, _ = orc.squash_walkers(r["chain_x"], r["accept_ratio"], r["chain_lp"])
    assert len(t) == niter // 2 and len(l) == niter // 2             # :41-42
    assert ar > 0.1                                                  # :43
    _moments_ok(t if d > 1 else t[:, 0], mean, std, tol, median, skew)


def test_oracle_readme_exponential(orc):
    """README.md:15,25 -- Exp(1): mean 1, std 1; acceptance ~0.746 at a=2."""
    dens = orc.Density("exponential", 1)
    x0 = np.abs(cases.ball(0.5, 0.1, 100, 1))
    r = orc.emcee(dens, x0, 1000, 500, 1, 2.0, seed=7)
    t, ar, _, _ = orc.squash_walkers(r["chain_x"], r["accept_ratio"])
    assert t.shape == (50000, 1) and np.all(t >= 0)
    assert abs(t.mean() - 1) < 0.1 and abs(t.std() - 1) < 0.1
    assert abs(ar - 0.746) < 0.02


@pytest.mark.parametrize("nitw,nbw,nthin", [(10, 5, 1), (10, 0, 3), (7, 3, 2), (5, 5, 1), (0, 0, 1), (9, 2, 4)])
def test_oracle_sample_counts(orc, nitw, nbw, nthin):
    """ns = ((niter/nw) - (nburnin/nw)) / nthin  (src/samplers.jl:234) and the burn-in reset (:285-288)."""
    dens = orc.Density("rosenbrock", 2, [1, 100, 20])
    r = orc.emcee(dens, cases.ball([0, 0], 0.1, 8, 0), nitw, nbw, nthin, 2.0, seed=1, trace=True)
    ns = (nitw - nbw) // nthin
    assert r["chain_x"].shape == (8, ns, 2)
    acc = r["trace"][3].reshape(nitw, 2, 4)
    post = acc[nbw:]      # decisions after the counters were reset at n == 0
    want = np.concatenate([post[:, 0, :].sum(0), post[:, 1, :].sum(0)])
    assert np.array_equal(r["naccept"], want)


def test_oracle_replay_equals_philox(orc):
    """Feeding the oracle its own trace back reproduces the run bit-for-bit."""
    name, d, params, th0, rad = cases.plugin_specs()["mvn2"]
    dens = orc.Density(name, d, params)
    x0 = cases.ball(th0, rad, 12, 5)
    a = orc.emcee(dens, x0, 25, 10, 2, 2.0, seed=9, trace=True)
    b = orc.emcee(dens, x0, 25, 10, 2, 2.0, replay=a["trace"][:3], trace=True)
    for k in ("chain_x", "chain_lp", "accept_ratio", "x", "lp"):
        assert np.array_equal(a[k], b[k])
    assert np.array_equal(a["trace"][3], b["trace"][3])


def test_oracle_threads_do_not_change_results(orc):
    name, d, params, th0, rad = cases.plugin_specs()["rosenbrock"]
    dens = orc.Density(name, d, params)
    x0 = cases.ball(th0, rad, 64, 5)
    a = orc.emcee(dens, x0, 30, 10, 1, 2.0, seed=2, nthreads=1)
    b = orc.emcee(dens, x0, 30, 10, 1, 2.0, seed=2, nthreads=4)
    assert np.array_equal(a["chain_x"], b["chain_x"]) and np.array_equal(a["naccept"], b["naccept"])


def test_accept_edge_cases(orc):
    """src/samplers.jl:260: `>=` (not `>`); p1=-Inf rejects unless u==0; NaN rejects."""
    dens = orc.Density("exponential", 1)
    x0 = np.array([[1.0], [6.0], [3.0], [4.0]])
    partner, z, u = edge_case_draws()
    r = orc.emcee(dens, x0, 1, 0, 1, 2.0, replay=(partner, z, u), trace=True)
    assert list(r["trace"][3]) == [0, 1, 1, 0]
    assert r["x"][:, 0].tolist() == [1.0, 6.0, -1.5, 4.0]
    assert r["lp"].tolist() == [-1.0, -6.0, -np.inf, -4.0]
    assert r["naccept"].tolist() == [0, 1, 1, 0]


def edge_case_draws():
    """batch 0: walkers 0,1 active (partners 2,3); batch 1: walkers 2,3 active (partners 1,0).
    w0: y = 3+2(1-3) = -1 -> p1=-Inf, u=.5 -> reject.   w1: z=1 -> y=x, lhs = 0 >= log(1)=0 -> accept on equality.
    w2: y = 6+2.5(3-6) = -1.5 -> p1=-Inf, u=0 -> -Inf >= -Inf -> accept.   w3: u=NaN -> reject."""
    return np.array([2, 3, 1, 0]), np.array([2.0, 1.0, 2.5, 0.5]), np.array([0.5, 1.0, 0.0, np.nan])


@pytest.mark.parametrize("case", ["exponential", "exponential3", "rosenbrock", "normal", "mvn2", "mvn10", "lognormal"])
def test_oracle_reproduces_golden(orc, case):
    g = np.load(GOLDEN / f"{case}.npz")
    dens = orc.Density(str(g["name"]), int(g["d"]), g["params"])
    args = (int(g["niter_walker"]), int(g["nburnin_walker"]), int(g["nthin"]), float(g["a_scale"]))
    r = orc.emcee(dens, g["theta0s"], *args, seed=int(g["seed"]), trace=True)
    assert np.array_equal(r["trace"][0], g["partner"]) and np.array_equal(r["trace"][1], g["z"])
    assert np.array_equal(r["trace"][2], g["u"]) and np.array_equal(r["trace"][3], g["accept"])
    assert np.array_equal(r["chain_x"], g["chain_x"]) and np.array_equal(r["chain_lp"], g["chain_lp"])
    assert np.array_equal(r["accept_ratio"], g["accept_ratio"])
    r2 = orc.emcee(dens, g["theta0s"], *args, replay=(g["partner"], g["z"], g["u"]))
    assert np.array_equal(r2["chain_x"], g["chain_x"]) and np.array_equal(r2["x"], g["final_x"])
    assert float(g["min_margin"]) > 1e-9     # no decision anywhere near a tie


def test_gaussian_matches_scipy(orc):
    from scipy.stats import multivariate_normal, norm, lognorm
    name, d, params, *_ = cases.plugin_specs()["mvn2"]
    dens = orc.Density(name, d, params)
    pts = np.random.default_rng(0).standard_normal((20, 2)) * 3
    want = multivariate_normal(cases.MVN_MEAN, cases.MVN_COV).logpdf(pts)
    np.testing.assert_allclose(dens.eval(pts), want, rtol=1e-12, atol=1e-12)
    name, d, params, *_ = cases.plugin_specs()["normal"]
    np.testing.assert_allclose(orc.Density(name, d, params).eval(pts[:, :1]), norm(-5, 3).logpdf(pts[:, 0]), rtol=1e-12)
    name, d, params, *_ = cases.plugin_specs()["lognormal"]
    xs = np.abs(pts[:, :1]) + 0.01
    np.testing.assert_allclose(orc.Density(name, d, params).eval(xs), lognorm(1.0).logpdf(xs[:, 0]), rtol=1e-12, atol=1e-13)
    assert np.isneginf(orc.Density(name, d, params).logpdf([-1.0])) and np.isneginf(orc.Density(name, d, params).logpdf([0.0]))


@pytest.mark.parametrize("case,nw,niter,nburnin,nthin,a", [
    ("exponential", 6, 6 * 9, 6 * 4, 2, 2.0),       # burn-in reset at n == 0, thinning
    ("exponential3", 10, 10 * 7 + 3, 10 * 2 + 9, 1, 2.5),   # niter, nburnin not multiples of nw (÷ truncation)
    ("rosenbrock", 8, 8 * 12, 0, 3, 2.0),           # no burn-in: the counters are never reset
    ("mvn2", 12, 12 * 8, 12 * 8, 1, 1.5),           # everything is burn-in: zero samples, accept_ratio = 0/0
    ("rosenbrock", 4, 4 * 5, 4 * 2, 4, 2.0),        # nthin larger than the post-burn-in run: zero samples
])
def test_oracle_equals_literal_julia_transliteration(orc, case, nw, niter, nburnin, nthin, a):
    """The C oracle against tests/julia_literal.py, an independent statement-by-statement transliteration of
    src/samplers.jl:188-293 (1-based ranges, circshift, push!, total-step arguments), on the same draws: bit-identical
    chains, log-densities, decisions, counters and final state."""
    from tests import julia_literal as jl
    name, d, params, th0, rad = cases.plugin_specs()[case]
    x0 = cases.ball(th0, rad, nw, 3)
    nitw, nbw = niter // nw, nburnin // nw
    want = orc.emcee(orc.Density(name, d, params), x0, nitw, nbw, nthin, a, seed=11, trace=True)
    pdf = {"exponential": lambda: jl.exponential, "rosenbrock": lambda: (lambda t: jl.rosenbrock(t, *params)),
           "gaussian": lambda: jl.gaussian(list(np.asarray(params, dtype=np.float64)), d)}[name]()
    src = jl.ReplaySource(*want["trace"][:3])
    th, ar, lps, blobs, nacc, xfin, pfin = jl.emcee(pdf, x0.tolist(), src, niter=niter, nburnin=nburnin, nthin=nthin,
                                                    a_scale=a, sample_z=src.sample_z)
    assert blobs is None and src.i == nitw * nw
    ns = (nitw - nbw) // nthin
    assert all(len(t) == ns for t in th)
    assert np.array_equal(np.array(th, dtype=np.float64).reshape(nw, ns, d), want["chain_x"])
    assert np.array_equal(np.array(lps, dtype=np.float64).reshape(nw, ns), want["chain_lp"])
    assert nacc == want["naccept"].tolist()
    assert np.array_equal(np.array(ar), want["accept_ratio"], equal_nan=True)
    assert np.array_equal(np.array(xfin), want["x"]) and np.array_equal(np.array(pfin), want["lp"])


def test_literal_g_helpers_equal_oracle(orc):
    """g_pdf / cdf_g_inv (src/samplers.jl:224,:227): literal transliteration == C oracle, bit for bit."""
    from tests import julia_literal as jl
    rng = np.random.default_rng(4)
    for a in (1.5, 2.0, 3.7):
        for u in list(rng.random(50)) + [0.0, 1.0 - 2.0 ** -53]:
            z = jl.cdf_g_inv(u, a)
            assert z == orc.cdf_g_inv(u, a)
            assert jl.g_pdf(z, a) == orc.g_pdf(z, a)
        assert jl.g_pdf(a * 1.0001, a) == 0.0 == orc.g_pdf(a * 1.0001, a)
