"""GPU tests (`-m gpu`) of the rows either side of the sampler that now run on the device, through the C-ABI, against
the oracle's line-by-line restatements of the reference:

  kmc_emcee_squash   squash_walkers        /root/reference/src/samplers.jl:372-428
  kmc_make_theta0s   make_theta0s          /root/reference/src/samplers.jl:311-349
  kmc_g_pdf / kmc_cdf_g_inv / kmc_sample_g  /root/reference/src/samplers.jl:223-230, test/emcee.jl:2-14
"""
import warnings

import numpy as np
import pytest

from tests import cases

pytestmark = pytest.mark.gpu


# ---------------------------------------------------------------------------------- squash_walkers on the device

def _run(km, case, nw, nitw, nbw, nthin, seed=3, spread=None):
    name, d, params, th0, rad = cases.plugin_specs()[case]
    ld = km.LogDensity(name, d, params)
    x0 = cases.ball(th0, rad, nw, seed)
    if spread is not None:                      # a few walkers start far out: they accept rarely (low accept ratio)
        x0[::17] += spread
    s = km.Sampler(ld, x0, nitw, nbw, nthin, 2.0, seed)
    s.run(-1)
    return s


@pytest.mark.parametrize("case,nw,spread", [("rosenbrock", 512, 30.0), ("mvn10", 1000, 40.0), ("normal", 300, None)])
@pytest.mark.parametrize("drop,order", [(False, False), (False, True), (True, False), (True, True)])
def test_device_squash_matches_reference_loop(km, orc, case, nw, spread, drop, order):
    s = _run(km, case, nw, 60, 20, 3, spread=spread)
    th, lp, ar = s.results()
    want_t, want_a, want_l = orc.squash_walkers(th, ar, lp, drop_low_accept_ratio=drop, drop_fact=1.0, order=order)[:3]
    got_t, got_a, got_l, blobs = s.squash(drop_low_accept_ratio=drop, drop_fact=1.0, order=order)
    nk, med, sd = s.squash_stats
    s.close()
    assert blobs is None
    want_t = np.asarray(want_t).reshape(len(want_l), -1)
    assert got_t.shape == want_t.shape and np.array_equal(got_t, want_t)      # same walkers kept, same order, same bits
    assert np.array_equal(got_l, want_l)
    np.testing.assert_allclose(got_a, want_a, rtol=1e-13)
    assert med == np.median(ar)
    np.testing.assert_allclose(sd, np.std(ar, ddof=1), rtol=1e-12)
    if drop and spread is not None:
        assert nk < nw                           # the far-out walkers were dropped
    if not drop:
        assert nk == nw
    # the host mirror of the reference function agrees as well
    host = km.squash_walkers(th, ar, lp, drop_low_accept_ratio=drop, drop_fact=1.0, order=order, verbose=False)
    assert np.array_equal(host[0].reshape(got_t.shape), got_t)


def test_device_squash_scalar_theta_and_no_samples(km):
    ld = km.exponential()
    x0 = np.abs(cases.ball(0.5, 0.1, 100, 1))
    s = km.Sampler(ld, x0, 40, 40, 1, 2.0, 1)            # burn-in == niter: nothing stored
    s.run(-1)
    t, a, l, _ = s.squash()
    assert t.shape == (0, 1) and l.shape == (0,)
    s.close()


# ---------------------------------------------------------------------------------- make_theta0s on the device

@pytest.mark.parametrize("theta0,radius,case,kw", [
    (0.5, 0.1, "exponential", {}),
    (0.02, 0.5, "exponential", {}),                       # half the ball has zero density: rejections
    ([0.0, 0.0], 0.1, "rosenbrock", {}),
    ([0.4, 0.3], [0.1, 0.2], "mvn2", {}),                 # vector ball radius (:316-319)
    (0.02, 0.5, "lognormal", {}),
    (-2.0, 0.3, "exponential", dict(ntries=5)),           # every k = 1 try fails: the radius-halving path (:324-326)
    (-0.35, 0.3, "exponential", dict(ntries=3)),          # a few walkers need a smaller ball; later walkers are redone
])
def test_device_make_theta0s_matches_reference_loop(km, orc, theta0, radius, case, kw):
    """Device make_theta0s == the sequential loop of src/samplers.jl:311-349 fed the normals the device used."""
    name, d, params, *_ = cases.plugin_specs()[case]
    ld, od = km.LogDensity(name, d, params), orc.Density(name, d, params)
    nw, seed = 300, 99
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = km.make_theta0s(theta0, radius, ld, nw, seed=seed, **kw)
    randn = lambda i, k, j: km.ball_randn_device(seed, i, 1, k, j, d)[0]
    want = orc.make_theta0s(theta0, radius, od.logpdf, nw, randn, **kw)
    assert got.shape == want.shape
    assert np.array_equal(got, want)
    assert got.ndim == (1 if np.ndim(theta0) == 0 else 2)
    if len(got):
        assert np.all(od.eval(np.reshape(got, (len(got), -1))) > -np.inf)


def test_device_make_theta0s_batched_plugin_and_statistics(km):
    """A batched plugin (dense Gaussian d = 40 goes through the wide FP64 kernel) and the ball's moments at scale."""
    d = 40
    ld = km.gaussian(np.zeros(d), cases.spd_cov(d, 2))
    x = km.make_theta0s(np.linspace(-1, 1, d), 0.25, ld, 20000, seed=5)
    assert x.shape == (20000, d)
    np.testing.assert_allclose(x.mean(0), np.linspace(-1, 1, d), atol=0.01)
    np.testing.assert_allclose(x.std(0), 0.25, rtol=0.03)
    z = km.ball_randn_device(5, 0, 20000, 1, 1, d)
    assert np.array_equal(x, np.linspace(-1, 1, d) + z * 0.25)             # :328-332: mul, then add
    assert abs(np.corrcoef(z[:, 0], z[:, 1])[0, 1]) < 0.03                 # the two Box-Muller outputs are independent
    assert np.array_equal(z[7:9], km.ball_randn_device(5, 7, 2, 1, 1, d))  # a pure function of the counter


def test_make_theta0s_through_emcee_on_device(km):
    ld = km.exponential()
    th, ar, lp, _ = km.emcee(ld, km.make_theta0s(0.5, 0.1, ld, 100, seed=2), niter=10**5, use_progress_meter=False)
    assert th.shape == (100, 500) and ar.mean() > 0.5


# ---------------------------------------------------------------------------------- the g distribution

def test_g_dist_like_the_reference(km, orc):
    """test/emcee.jl:2-14 on the device draw path, plus bit-equality of g_pdf / cdf_g_inv with the oracle."""
    a = 3.5
    z = km.sample_g(a, 50000, seed=11)
    assert np.all((z >= 1 / a) & (z <= a))
    assert np.isclose(km.cdf_g_inv(1.0, a), a) and np.isclose(km.cdf_g_inv(0.0, a), 1 / a)
    grid = np.arange(1 / a, a, 0.01)
    pdf = km.g_pdf(grid, a)
    mean = np.sum(grid * pdf) * 0.01
    std = np.sqrt(np.sum((grid - mean) ** 2 * pdf) * 0.01)
    assert abs(z.mean() - mean) < 1e-2 and abs(z.std() - std) < 1e-2
    L = orc.lib()
    for v in (0.1, 1 / a, 0.5, 1.0, 2.0, a, 3.6):
        assert km.g_pdf(v, a) == L.kmo_g_pdf(v, a)
    for u in (0.0, 0.25, 0.999, 1.0):
        assert km.cdf_g_inv(u, a) == L.kmo_cdf_g_inv(u, a)
    assert km.g_pdf(0.1, a) == 0.0 and km.g_pdf(3.6, a) == 0.0
    assert isinstance(km.sample_g(2.0, seed=3), float)
    # the device z of sample i is cdf_g_inv of the walker-step uniform of (walker i, iteration 0, batch 0)
    uz = np.array([orc.draw(11, i, 0, 0, 2)[1] for i in range(8)])
    assert np.array_equal(z[:8], km.cdf_g_inv(uz, a))
