"""GPU parity tests (`-m gpu`): the CUDA path, called through the C-ABI, against the CPU oracle
and the committed golden fixtures.  Bit-exact on decisions, partner indices, chains and
log-densities for the scalar FP64 plugins (tolerance 0): both sides perform the same IEEE
binary64 operations without FMA contraction.  The only non-shared arithmetic is log(): the
fixtures record the smallest |lhs - log u| of any decision (>= 1e-9), so a 1-ulp difference
between glibc and CUDA log cannot flip a decision."""
import math
from pathlib import Path

import numpy as np
import pytest

from tests import cases
from tests.test_oracle import REFERENCE_CASES, _moments_ok, edge_case_draws

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden"
ALL_CASES = ["exponential", "exponential3", "rosenbrock", "normal", "mvn2", "mvn10", "lognormal"]


def _pair(km, orc, case):
    name, d, params, th0, rad = cases.plugin_specs()[case]
    return km.LogDensity(name, d, params), orc.Density(name, d, params), th0, rad


def _start(case, th0, rad, nw, seed):
    x = cases.ball(th0, rad, nw, seed)
    return np.abs(x) if case.startswith(("exponential", "lognormal")) else x


def _run_gpu(km, ld, x0, nitw, nbw, nthin, a, seed=0, replay=None, launch_mode=0, chunks=None):
    mode = km.MODE_REPLAY if replay is not None else km.MODE_PHILOX
    s = km.Sampler(ld, x0, nitw, nbw, nthin, a, seed, mode, launch_mode=launch_mode)
    if replay is not None:
        s.set_replay(*replay)
    if chunks:
        for c in chunks:
            s.run(c)
    s.run(-1)
    th, lp, ar = s.results()
    x, l, na = s.state()
    s.close()
    return dict(chain_x=th, chain_lp=lp, accept_ratio=ar, x=x, lp=l, naccept=na)


def _assert_same(a, b):
    for k in ("chain_x", "chain_lp", "accept_ratio", "x", "lp", "naccept"):
        assert np.array_equal(a[k], b[k], equal_nan=True), k


def test_device_present(km):
    assert km.device_count() >= 1


@pytest.mark.parametrize("case", ALL_CASES)
def test_density_eval_matches_oracle(km, orc, case):
    ld, od, th0, rad = _pair(km, orc, case)
    rng = np.random.default_rng(0)
    pts = np.atleast_1d(np.asarray(th0, dtype=float))[None, :] + 3.0 * rng.standard_normal((1000, ld.d))
    pts[0] = 0.0
    pts[1] = -1.0
    pts[2] = np.nan
    pts[3] = 1e300
    got, want = ld.eval(pts), od.eval(pts)
    if case == "lognormal":     # log() is the one routine not shared bit-for-bit (CUDA vs glibc, <= 1 ulp each)
        fin = np.isfinite(want)
        assert np.array_equal(np.isfinite(got), fin) and np.array_equal(got[~fin], want[~fin], equal_nan=True)
        np.testing.assert_allclose(got[fin], want[fin], rtol=1e-14, atol=1e-15)
    else:
        assert np.array_equal(got, want, equal_nan=True)
    assert ld(pts[5]) == want[5] if ld.d > 1 else ld(float(pts[5, 0])) == want[5]


@pytest.mark.parametrize("case", ALL_CASES)
def test_golden_replay(km, case):
    """Replay mode against the committed fixture (no oracle involved): uploaded partner/z/u
    must give bit-identical chains, log-densities, accept counts and final state."""
    g = np.load(GOLDEN / f"{case}.npz")
    ld = km.LogDensity(str(g["name"]), int(g["d"]), g["params"])
    r = _run_gpu(km, ld, g["theta0s"], int(g["niter_walker"]), int(g["nburnin_walker"]), int(g["nthin"]),
                 float(g["a_scale"]), replay=(g["partner"], g["z"], g["u"]))
    tol = dict(rtol=1e-14, atol=0) if case == "lognormal" else None
    for k, gk in (("chain_x", "chain_x"), ("chain_lp", "chain_lp"), ("accept_ratio", "accept_ratio"),
                  ("x", "final_x"), ("lp", "final_lp"), ("naccept", "naccept")):
        if tol and k in ("chain_lp", "lp"):
            np.testing.assert_allclose(r[k], g[gk], **tol)
        else:
            assert np.array_equal(r[k], g[gk]), k


@pytest.mark.parametrize("case", ALL_CASES)
def test_golden_philox(km, case):
    """Free-running mode: the device Philox stream, partner mapping and z transform reproduce the
    fixture's recorded draws, so the whole run is bit-identical."""
    g = np.load(GOLDEN / f"{case}.npz")
    ld = km.LogDensity(str(g["name"]), int(g["d"]), g["params"])
    r = _run_gpu(km, ld, g["theta0s"], int(g["niter_walker"]), int(g["nburnin_walker"]), int(g["nthin"]),
                 float(g["a_scale"]), seed=int(g["seed"]))
    assert np.array_equal(r["chain_x"], g["chain_x"]) and np.array_equal(r["naccept"], g["naccept"])
    assert np.array_equal(r["x"], g["final_x"])
    if case == "lognormal":
        np.testing.assert_allclose(r["chain_lp"], g["chain_lp"], rtol=1e-14)
    else:
        assert np.array_equal(r["chain_lp"], g["chain_lp"])


@pytest.mark.parametrize("case,nw,nitw,nbw,nthin,a", [
    ("exponential", 100, 1000, 500, 1, 2.0),        # README config (BASELINE.json configs[0])
    ("rosenbrock", 4096, 200, 100, 7, 2.0),
    ("rosenbrock", 4, 50, 0, 1, 2.0),               # minimum ensemble: nw = d + 2
    ("mvn2", 1000, 120, 40, 3, 3.5),
    ("mvn10", 512, 60, 20, 2, 2.0),
    ("exponential3", 300, 80, 80, 1, 2.0),          # burn-in == niter -> no samples
    ("normal", 50, 0, 0, 1, 2.0),                   # niter < nwalkers -> niter_walker = 0
])
def test_seeded_oracle_parity(km, orc, case, nw, nitw, nbw, nthin, a):
    """Same seed, same inputs: oracle (Philox) vs CUDA (Philox) vs CUDA (replay of the oracle's trace)."""
    ld, od, th0, rad = _pair(km, orc, case)
    x0 = _start(case, th0, rad, nw, 21)
    want = orc.emcee(od, x0, nitw, nbw, nthin, a, seed=99, trace=True, nthreads=4)
    if want["trace"][0].size:
        assert want["min_margin"] > 1e-11
    got = _run_gpu(km, ld, x0, nitw, nbw, nthin, a, seed=99)
    _assert_same(got, want)
    if nitw:
        rep = _run_gpu(km, ld, x0, nitw, nbw, nthin, a, replay=want["trace"][:3])
        _assert_same(rep, want)
    assert got["chain_x"].shape == (nw, (nitw - nbw) // nthin, ld.d)


def test_launch_modes_and_chunking_agree(km):
    """Persistent kernel (grid barrier) == one launch per half-step == arbitrary run() chunking."""
    ld = km.rosenbrock()
    x0 = cases.ball([0, 0], 0.1, 20000, 3)
    a = _run_gpu(km, ld, x0, 60, 25, 4, 2.0, seed=5, launch_mode=0)
    b = _run_gpu(km, ld, x0, 60, 25, 4, 2.0, seed=5, launch_mode=1)
    c = _run_gpu(km, ld, x0, 60, 25, 4, 2.0, seed=5, launch_mode=0, chunks=[1, 7, 24, 1, 13])
    _assert_same(a, b)
    _assert_same(a, c)


def test_accept_edge_cases_on_device(km, orc):
    """src/samplers.jl:260 `>=`, -Inf proposals, u == 0, NaN -- same decisions as the oracle."""
    x0 = np.array([[1.0], [6.0], [3.0], [4.0]])
    r = _run_gpu(km, km.exponential(), x0, 1, 0, 1, 2.0, replay=edge_case_draws())
    assert r["x"][:, 0].tolist() == [1.0, 6.0, -1.5, 4.0]
    assert r["lp"].tolist() == [-1.0, -6.0, -np.inf, -4.0]
    assert r["naccept"].tolist() == [0, 1, 1, 0]


def test_replay_validation(km):
    ld = km.rosenbrock()
    x0 = cases.ball([0, 0], 0.1, 8, 0)
    s = km.Sampler(ld, x0, 4, 0, 1, 2.0, 0, km.MODE_REPLAY)
    with pytest.raises(km.KmcError):
        s.run(1)                                   # no draws uploaded
    bad = np.full(8, 99, dtype=np.int64)
    with pytest.raises(km.KmcError):
        s.set_replay(bad, np.ones(8), np.ones(8))  # partner outside the ensemble
    s.close()
    with pytest.raises(km.KmcError, match="even number"):
        km.Sampler(ld, x0[:7], 4, 0)
    with pytest.raises(km.KmcError, match="a_scale"):
        km.Sampler(ld, x0, 4, 0, a_scale=0.5)
    with pytest.raises(km.KmcError, match="DOF"):
        km.Sampler(ld, x0[:2], 4, 0)


@pytest.mark.parametrize("case,niter,tol,mean,std,median,skew", REFERENCE_CASES)
def test_reference_testcases_through_public_api(km, case, niter, tol, mean, std, median, skew):
    """The reference's own emcee testset (test/emcee.jl:17-48) through make_theta0s -> emcee ->
    squash_walkers: shapes, 4th output None, accept_ratio > 0.1, moments within tol*std."""
    name, d, params, th0, rad = cases.plugin_specs()[case]
    ld = km.LogDensity(name, d, params)
    nw = 100
    theta0s = km.make_theta0s(th0, rad, ld, nw, seed=4)
    samples = km.emcee(ld, theta0s, niter=niter, use_progress_meter=False, seed=8)
    assert tuple(len(v) for v in samples[:3]) == (nw, nw, nw)
    assert samples[3] is None
    assert len(samples[0][0]) == niter // nw // 2
    thetas, ar, logd, blobs = km.squash_walkers(*samples, verbose=False)
    assert blobs is None and len(thetas) == niter // 2 and len(logd) == niter // 2
    assert ar > 0.1
    _moments_ok(thetas, mean, std, tol, median, skew)


def test_readme_example(km, capsys):
    """README.md:12-34 (BASELINE.json configs[0]) with the default progress meter on."""
    ld = km.exponential()
    thetas, ar, _, _ = km.emcee(ld, km.make_theta0s(0.5, 0.1, ld, 100), niter=10**5)
    t, a, _, _ = km.squash_walkers(thetas, ar)
    assert thetas.shape == (100, 500) and t.shape == (50000,)
    assert np.all(t >= 0) and abs(t.mean() - 1) < 0.1 and abs(t.std() - 1) < 0.1
    assert abs(a - 0.746) < 0.03
    assert "accept_ratio_mean" in capsys.readouterr().err


def test_progress_statistics(km):
    """kmc_emcee_progress == mean / sqrt(var) / outlier count of naccept (src/samplers.jl:276-278)."""
    ld = km.rosenbrock()
    s = km.Sampler(ld, cases.ball([0, 0], 0.1, 5000, 1), 80, 0, 1, 2.0, 3)
    s.run(50)
    it, mean, sd, outl = s.progress()
    _, _, na = s.state()
    assert it == 50 and mean == pytest.approx(na.mean(), rel=1e-12)
    assert sd == pytest.approx(na.std(ddof=1), rel=1e-9)
    assert outl == int(np.sum(np.abs(na - na.mean()) > 2 * na.std(ddof=1)))
    s.close()


def test_full_size_properties_config2(km):
    """BASELINE.json configs[1] at full width (2^20 walkers, Rosenbrock) for 64 iterations, checked
    through size-independent properties: (1) every stored log-density equals the plugin evaluated
    at the stored theta; (2) per-walker accept counts equal the number of state changes in the
    unthinned chain; (3) persistent and per-half-step launch modes give identical checksums."""
    ld = km.rosenbrock()
    nw = 1 << 20
    x0 = cases.ball([0, 0], 0.1, nw, 77)
    s = km.Sampler(ld, x0, 64, 32, 1, 2.0, 1234)
    s.run(-1)
    th, lp, ar = s.results()
    xf, lf, na = s.state()
    s.close()
    assert th.shape == (nw, 32, 2)
    sub = slice(0, nw, 37)
    assert np.array_equal(ld.eval(th[sub].reshape(-1, 2)), lp[sub].reshape(-1))
    assert np.array_equal(th[:, -1], xf) and np.array_equal(lp[:, -1], lf)
    changes = np.sum(np.any(th[:, 1:] != th[:, :-1], axis=2), axis=1)
    assert np.all(changes <= na) and np.all(na - changes <= 1)      # the first post-burn-in move has no predecessor stored
    assert 0.2 < ar.mean() < 0.8
    s2 = km.Sampler(ld, x0, 64, 32, 1, 2.0, 1234, launch_mode=1)
    s2.run(-1)
    x2, l2, na2 = s2.state()
    s2.close()
    assert np.array_equal(xf, x2) and np.array_equal(lf, l2) and np.array_equal(na, na2)


def test_device_chain_moments(km):
    """kmc_emcee_chain_moments == numpy mean / var (ddof=1) of the squashed chain."""
    ld = km.gaussian([0.5, -0.25], [[0.47, 1.8], [1.8, 7.0]])
    s = km.Sampler(ld, cases.ball([0.4, 0.3], 0.1, 1000, 2), 300, 100, 2, 2.0, 5)
    s.run(-1)
    mean, var, n = s.chain_moments()
    th, _, ar = s.results()
    s.close()
    t, _, _, _ = km.squash_walkers(th, ar)
    assert n == len(t) == 1000 * 100
    np.testing.assert_allclose(mean, t.mean(0), rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(var, t.var(0, ddof=1), rtol=1e-10)


@pytest.mark.parametrize("plugin,d", [("exponential", 2), ("exponential", 4), ("exponential", 5), ("exponential", 6),
                                      ("exponential", 8), ("exponential", 16), ("gaussian", 3), ("gaussian", 4),
                                      ("gaussian", 5), ("gaussian", 6), ("gaussian", 8), ("gaussian", 12),
                                      ("gaussian", 16)])
def test_every_compiled_dimension_matches_oracle(km, orc, plugin, d):
    """Every (plugin, d) instantiation, on every kernel family it can take -- shared-memory kernel
    (d <= 4, small ensembles), bulk/TMA kernel (even d >= 6, Philox), general kernel (odd d, replay,
    one launch per half-step) -- bit-identical to the oracle in Philox and replay mode."""
    if plugin == "exponential":
        params, x0 = [], np.abs(cases.ball(np.full(d, 1.0), 0.3, 600, d))
    else:
        params = cases.gaussian_params(np.linspace(-1, 1, d), cases.spd_cov(d, d))
        x0 = cases.ball(np.zeros(d), 0.5, 600, d)
    ld, od = km.LogDensity(plugin, d, params), orc.Density(plugin, d, params)
    want = orc.emcee(od, x0, 14, 5, 3, 2.0, seed=31 + d, trace=True, nthreads=4)
    assert want["min_margin"] > 1e-11
    for kw in (dict(), dict(launch_mode=1), dict(replay=want["trace"][:3])):
        got = _run_gpu(km, ld, x0, 14, 5, 3, 2.0, seed=31 + d, **kw)
        _assert_same(got, want)


def test_large_ensemble_general_kernels(km, orc):
    """Ensembles too large for the shared-memory kernel: 2-D Rosenbrock with 2^21 walkers (general kernel)
    and a 10-D Gaussian with 2^18 walkers (bulk kernel, several groups per CTA, partial last group) agree
    with one-launch-per-half-step runs and, on a sub-sampled oracle check, with the plugin itself."""
    ld = km.rosenbrock()
    x0 = cases.ball([0, 0], 0.1, 1 << 21, 9)
    a = _run_gpu(km, ld, x0, 6, 2, 2, 2.0, seed=3)
    b = _run_gpu(km, ld, x0, 6, 2, 2, 2.0, seed=3, launch_mode=1)
    _assert_same(a, b)
    prm = cases.gaussian_params(np.linspace(-1, 1, 10), cases.spd_cov(10, 1))
    ld = km.LogDensity("gaussian", 10, prm)
    x0 = cases.ball(np.zeros(10), 0.3, (1 << 18) + 2 * 77, 4)
    a = _run_gpu(km, ld, x0, 6, 2, 2, 2.0, seed=5)
    b = _run_gpu(km, ld, x0, 6, 2, 2, 2.0, seed=5, launch_mode=1)
    _assert_same(a, b)
    od = orc.Density("gaussian", 10, prm)
    sub = slice(0, None, 997)
    assert np.array_equal(od.eval(a["x"][sub]), a["lp"][sub])


def test_rosenbrock_analytic_moments_at_scale(km):
    """Free-running check against the ANALYTIC posterior of the reference's Rosenbrock/20 test density
    (test/runtests.jl:68): y|x ~ N(x^2, 0.1), x ~ N(1, 10)  =>  E = [1, 11], std = [sqrt(10), sqrt(240.1)].
    (The reference quotes [0.98, 10.3] / [3.1, 13.8] from its own 10^9-step run, test/runtests.jl:70-72, and
    tests within 0.6*std of those.)  9.8e8 walker-steps, moments reduced on the device; the tails mix
    slowly (acceptance 0.22), hence the 3-4 % bands."""
    ld = km.rosenbrock()
    nw, nitw = 1 << 14, 60000
    x0 = km.make_theta0s(np.array([0.0, 0.0]), 0.1, ld, nw, seed=1)
    s = km.Sampler(ld, x0, nitw, nitw // 3, 50, 2.0, seed=11)
    s.run(-1)
    mean, var, n = s.chain_moments()
    s.close()
    std = np.sqrt(var)
    assert n == nw * ((nitw - nitw // 3) // 50)
    assert abs(mean[0] - 1.0) < 0.1 and abs(std[0] - 10 ** 0.5) < 0.12
    assert abs(mean[1] - 11.0) < 0.7 and abs(std[1] - 240.1 ** 0.5) < 0.9
    # and inside the reference's own acceptance band around its quoted values
    assert np.all(np.abs(mean - [0.98, 10.3]) < 0.6 * np.array([3.1, 13.8]))
    assert np.all(np.abs(std - [3.1, 13.8]) < 0.6 * np.array([3.1, 13.8]))


@pytest.mark.parametrize("d", [7, 9, 15, 20, 100])
def test_exponential_any_dimension_matches_oracle(km, orc, d):
    """d without a compiled fused kernel (7, 9, 11, 13-15, > 16) runs the batched half-step with a thread-per-point sum in
    the oracle's order: eval and a seeded run are bit-identical to the oracle."""
    ld, od = km.exponential(d), orc.Density("exponential", d, [])
    assert ld.info("batched") == 3.0
    rng = np.random.default_rng(d)
    pts = rng.standard_normal((500, d)) + 1.0
    assert np.array_equal(ld.eval(pts), od.eval(pts))
    nw = 2 * (d + 3)
    x0 = np.abs(1.0 + 0.1 * rng.standard_normal((nw, d)))
    want = orc.emcee(od, x0, 30, 10, 3, 2.0, seed=5, nthreads=2)
    for launch_mode in (0, 1):
        r = _run_gpu(km, ld, x0, 30, 10, 3, 2.0, seed=5, launch_mode=launch_mode)
        assert np.array_equal(r["chain_x"], want["chain_x"]) and np.array_equal(r["chain_lp"], want["chain_lp"])
        assert np.array_equal(r["accept_ratio"], want["accept_ratio"])


def test_very_long_chain_of_few_walkers_copies_out(km, orc):
    """Few walkers, > 2.09 M stored samples each (gridDim.y of one transpose launch would overflow): results come back
    and equal the oracle's at both ends of the chain."""
    ld, od = km.exponential(), orc.Density("exponential", 1, [])
    x0 = np.array([[0.5], [0.7], [0.9], [1.1]])
    nitw = 65535 * 32 + 1000
    s = km.Sampler(ld, x0, nitw, 0, 1, 2.0, seed=3)
    s.run(-1)
    th, lp, ar = s.results()
    s.close()
    assert th.shape == (4, nitw, 1) and np.all(np.isfinite(lp)) and np.all(th >= 0)
    want = orc.emcee(od, x0, 2000, 0, 1, 2.0, seed=3, nthreads=1)
    assert np.array_equal(th[:, :2000], want["chain_x"]) and np.array_equal(lp[:, :2000], want["chain_lp"])
    assert np.array_equal(lp[:, -1], -th[:, -1, 0]) and 0.3 < ar.mean() < 0.95
