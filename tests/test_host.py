"""CPU tests (no GPU) of the host side: the C-ABI library loads and exports every declared
symbol, argument checks mirror the reference's asserts, make_theta0s / squash_walkers match
the numpy restatements in oracle/oracle.py."""
import re
from pathlib import Path

import numpy as np
import pytest

from tests import cases

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol(km):
    header = (ROOT / "include" / "kissmcmc_cuda.h").read_text()
    declared = set(re.findall(r"\b(kmc_[a-z_0-9]+)\s*\(", header))
    assert declared == set(km.SYMBOLS)
    for name in declared:
        assert hasattr(km.lib, name), name
    assert km.lib.kmc_version() >= 100


def test_density_create_validation(km):
    with pytest.raises(km.KmcError) as e:
        km.LogDensity("no_such_plugin", 1)
    assert e.value.code == 1
    with pytest.raises(km.KmcError) as e:
        km.LogDensity("rosenbrock", 2, [1.0])           # wrong parameter count
    assert e.value.code == 1
    with pytest.raises(km.KmcError) as e:
        km.LogDensity("rosenbrock", 3, [1.0, 100.0, 20.0])
    assert e.value.code == 3
    km.rosenbrock(), km.exponential(), km.gaussian([0, 0], np.eye(2)), km.lognormal()   # creation needs no GPU


def test_emcee_argument_checks(km):
    """src/samplers.jl:200-205 asserts fire before any device work."""
    ld = km.rosenbrock()
    x = cases.ball([0, 0], 0.1, 10, 0)
    with pytest.raises(AssertionError):
        km.emcee(ld, x, a_scale=1.0, use_progress_meter=False)
    with pytest.raises(AssertionError, match="even number"):
        km.emcee(ld, x[:9], use_progress_meter=False)
    with pytest.raises(AssertionError, match="DOF\\+2"):
        km.emcee(ld, x[:2], use_progress_meter=False)
    with pytest.raises(NotImplementedError):
        km.emcee(ld, x, hasblob=True)
    with pytest.raises(TypeError, match="plugin"):
        km.emcee(lambda t: -np.sum(t ** 2), x)
    with pytest.raises(TypeError, match="plugin"):
        km.make_theta0s(0.5, 0.1, lambda t: -t, 10)


def test_no_cpu_fallback(km):
    """Without a CUDA device the compute entry points fail loudly."""
    try:
        n = km.device_count()
    except km.KmcError:
        n = 0
    if n > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(km.KmcError):
        km.exponential().eval(np.ones((4, 1)))
    with pytest.raises(km.KmcError):
        km.emcee(km.rosenbrock(), cases.ball([0, 0], 0.1, 10, 0), niter=100, use_progress_meter=False)


class _OracleBackedDensity:
    """Test double: a LogDensity whose eval() is the CPU oracle, to exercise host logic without a GPU."""

    def __new__(cls, km, od):
        class _D(km.LogDensity):
            def __init__(self, od):
                self.d, self._od, self._h = od.d, od, None
                self.calls = 0

            def eval(self, thetas):
                self.calls += 1
                return self._od.eval(thetas)
        return _D(od)


@pytest.mark.parametrize("theta0,radius,case", [(0.5, 0.1, "exponential"), ([0.0, 0.0], 0.1, "rosenbrock"),
                                                ([0.4, 0.3], [0.1, 0.2], "mvn2"), (0.02, 0.5, "lognormal"),
                                                (-2.0, 0.3, "exponential")])
def test_make_theta0s_matches_reference_loop(km, orc, theta0, radius, case):
    """Batched make_theta0s == the sequential loop of src/samplers.jl:311-349 on the same draws,
    including rejections (exponential / lognormal have zero density on half the ball) and the
    cumulative radius halving when a walker exhausts its tries (theta0=-2: every k=1 try fails)."""
    name, d, params, *_ = cases.plugin_specs()[case]
    od = orc.Density(name, d, params)
    nw, seed = 40, 123
    fake = _OracleBackedDensity(km, od)
    kw = dict(ntries=5 if theta0 == -2.0 else 100)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = km.make_theta0s(theta0, radius, fake, nw, seed=seed, **kw)
    want = orc.make_theta0s(theta0, radius, od.logpdf, nw,
                            lambda i, k, j: km.ball_randn(seed, [i], k, j, d)[0], **kw)
    assert got.shape == want.shape
    assert np.array_equal(got, want)
    if theta0 != -2.0:
        assert len(got) == nw and fake.calls < 60          # batched, not nw serial calls
        assert np.all(od.eval(np.reshape(got, (nw, -1))) > -np.inf)
    assert got.ndim == (1 if np.ndim(theta0) == 0 else 2)  # scalar theta0 -> [nw]


def test_ball_randn_is_standard_normal(km):
    z = km.ball_randn(7, np.arange(20000), 1, 1, 3)
    assert z.shape == (20000, 3)
    assert abs(z.mean()) < 0.02 and abs(z.std() - 1) < 0.02
    assert np.array_equal(z[5:9], km.ball_randn(7, np.arange(5, 9), 1, 1, 3))      # pure function of the counter
    assert not np.array_equal(z[:4], km.ball_randn(8, np.arange(4), 1, 1, 3))


def test_host_philox_matches_oracle(km, orc):
    rng = np.random.default_rng(0)
    for _ in range(20):
        c = [int(v) for v in rng.integers(0, 2**32, 4)]
        k = [int(v) for v in rng.integers(0, 2**32, 2)]
        got = [int(v) for v in km.philox4x32_10(*c, *k)]
        assert got == orc.philox4x32_10(c, k)


@pytest.mark.parametrize("d", [None, 3])
@pytest.mark.parametrize("kw", [dict(), dict(order=True), dict(drop_low_accept_ratio=True, verbose=False),
                                dict(drop_low_accept_ratio=True, drop_fact=1, order=True, verbose=False)])
def test_squash_walkers_matches_reference(km, orc, d, kw):
    """src/samplers.jl:372-428; walker-major concatenation, optional drop and time ordering."""
    rng = np.random.default_rng(1)
    nw, ns = 12, 7
    thetas = rng.standard_normal((nw, ns) if d is None else (nw, ns, d))
    logp = rng.standard_normal((nw, ns))
    ar = rng.uniform(0.2, 0.6, nw)
    ar[3] = 0.01                                            # a stuck walker
    okw = {k: v for k, v in kw.items() if k != "verbose"}
    t, a, l, b = km.squash_walkers(thetas, ar, logp, **kw)
    t0, a0, l0, b0 = orc.squash_walkers(thetas, ar, logp, **okw)
    assert b is None and b0 is None
    assert np.array_equal(t, t0) and np.array_equal(l, l0) and a == a0
    if kw.get("drop_low_accept_ratio"):
        assert len(t) % ns == 0 and len(t) <= (nw - 1) * ns     # at least the stuck walker is gone
        assert a > np.mean(ar)
    else:
        assert len(t) == nw * ns
    t1, a1, l1, _ = km.squash_walkers(thetas, ar, **kw)     # logdensities optional (:372)
    assert l1 is None and np.array_equal(t1, t)
