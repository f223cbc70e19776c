"""int_acorr / acor1d / auto_window / eff_samples (reference: commented code in src/analysis.jl).
Known answer: an AR(1) process x_t = phi x_{t-1} + e_t has tau = (1+phi)/(1-phi)."""
import numpy as np
import pytest


def _ar1(phi, nchains, n, seed):
    rng = np.random.default_rng(seed)
    x = np.zeros((nchains, n))
    e = rng.standard_normal((nchains, n))
    x[:, 0] = e[:, 0] / np.sqrt(1 - phi * phi)
    for t in range(1, n):
        x[:, t] = phi * x[:, t - 1] + e[:, t]
    return x


@pytest.mark.parametrize("phi", [0.0, 0.5, 0.9])
def test_int_acorr_ar1(km, phi):
    x = _ar1(phi, 64, 20000, 1)
    tau, conv = km.int_acorr(x, warn=False)
    want = (1 + phi) / (1 - phi)
    assert tau.shape == (1,) and abs(tau[0] - want) < 0.08 * want + 0.05
    assert conv[0] == pytest.approx(20000 / tau[0])


def test_acor1d_and_window(km):
    x = _ar1(0.7, 1, 4096, 2)[0]
    rho = km.acor1d(x)
    assert len(rho) == 2048 and rho[0] == pytest.approx(1.0) and abs(rho[1] - 0.7) < 0.05
    un = km.acor1d(x, norm=False)
    assert un[0] == pytest.approx(np.sum((x - x.mean()) ** 2) / (4 * len(x)))     # the reference's /(4 n) scaling
    assert km.auto_window(np.array([3.0, 3.0, 3.0, 0.5, 0.1]), 1.5) == 4          # first 1-based i >= c*tau_i
    assert km.auto_window(np.array([9.0, 9.0, 9.0]), 5) == 2                      # none -> len-1


def test_int_acorr_multi_theta_nan_and_warning(km):
    x = np.stack([_ar1(0.5, 8, 5000, 3), _ar1(0.8, 8, 5000, 4)], axis=-1)
    tau, conv = km.int_acorr(x, warn=False)
    assert tau.shape == (2,) and abs(tau[0] - 3) < 0.5 and abs(tau[1] - 9) < 1.5
    neff, thin, mconv, neffs, taus, convs = km.eff_samples(x)
    assert np.array_equal(taus, tau) and neff == int(round((5000 / tau * 8).mean()))
    bad = x.copy()
    bad[0, 0, 0] = np.nan
    t2, c2 = km.int_acorr(bad, warn=False)
    assert np.all(t2 == -1) and np.all(c2 == -1)
    with pytest.warns(UserWarning):
        km.int_acorr(_ar1(0.99, 4, 400, 5))        # far too short: nsamples/tau < 50


@pytest.mark.gpu
def test_int_acorr_of_emcee_chains_and_moment_errors(km):
    """Free-running statistical check of north_star: posterior moments within Monte Carlo error,
    with the error bar taken from the integrated autocorrelation time of the chains themselves."""
    mean, cov = [0.5, -0.25], [[0.47, 1.8], [1.8, 7.0]]
    ld = km.gaussian(mean, cov)
    nw = 256
    x0 = km.make_theta0s(np.array([0.4, 0.3]), 0.1, ld, nw, seed=1)
    th, ar, lp, _ = km.emcee(ld, x0, niter=4000 * nw, nburnin=1000 * nw, use_progress_meter=False, seed=2)
    tau, conv = km.int_acorr(th, warn=False)
    assert np.all(tau > 1) and np.all(tau < 60) and np.all(conv > 50)
    neff = th.shape[0] * th.shape[1] / tau
    sd = np.sqrt(np.diag(cov))
    err = sd / np.sqrt(neff)
    assert np.all(np.abs(th.reshape(-1, 2).mean(0) - mean) < 6 * err + 1e-3)


def test_evaluate_convergence_rhat(km):
    """R-hat ~ 1 for runs of the same stationary process, >> 1.1 when one run sits elsewhere."""
    a, b = _ar1(0.6, 32, 4000, 7), _ar1(0.6, 32, 4000, 8)
    rhat, neff, nthin = km.evaluate_convergence(a, b)
    assert rhat.shape == (1,) and abs(rhat[0] - 1) < 0.02 and neff > 10000 and 2 <= nthin <= 8
    rhat_bad, _, _ = km.evaluate_convergence(a, b + 3.0)
    assert rhat_bad[0] > 1.5
    with pytest.raises(AssertionError):
        km.evaluate_convergence(a)


@pytest.mark.gpu
def test_rhat_of_two_independent_gpu_ensembles(km):
    ld = km.rosenbrock()
    runs = []
    for seed in (1, 2):
        x0 = km.make_theta0s(np.array([0.0, 0.0]), 0.1, ld, 512, seed=seed)
        th, ar, _, _ = km.emcee(ld, x0, niter=6000 * 512, nburnin=2000 * 512, nthin=4, seed=seed, use_progress_meter=False)
        runs.append(th)
    rhat, neff, nthin = km.evaluate_convergence(*runs)
    assert np.all(rhat < 1.1) and neff > 1000


def test_int_acorr_torch_path_matches_numpy(km):
    import torch
    x = np.stack([_ar1(0.5, 16, 3000, 3), _ar1(0.8, 16, 3000, 4)], axis=-1)
    t0, c0 = km.int_acorr(x, warn=False)
    t1, c1 = km.int_acorr(torch.from_numpy(x), warn=False)          # CPU tensor: same code path as CUDA tensors
    np.testing.assert_allclose(t1, t0, rtol=1e-9)


@pytest.mark.gpu
def test_int_acorr_cuda_tensor_matches_numpy(km):
    """The batched device FFT path on a CUDA tensor (the branch the CPU suite cannot reach)."""
    import torch
    x = np.stack([_ar1(0.5, 16, 3000, 3), _ar1(0.8, 16, 3000, 4)], axis=-1)
    t0, c0 = km.int_acorr(x, warn=False)
    t2, c2 = km.int_acorr(torch.from_numpy(x).cuda(), warn=False)
    np.testing.assert_allclose(t2, t0, rtol=1e-9)
    np.testing.assert_allclose(c2, c0, rtol=1e-9)


def test_zero_padded_autocorrelation_is_the_linear_one(km):
    """zero_pad=True (an extension; the reference's correlation is circular): equals the direct lag sums."""
    x = _ar1(0.6, 1, 400, 9)[0]
    xc = x - x.mean()
    direct = np.array([np.dot(xc[:len(x) - k], xc[k:]) for k in range(len(x) // 2)])
    np.testing.assert_allclose(km.acor1d(x, zero_pad=True), direct / direct[0], rtol=1e-10, atol=1e-12)
    assert not np.allclose(km.acor1d(x), direct / direct[0], atol=1e-6)          # the default stays circular
    t_lin, _ = km.int_acorr(_ar1(0.5, 8, 4000, 1), warn=False, zero_pad=True)
    assert abs(t_lin[0] - 3.0) < 0.5                                               # AR(1): tau = (1 + phi) / (1 - phi)
