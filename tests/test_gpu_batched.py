"""GPU parity tests (`-m gpu`) of the batched plugins -- dense Gaussian with d up to 128 (BASELINE.json
configs[2]) and Bayesian logistic regression (configs[3]) -- through the C-ABI against the CPU oracle.

These log-densities are long dot products: the CUDA kernels accumulate with FMA in a different
(parallel) order than the oracle's sequential loop, so log-densities agree to a stated relative
tolerance (1e-12 Gaussian, 1e-10 logistic) instead of bit-for-bit.  Decisions are compared in
REPLAY mode, where both sides see the same draws: the oracle's smallest decision margin
|lhs - log u| is asserted to be far above the log-density tolerance, so every accept/reject must
agree exactly; states x are produced by exact IEEE sub/mul/add and are compared bit-for-bit."""
import numpy as np
import pytest

from tests import cases

pytestmark = pytest.mark.gpu


def _gauss(km, orc, d, seed=3):
    prm = cases.gaussian_params(np.linspace(-1, 1, d), cases.spd_cov(d, seed))
    return km.LogDensity("gaussian", d, prm), orc.Density("gaussian", d, prm)


@pytest.mark.parametrize("d", [17, 32, 100, 128])
def test_wide_gaussian_eval(km, orc, d):
    ld, od = _gauss(km, orc, d)
    pts = np.random.default_rng(0).standard_normal((777, d)) * 2
    np.testing.assert_allclose(ld.eval(pts), od.eval(pts), rtol=1e-12, atol=0)


@pytest.mark.parametrize("d,npts", [(129, 777), (200, 333), (512, 130), (1000, 9)])
def test_gaussian_beyond_128_dimensions(km, orc, d, npts):
    """128 < d <= 4096: the FP64 kernel with the matrix in L2 (the reference takes any d: a closure, src/samplers.jl:257).
    Same accumulation as the d <= 128 kernel -> the same stated tolerance against the oracle; no tensor-core path."""
    ld, od = _gauss(km, orc, d)
    pts = np.random.default_rng(1).standard_normal((npts, d)) * 2
    np.testing.assert_allclose(ld.eval(pts), od.eval(pts), rtol=1e-12, atol=0)
    assert ld.info("tensor_cores_available") == 0
    with pytest.raises(km.KmcError):
        ld.set_option("tensor_cores", 1)


def test_gaussian_dimension_limit(km):
    with pytest.raises(km.KmcError) as e:
        km.LogDensity("gaussian", 4097, np.zeros(4097 + 4097 * 4097 + 1))
    assert e.value.code == 3


@pytest.mark.parametrize("d,nw,nitw,nbw,nthin", [(100, 512, 12, 4, 2), (40, 200, 9, 0, 1), (160, 324, 8, 2, 2)])
def test_wide_gaussian_replay_parity(km, orc, d, nw, nitw, nbw, nthin):
    ld, od = _gauss(km, orc, d)
    x0 = cases.ball(np.zeros(d), 0.5, nw, 5)
    want = orc.emcee(od, x0, nitw, nbw, nthin, 2.0, seed=21, trace=True, nthreads=4)
    assert want["min_margin"] > 1e-8            # no decision within reach of a 1e-12 relative log-density error
    for replay in (want["trace"][:3], None):    # replay of the oracle's draws, then the device's own Philox stream
        s = km.Sampler(ld, x0, nitw, nbw, nthin, 2.0, 21, km.MODE_REPLAY if replay is not None else km.MODE_PHILOX)
        if replay is not None:
            s.set_replay(*replay)
        s.run(-1)
        th, lp, ar = s.results()
        x, l, na = s.state()
        s.close()
        assert np.array_equal(ar, want["accept_ratio"]) and np.array_equal(na, want["naccept"])   # decisions exact
        assert np.array_equal(th, want["chain_x"]) and np.array_equal(x, want["x"])               # states bit-exact
        np.testing.assert_allclose(lp, want["chain_lp"], rtol=1e-12)
        np.testing.assert_allclose(l, want["lp"], rtol=1e-12)


def test_wide_gaussian_statistics(km):
    """Free-running: posterior mean / marginal std of a 100-D correlated Gaussian within Monte Carlo error."""
    d, nw = 100, 2048
    mean, cov = np.linspace(-1, 1, d), cases.spd_cov(d, 3)
    ld = km.gaussian(mean, cov)
    x0 = mean + cases.ball(np.zeros(d), 1.0, nw, 1)
    th, ar, lp, _ = km.emcee(ld, x0, niter=600 * nw, nburnin=400 * nw, nthin=20, use_progress_meter=False, seed=2)
    t, a, _, _ = km.squash_walkers(th, ar)
    sd = np.sqrt(np.diag(cov))
    assert a > 0.1
    assert np.all(np.abs(t.mean(0) - mean) < 0.25 * sd)
    assert np.all(np.abs(t.std(0) - sd) < 0.25 * sd)


def test_logistic_eval_and_validation(km, orc):
    X, y, tstar = cases.logistic_problem(N=5000, d=32, seed=1)
    ld = km.logistic(X, y, prior_sigma=10.0)
    od = orc.Density("logistic", 32, [10.0], data=np.concatenate([X.ravel(), y]))
    pts = tstar + 0.05 * np.random.default_rng(2).standard_normal((70, 32))
    assert ld.info("tensor_cores") == 0.0 and ld.info("tensor_cores_available") == 1.0
    np.testing.assert_allclose(ld.eval(pts), od.eval(pts), rtol=1e-10)            # default: exact FP64 kernel
    ld.set_option("tensor_cores", 1)                                              # opt-in: tcgen05 path
    np.testing.assert_allclose(ld.eval(pts), od.eval(pts), rtol=0, atol=2e-3)
    with pytest.raises(km.KmcError):
        km.LogDensity("logistic", 32, [10.0])                       # no data
    with pytest.raises(km.KmcError):
        km.LogDensity("logistic", 32, [-1.0], data=np.concatenate([X.ravel(), y]))


def test_logistic_replay_parity(km, orc):
    d, N, nw, nitw, nbw, nthin = 8, 3000, 64, 10, 4, 3
    X, y, tstar = cases.logistic_problem(N=N, d=d, seed=4)
    ld = km.logistic(X, y, prior_sigma=10.0)
    od = orc.Density("logistic", d, [10.0], data=np.concatenate([X.ravel(), y]))
    x0 = tstar + cases.ball(np.zeros(d), 0.02, nw, 7)
    want = orc.emcee(od, x0, nitw, nbw, nthin, 2.0, seed=9, trace=True, nthreads=4)
    assert want["min_margin"] > 1e-6
    s = km.Sampler(ld, x0, nitw, nbw, nthin, 2.0, 9, km.MODE_REPLAY)
    s.set_replay(*want["trace"][:3])
    s.run(-1)
    th, lp, ar = s.results()
    s.close()
    assert np.array_equal(ar, want["accept_ratio"]) and np.array_equal(th, want["chain_x"])
    np.testing.assert_allclose(lp, want["chain_lp"], rtol=1e-10)


def test_logistic_posterior_recovers_truth(km):
    """Free-running on N=20000, d=8: posterior mean within 4 posterior std of theta*, std ~ 1/sqrt(N) scale."""
    d, N, nw = 8, 20000, 256
    X, y, tstar = cases.logistic_problem(N=N, d=d, seed=5)
    ld = km.logistic(X, y, prior_sigma=10.0)
    x0 = tstar + cases.ball(np.zeros(d), 0.01, nw, 3)
    th, ar, lp, _ = km.emcee(ld, x0, niter=300 * nw, nburnin=150 * nw, nthin=5, use_progress_meter=False, seed=6)
    t, a, _, _ = km.squash_walkers(th, ar)
    assert a > 0.1
    sd = t.std(0)
    assert np.all(sd < 0.05) and np.all(sd > 0.005)
    assert np.all(np.abs(t.mean(0) - tstar) < 5 * sd + 0.02)


# ---------------------------------------------------------------------------------- tcgen05 logistic (K3)

def test_logistic_tensor_core_path_matches_fp64(km):
    """d = 32 with bf16-representable X runs the logits GEMM on tcgen05 (theta split into 3 bf16 pieces,
    FP32 accumulation in TMEM, FP32 softplus).  Stated tolerance against the FP64 CUDA-core kernel of the
    same plugin: |logp_tc - logp_fp64| <= 2e-3 absolute (N = 50k) and the same bound on log-density
    DIFFERENCES between nearby points, which is what accept decisions see.  N is not a multiple of the
    256-row tile (tail masking) and W is not a multiple of 128 (padded rows)."""
    N, d = 50_001, 32
    X, y, tstar = cases.logistic_problem(N=N, d=d, seed=11)
    assert km.logistic(X, y).info("tensor_cores") == 0.0           # approximate path: never on by default
    ld = km.logistic(X, y, prior_sigma=10.0, tensor_cores=True)
    assert ld.info("tensor_cores") == 1.0 and ld.info("batched") == 2.0
    pts = tstar + 0.01 * np.random.default_rng(3).standard_normal((300, d))
    got = ld.eval(pts)
    ld.set_option("tensor_cores", 0)
    assert ld.info("tensor_cores") == 0.0
    want = ld.eval(pts)
    assert np.all(np.isfinite(got))
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-3)
    np.testing.assert_allclose(got[1:] - got[:-1], want[1:] - want[:-1], rtol=0, atol=2e-3)
    # a single point and an exact multiple of the tile sizes
    ld.set_option("tensor_cores", 1)
    np.testing.assert_allclose(ld.eval(pts[:1]), want[:1], rtol=0, atol=2e-3)
    np.testing.assert_allclose(ld.eval(pts[:256]), want[:256], rtol=0, atol=2e-3)


def test_logistic_tensor_core_not_used_when_ineligible(km):
    X, y, _ = cases.logistic_problem(N=2000, d=32, seed=1)
    X2 = X + np.float32(1e-4)                                   # not bf16-representable any more
    ld = km.logistic(X2, y)
    assert ld.info("tensor_cores_available") == 0.0 and ld.info("tensor_cores") == 0.0
    with pytest.raises(km.KmcError) as e:                       # asking for it is an error, never a silent FP64 run
        ld.set_option("tensor_cores", 1)
    assert e.value.code == 3
    with pytest.raises(km.KmcError):
        km.logistic(X2, y, tensor_cores=True)
    X65, y65, _ = cases.logistic_problem(N=500, d=65, seed=1)   # d > 64: FP64 kernel only
    ld65 = km.logistic(X65, y65)
    assert ld65.info("tensor_cores_available") == 0.0
    with pytest.raises(km.KmcError):
        ld65.set_option("tensor_cores", 1)
    X513, y513, _ = cases.logistic_problem(N=50, d=513, seed=1)  # d > 512: no kernel at all
    with pytest.raises(km.KmcError) as e:
        km.logistic(X513, y513)
    assert e.value.code == 3


@pytest.mark.parametrize("d,N", [(65, 3000), (200, 1500), (512, 700)])
def test_logistic_fp64_beyond_64_dimensions(km, orc, d, N):
    """64 < d <= 512 on the exact FP64 kernel (more than 48 KB of shared memory for the 32 points' theta beyond d = 191)."""
    X, y, tstar = cases.logistic_problem(N=N, d=d, seed=4)
    ld = km.logistic(X, y, prior_sigma=3.0)
    od = orc.Density("logistic", d, [3.0], data=np.concatenate([X.ravel(), y]))
    pts = tstar + cases.ball(np.zeros(d), 0.05, 70, 2)
    np.testing.assert_allclose(ld.eval(pts), od.eval(pts), rtol=1e-10)


def test_logistic_tensor_core_sampling(km):
    """Same seeded Philox run on the tcgen05 path and on the FP64 path: the log-density noise of the
    tensor path (<= 2e-3) may flip only decisions that were within that distance of a tie, so almost all
    accept counts agree and the posterior moments agree within Monte Carlo error."""
    N, d, nw = 30_000, 32, 512
    X, y, tstar = cases.logistic_problem(N=N, d=d, seed=12)
    ld = km.logistic(X, y, prior_sigma=10.0, tensor_cores=True)
    x0 = tstar + cases.ball(np.zeros(d), 0.01, nw, 2)
    kw = dict(niter=120 * nw, nburnin=60 * nw, nthin=4, use_progress_meter=False, seed=3)
    th_tc, ar_tc, lp_tc, _ = km.emcee(ld, x0, **kw)
    ld.set_option("tensor_cores", 0)
    th64, ar64, lp64, _ = km.emcee(ld, x0, **kw)
    assert ar_tc.mean() > 0.05 and abs(ar_tc.mean() - ar64.mean()) < 0.02
    t_tc, t64 = th_tc.reshape(-1, d), th64.reshape(-1, d)
    sd = t64.std(0)
    assert np.all(np.abs(t_tc.mean(0) - t64.mean(0)) < 0.5 * sd)
    assert np.all(np.abs(t_tc.std(0) - sd) < 0.3 * sd)
    assert np.all(np.abs(t_tc.mean(0) - tstar) < 6 * sd + 0.03)


@pytest.mark.parametrize("d,N", [(16, 20_003), (48, 20_003), (64, 7_000), (5, 3_001), (33, 4_096)])
def test_logistic_tensor_core_any_d_up_to_64(km, orc, d, N):
    """d is zero-padded to the GEMM's K: 32 (64-byte rows, SWIZZLE_64B) for d <= 32, 64 (128-byte rows, SWIZZLE_128B)
    for 32 < d <= 64.  Against the ORACLE density, stated tolerance 2e-3 on the value and on differences."""
    X, y, tstar = cases.logistic_problem(N=N, d=d, seed=20 + d)
    ld = km.logistic(X, y, prior_sigma=10.0, tensor_cores=True)
    assert ld.info("tensor_cores") == 1.0
    od = orc.Density("logistic", d, [10.0], data=np.concatenate([X.ravel(), y]))
    pts = tstar + 0.02 * np.random.default_rng(d).standard_normal((200, d))
    got, want = ld.eval(pts), od.eval(pts)
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-3)
    np.testing.assert_allclose(got[1:] - got[:-1], want[1:] - want[:-1], rtol=0, atol=2e-3)
    # and a short sampling run on the tensor path stays a valid chain of the same target
    nw = 2 * (d + 3)
    x0 = tstar + cases.ball(np.zeros(d), 0.02, nw, 1)
    th, ar, lp, _ = km.emcee(ld, x0, niter=12 * nw, nburnin=4 * nw, use_progress_meter=False, seed=2)
    np.testing.assert_allclose(lp, od.eval(th.reshape(-1, d)).reshape(lp.shape), rtol=0, atol=2e-3)


# ---------------------------------------------------------------------------------- tcgen05 dense Gaussian (K2)

def _check_tensor_gaussian_replay(od, prm, want, th, lp, x, l, na, min_same=0.98):
    """Replay of the oracle's draws on a tensor-core Gaussian path, every assertion unconditional.
    A walker's whole history agrees with the oracle's iff its stored chain and final state are bit-identical (states
    are produced by the same three IEEE operations; one different decision anywhere upstream changes the bits).
      * those walkers: accept counts equal, stored log-densities within the path's stated tolerance 1e-5 (1 + |y|^2);
      * they must be nearly all walkers (a decision can only differ within the tolerance of a tie);
      * EVERY walker, agreeing or not: each stored log-density is the oracle density of the stored position within
        the tolerance, and the final log-density that of the final position -- a diverged walker is still a valid
        chain of the same target."""
    nw = len(na)
    same = np.all(th.reshape(nw, -1) == want["chain_x"].reshape(nw, -1), axis=1) & np.all(x == want["x"], axis=1)
    assert same.mean() >= min_same, same.mean()
    assert np.array_equal(na[same], want["naccept"][same])
    tol = lambda ref: 1e-5 * (1.0 + 2.0 * (prm[-1] - ref))
    assert np.all(np.abs(lp[same] - want["chain_lp"][same]) <= tol(want["chain_lp"][same]))
    ref = od.eval(th.reshape(-1, th.shape[-1])).reshape(lp.shape)
    assert np.all(np.abs(lp - ref) <= tol(ref)), np.max(np.abs(lp - ref) / tol(ref))
    ref = od.eval(x)
    assert np.all(np.abs(l - ref) <= tol(ref))


@pytest.mark.parametrize("d,npts", [(100, 1000), (128, 129), (17, 5), (64, 4096)])
def test_gaussian_tensor_core_path_matches_fp64(km, d, npts):
    """Opt-in tcgen05 Mahalanobis GEMM (both operands split into 3 bf16 pieces, 6 piece pairs, FP32
    accumulation in TMEM).  Stated tolerance against the FP64 kernel: 1e-5 * (1 + |y|^2) absolute."""
    prm = cases.gaussian_params(np.linspace(-1, 1, d), cases.spd_cov(d, 3))
    ld = km.LogDensity("gaussian", d, prm)
    assert ld.info("tensor_cores") == 0.0            # exact FP64 by default
    pts = np.linspace(-1, 1, d) + np.random.default_rng(0).standard_normal((npts, d)) * 1.5
    want = ld.eval(pts)
    ld.set_option("tensor_cores", 1)
    assert ld.info("tensor_cores") == 1.0
    got = ld.eval(pts)
    ss = 2.0 * (prm[-1] - want)
    assert np.all(np.abs(got - want) <= 1e-5 * (1.0 + ss)), np.max(np.abs(got - want) / (1.0 + ss))


def test_gaussian_tensor_core_sampling(km):
    d, nw = 100, 2048
    mean, cov = np.linspace(-1, 1, d), cases.spd_cov(d, 3)
    ld = km.gaussian(mean, cov)
    ld.set_option("tensor_cores", 1)
    x0 = mean + cases.ball(np.zeros(d), 1.0, nw, 1)
    th, ar, lp, _ = km.emcee(ld, x0, niter=600 * nw, nburnin=400 * nw, nthin=20, use_progress_meter=False, seed=2)
    t, a, _, _ = km.squash_walkers(th, ar)
    sd = np.sqrt(np.diag(cov))
    assert a > 0.1
    assert np.all(np.abs(t.mean(0) - mean) < 0.25 * sd) and np.all(np.abs(t.std(0) - sd) < 0.25 * sd)


def test_gaussian_tensor_core_pipeline_states_exact(km, orc):
    """The Y-free tcgen05 pipeline (propose emits bf16 pieces, accept recomputes y for accepted walkers):
    in replay mode every decision whose margin is above the tensor path's log-density error agrees with
    the oracle, and then the states -- recomputed with the same three IEEE operations -- are bit-identical."""
    d, nw, nitw, nbw, nthin = 40, 200, 9, 0, 1
    prm = cases.gaussian_params(np.linspace(-1, 1, d), cases.spd_cov(d, 3))
    ld, od = km.LogDensity("gaussian", d, prm), orc.Density("gaussian", d, prm)
    ld.set_option("tensor_cores", 1)
    x0 = cases.ball(np.zeros(d), 0.5, nw, 5)
    want = orc.emcee(od, x0, nitw, nbw, nthin, 2.0, seed=21, trace=True, nthreads=4)
    s = km.Sampler(ld, x0, nitw, nbw, nthin, 2.0, 21, km.MODE_REPLAY)
    s.set_replay(*want["trace"][:3])
    s.run(-1)
    th, lp, ar = s.results()
    x, l, na = s.state()
    s.close()
    _check_tensor_gaussian_replay(od, prm, want, th, lp, x, l, na)


def test_gaussian_fused_kernel_equals_three_kernel_pipeline(km):
    """K2F (one persistent fused tcgen05 kernel, launch_mode 0) == propose / GEMM / accept kernels
    (launch_mode 1): same bf16 pieces, same MMA order, same epilogue -> bit-identical chains.
    Sizes exercise partial tiles (nw/2 not a multiple of 128), several tiles per CTA and odd d."""
    for d, nw in ((100, 2 * 1000), (33, 2 * 130), (64, 2 * 128 * 150)):
        prm = cases.gaussian_params(np.linspace(-1, 1, d), cases.spd_cov(d, 3))
        ld = km.LogDensity("gaussian", d, prm)
        ld.set_option("tensor_cores", 1)
        ld.set_option("fused_variant", 1)          # K2F; the default K2G sums |y|^2 in a different order
        x0 = np.linspace(-1, 1, d) + cases.ball(np.zeros(d), 0.7, nw, 5)
        out = []
        for mode in (0, 1):
            s = km.Sampler(ld, x0, 8, 3, 2, 2.0, 77, launch_mode=mode)
            s.run(3)
            s.run(-1)
            th, lp, ar = s.results()
            x, l, na = s.state()
            s.close()
            out.append((th, lp, ar, x, l, na))
        for a, b in zip(*out):
            assert np.array_equal(a, b)
        assert 0.02 < out[0][2].mean() < 0.9


def test_gaussian_fused_tmem_variant(km, orc):
    """K2G (fused_variant 2: matrix pieces resident in TMEM as the A operand, walker pieces double-buffered, the GEMM
    of a tile hidden behind the next tile's proposal phase) against K2F on the same Philox draws.  The two sum |y|^2 in
    a different order (FP32 lane tree vs column-sequential), so log-densities agree to the tensor path's stated
    tolerance 1e-5 (1 + |y|^2) and a decision can differ only within that distance of a tie: nearly every walker ends
    bit-identical.  Sizes: partial tiles, one / two / three tiles per CTA, odd d.  Replay against the oracle as well."""
    for d, nw in ((100, 2 * 1000), (33, 2 * 130), (64, 2 * 128 * 150), (100, 2 * (148 * 128 * 2 + 777))):
        prm = cases.gaussian_params(np.linspace(-1, 1, d), cases.spd_cov(d, 3))
        ld = km.LogDensity("gaussian", d, prm)
        ld.set_option("tensor_cores", 1)
        x0 = np.linspace(-1, 1, d) + cases.ball(np.zeros(d), 0.7, nw, 5)
        out = []
        for variant in (1, 2):
            ld.set_option("fused_variant", variant)
            assert ld.info("fused_variant") == float(variant)
            s = km.Sampler(ld, x0, 6, 2, 2, 2.0, 77)
            s.run(2)
            s.run(-1)
            th, lp, ar = s.results()
            x, l, na = s.state()
            s.close()
            out.append((th, lp, ar, x, l, na))
        (th1, lp1, ar1, x1, l1, na1), (th2, lp2, ar2, x2, l2, na2) = out
        same = np.all(x1 == x2, axis=1)
        assert same.mean() > 0.995, (d, nw, same.mean())
        ss = 2.0 * (prm[-1] - l1[same])
        assert np.all(np.abs(l2[same] - l1[same]) <= 1e-5 * (1.0 + ss))
        assert np.all(np.isfinite(l2)) and abs(ar2.mean() - ar1.mean()) < 0.01
        first = np.all(th1[:, 0] == th2[:, 0], axis=1)     # the first stored sample (after 3 iterations)
        assert first.mean() > 0.997
    # replay mode against the oracle's exact FP64 run
    d, nw, nitw = 40, 200, 9
    prm = cases.gaussian_params(np.linspace(-1, 1, d), cases.spd_cov(d, 3))
    ld, od = km.LogDensity("gaussian", d, prm), orc.Density("gaussian", d, prm)
    ld.set_option("tensor_cores", 1)
    ld.set_option("fused_variant", 2)
    x0 = cases.ball(np.zeros(d), 0.5, nw, 5)
    want = orc.emcee(od, x0, nitw, 0, 1, 2.0, seed=21, trace=True, nthreads=4)
    s = km.Sampler(ld, x0, nitw, 0, 1, 2.0, 21, km.MODE_REPLAY)
    s.set_replay(*want["trace"][:3])
    s.run(-1)
    th, lp, ar = s.results()
    x, l, na = s.state()
    s.close()
    _check_tensor_gaussian_replay(od, prm, want, th, lp, x, l, na)


# ---------------------------------------------------------------------------------- BASELINE.json sizes vs the oracle

def test_gaussian100d_full_size_against_oracle(km, orc):
    """BASELINE.json configs[2] at FULL size -- 100-D dense Gaussian, 2^16 walkers, the default fused tcgen05 kernel
    (K2G) -- against the oracle: (a) the oracle density of a 1/997 sub-sample of every stored chain entry equals the
    stored log-density within the stated tolerance 1e-5 (1 + |y|^2); (b) accept counts equal the number of state changes
    in the (unthinned) chain; (c) the same sampler at 4096 walkers replays the oracle's draws (decision by decision)."""
    d, nw, nitw = 100, 1 << 16, 8
    prm = cases.gaussian_params(np.linspace(-1, 1, d), cases.spd_cov(d, 1))     # bench.py's configs[2] density
    ld, od = km.LogDensity("gaussian", d, prm), orc.Density("gaussian", d, prm)
    ld.set_option("tensor_cores", 1)
    assert ld.info("fused_variant") == 2.0
    x0 = 0.1 * np.random.default_rng(1000).standard_normal((nw, d))
    s = km.Sampler(ld, x0, nitw, 0, 1, 2.0, seed=5)
    s.run(-1)
    th, lp, ar = s.results()
    x, l, na = s.state()
    s.close()
    flat_th, flat_lp = th.reshape(-1, d), lp.reshape(-1)
    idx = np.arange(0, flat_lp.size, 997)
    ref = od.eval(flat_th[idx])
    tol = 1e-5 * (1.0 + 2.0 * (prm[-1] - ref))
    assert np.all(np.abs(flat_lp[idx] - ref) <= tol), np.max(np.abs(flat_lp[idx] - ref) / tol)
    full = np.concatenate([x0[:, None, :], th], axis=1)
    changes = np.any(full[:, 1:] != full[:, :-1], axis=2).sum(axis=1)
    assert np.array_equal(changes, na) and np.array_equal(ar, na / nitw)
    assert np.array_equal(th[:, -1], x) and np.array_equal(lp[:, -1], l)
    assert 0.01 < ar.mean() < 0.9
    # (c) replay of the oracle at 4096 walkers, same density, same kernel
    nw2, nit2 = 4096, 6
    x02 = 0.1 * np.random.default_rng(7).standard_normal((nw2, d))
    want = orc.emcee(od, x02, nit2, 1, 1, 2.0, seed=33, trace=True, nthreads=8)
    s = km.Sampler(ld, x02, nit2, 1, 1, 2.0, 33, km.MODE_REPLAY)
    s.set_replay(*want["trace"][:3])
    s.run(-1)
    th, lp, ar = s.results()
    x, l, na = s.state()
    s.close()
    _check_tensor_gaussian_replay(od, prm, want, th, lp, x, l, na, min_same=0.97)


def test_logistic32d_full_size_against_oracle(km, orc):
    """BASELINE.json configs[3] at FULL data size -- d = 32, N = 10^6, the tcgen05 kernel (K3) -- against the ORACLE's
    FP64 density on 64 points of a realistic ensemble (ball of radius 1e-3 about theta*).  Stated tolerance: differences
    between neighbouring points (what accept decisions see) within 2e-3; the absolute value within 0.1 (a common
    offset of ~3e-8 per data row from the truncating FP32 accumulation, identical for every point)."""
    N, d = 1_000_000, 32
    X, y, tstar = cases.logistic_problem(N=N, d=d, seed=1001)
    data = np.concatenate([X.ravel(), y])
    ld = km.logistic(X, y, prior_sigma=10.0, tensor_cores=True)
    od = orc.Density("logistic", d, [10.0], data=data)
    pts = tstar + 1e-3 * np.random.default_rng(5).standard_normal((64, d))
    got, want = ld.eval(pts), od.eval(pts)
    assert np.all(np.isfinite(got))
    dd = (got[1:] - got[:-1]) - (want[1:] - want[:-1])
    assert np.max(np.abs(dd)) <= 2e-3, np.max(np.abs(dd))
    assert np.sqrt(np.mean(dd ** 2)) <= 1e-3
    assert np.max(np.abs(got - want)) <= 0.1, np.max(np.abs(got - want))
    off = got - want
    assert off.max() - off.min() <= 2e-3           # the offset is common to all points
    ld.set_option("tensor_cores", 0)               # and the exact FP64 kernel of the same plugin at this size
    np.testing.assert_allclose(ld.eval(pts[:8]), want[:8], rtol=1e-10)


def test_logistic_tensor_core_replay_against_oracle(km, orc):
    """The tcgen05 logistic path in REPLAY mode against the oracle's exact run (d = 32, N = 5000, 256 walkers): walkers
    whose whole history agrees are bit-identical in state; they are nearly all; and EVERY stored log-density is the
    oracle density of the stored position within 2e-3."""
    d, N, nw, nitw, nbw, nthin = 32, 5000, 256, 10, 2, 2
    X, y, tstar = cases.logistic_problem(N=N, d=d, seed=4)
    ld = km.logistic(X, y, prior_sigma=10.0, tensor_cores=True)
    od = orc.Density("logistic", d, [10.0], data=np.concatenate([X.ravel(), y]))
    x0 = tstar + cases.ball(np.zeros(d), 0.05, nw, 7)
    want = orc.emcee(od, x0, nitw, nbw, nthin, 2.0, seed=9, trace=True, nthreads=8)
    s = km.Sampler(ld, x0, nitw, nbw, nthin, 2.0, 9, km.MODE_REPLAY)
    s.set_replay(*want["trace"][:3])
    s.run(-1)
    th, lp, ar = s.results()
    x, l, na = s.state()
    s.close()
    same = np.all(th.reshape(nw, -1) == want["chain_x"].reshape(nw, -1), axis=1) & np.all(x == want["x"], axis=1)
    assert same.mean() >= 0.95, same.mean()
    assert np.array_equal(na[same], want["naccept"][same])
    np.testing.assert_allclose(lp[same], want["chain_lp"][same], rtol=0, atol=2e-3)
    ref = od.eval(th.reshape(-1, d)).reshape(lp.shape)
    np.testing.assert_allclose(lp, ref, rtol=0, atol=2e-3)
    np.testing.assert_allclose(l, od.eval(x), rtol=0, atol=2e-3)
