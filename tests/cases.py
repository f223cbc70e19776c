"""Shared test cases: the plugin instances used by both the oracle and the CUDA path."""
import math

import numpy as np

MVN_MEAN = [0.5, -0.25]                      # /root/reference/test/runtests.jl:61
MVN_COV = [[0.47, 1.8], [1.8, 7.0]]


def gaussian_params(mean, cov):
    # same construction as kissmcmc_b200.gaussian_params, restated so CPU tests do not depend on it
    mu = np.atleast_1d(np.asarray(mean, dtype=np.float64))
    cov = np.atleast_2d(np.asarray(cov, dtype=np.float64))
    prec = np.linalg.inv(cov)
    L = np.linalg.cholesky((prec + prec.T) / 2)
    lognorm = float(np.sum(np.log(np.diag(L))) - 0.5 * mu.size * math.log(2 * math.pi))
    return np.concatenate([mu, L.T.ravel(), [lognorm]])


def spd_cov(d, seed=0):
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((d, d))
    return a @ a.T / d + np.eye(d)


def plugin_specs():
    """name -> (plugin name, d, params, theta0, ball_radius)."""
    return {
        "exponential": ("exponential", 1, [], 0.5, 0.1),
        "exponential3": ("exponential", 3, [], [0.5, 1.0, 2.0], 0.1),
        "rosenbrock": ("rosenbrock", 2, [1.0, 100.0, 20.0], [0.0, 0.0], 0.1),
        "normal": ("gaussian", 1, gaussian_params(-5.0, 9.0), -4.0, 0.1),
        "mvn2": ("gaussian", 2, gaussian_params(MVN_MEAN, MVN_COV), [0.4, 0.3], 0.1),
        "mvn10": ("gaussian", 10, gaussian_params(np.linspace(-1, 1, 10), spd_cov(10, 1)), np.zeros(10), 0.1),
        "lognormal": ("lognormal", 1, [0.0, 1.0, 0.5 * math.log(2 * math.pi)], 0.4, 0.1),
    }


def ball(theta0, radius, nw, seed):
    rng = np.random.default_rng(seed)
    th0 = np.atleast_1d(np.asarray(theta0, dtype=np.float64))
    x = th0[None, :] + radius * rng.standard_normal((nw, th0.size))
    return x


def logistic_problem(N=20000, d=32, seed=0):
    """Synthetic Bayesian logistic regression (BASELINE.json configs[3], scaled): X ~ N(0,1) rounded to
    bf16-representable values, theta* ~ N(0,1)/sqrt(d), y ~ Bernoulli(sigmoid(X theta*))."""
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((N, d)).astype(np.float32)
    X = (X.view(np.uint32) & np.uint32(0xFFFF0000)).view(np.float32)        # exact in bf16
    tstar = rng.standard_normal(d) / np.sqrt(d)
    y = (rng.random(N) < 1.0 / (1.0 + np.exp(-(X.astype(np.float64) @ tstar)))).astype(np.float32)
    return X, y, tstar
