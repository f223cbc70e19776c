"""kissmcmc.jl_b200 -- B200-native emcee stretch-move hot path of KissMCMC.jl.

Holds only what the path needs: csrc/ (CUDA kernels + the C-ABI of libkissmcmc_cuda.so), the
Python host mirror of the reference's public API (api.py) and the Julia `CUDABackend` module
source (julia/).  Import it as `kissmcmc_b200` (the directory name has a dot in it; the
repo-root shim kissmcmc_b200.py registers it).
"""
from ._lib import KmcError, MODE_PHILOX, MODE_REPLAY, SYMBOLS, LIB_PATH, device_count, lib
from .api import (LogDensity, Sampler, ball_randn, emcee, exponential, gaussian, gaussian_params, logistic, lognormal,
                  make_theta0s, philox4x32_10, rosenbrock, squash_walkers)

from .analysis import acor1d, auto_window, eff_samples, evaluate_convergence, int_acorr  # noqa: E402
from . import distributed  # noqa: E402  (multi-GPU drivers; imports torch)

__all__ = [
    "distributed", "int_acorr", "acor1d", "auto_window", "eff_samples", "evaluate_convergence",
    "emcee", "make_theta0s", "squash_walkers", "LogDensity", "Sampler", "exponential", "rosenbrock", "gaussian",
    "gaussian_params", "lognormal", "logistic", "KmcError", "MODE_PHILOX", "MODE_REPLAY", "device_count", "ball_randn",
    "philox4x32_10", "SYMBOLS", "LIB_PATH", "lib",
]
