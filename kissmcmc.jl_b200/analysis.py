"""Integrated autocorrelation time and effective sample size of ensemble chains.

Restated from the (commented-out, pre-1.0 and untested) reference code in
/root/reference/src/analysis.jl:
    int_acorr    :140-167     acor1d  :252-273 (its text is corrupted by a pasted fragment at :256-262)
    auto_window  :281-286     eff_samples :88-95
Nothing in the reference pins these numbers (all of analysis.jl is commented out and has no
test), so parity is "unpinned"; the known answer used here is AR(1): tau = (1+phi)/(1-phi).

Layout: the reference takes `thetas[ntheta, nsamples, nchains]` (Julia column-major); here chains
come from emcee as [nchains(=walkers), nsamples(, ntheta)] -- the same memory order.
"""
from __future__ import annotations

import warnings

import numpy as np


def acor1d(x, norm: bool = True, zero_pad: bool = False) -> np.ndarray:
    """Autocorrelation function of one chain via FFT, first half kept (:252-273).  Like the reference there is NO zero
    padding by default (the correlation is circular, which biases tau for chains that are short against tau);
    zero_pad=True pads to 2n, the linear autocorrelation -- a deliberate extension, not reference behaviour."""
    x = np.asarray(x, dtype=np.float64)
    n = len(x)
    f = np.fft.fft(x - x.mean(), n=2 * n if zero_pad else n)
    acf = np.real(np.fft.ifft(f * np.conj(f)))[:n]
    acf /= 4 * n
    if norm:
        acf /= acf[0]
    return acf[:n // 2]


def auto_window(taus, c) -> int:
    """First (1-based) index i with i >= c*taus[i], else len-1 (:281-286).  Returns the 1-based index."""
    for i, t in enumerate(taus, start=1):
        if i >= c * t:
            return i
    return len(taus) - 1


def _mean_rho_torch(th, zero_pad=False):
    """Chain-averaged normalised autocorrelation for every parameter on the GPU (batched FFT over
    all chains and parameters at once): th [nchains, nsamples, ntheta] torch tensor -> [ntheta, nsamples//2]."""
    import torch
    x = th.to(torch.float64)
    x = x - x.mean(dim=1, keepdim=True)
    n = x.shape[1]
    f = torch.fft.fft(x, n=2 * n if zero_pad else n, dim=1)
    acf = torch.fft.ifft(f * torch.conj(f), dim=1).real[:, :n, :] / (4 * n)
    acf = acf / acf[:, :1, :]
    return acf[:, :x.shape[1] // 2, :].mean(dim=0).T.contiguous()


def int_acorr(thetas, c=5, warn=True, warnat=50, zero_pad=False):
    """Integrated autocorrelation time per parameter, averaged over chains (:140-167).

    thetas: [nchains, nsamples] or [nchains, nsamples, ntheta]; a numpy array (host FFT, chain by
    chain like the reference) or a torch tensor (one batched FFT on its device).
    Returns (tau[ntheta], converged[ntheta]); converged = nsamples / tau should be > ~50.
    If anything is NaN both are set to -1 (the reference's "hack").  zero_pad: see acor1d (default: the reference's
    circular correlation)."""
    assert c > 1
    rho_dev = None
    if type(thetas).__module__.startswith("torch"):      # device chains: one batched FFT on the GPU
        tt = thetas if thetas.ndim == 3 else thetas[:, :, None]
        nchains, nsamples, ntheta = tt.shape
        rho_dev = _mean_rho_torch(tt, zero_pad).cpu().numpy()
    else:
        th = np.asarray(thetas, dtype=np.float64)
        if th.ndim == 2:
            th = th[:, :, None]
        nchains, nsamples, ntheta = th.shape
    out = np.empty(ntheta)
    for n in range(ntheta):
        if rho_dev is not None:
            rho = rho_dev[n]
        else:
            rho = np.zeros(nsamples // 2)
            for cc in range(nchains):
                rho += acor1d(th[cc, :, n], zero_pad=zero_pad)
            rho /= nchains
        taus = 2 * np.cumsum(rho) - 1          # the -1: dfm/emcee issue 267
        window = auto_window(taus, c)
        out[n] = taus[window - 1]
    converged = nsamples / out
    if warn and np.any(converged < warnat):
        warnings.warn("Estimate of integrated autocorrelation likely not accurate!")
    if np.any(np.isnan(out)) or np.any(np.isnan(converged)):
        out = np.full(ntheta, -1.0)
        converged = np.full(ntheta, -1.0)
    return out, converged


def eff_samples(thetas, c=5):
    """(Neff, suggested thinning, mean convergence, Neff per theta, tau per theta, convergence per theta) (:88-95)."""
    th = np.asarray(thetas, dtype=np.float64)
    if th.ndim == 2:
        th = th[:, :, None]
    nchains, nsamples, _ = th.shape
    acorr, converged = int_acorr(th, c=c, warn=False)
    ns = nsamples / acorr * nchains
    return (int(round(ns.mean())), int(round(nsamples * nchains // ns.mean())), float(converged.mean()),
            np.round(ns).astype(int), acorr, converged)


def evaluate_convergence(*ensembles, c=5):
    """R-hat (potential scale reduction, should be < 1.1), total effective sample size and average
    thinning factor from two or more SEPARATE emcee runs (the reference only has the docstring of
    this function, src/analysis.jl:59-77: "needs input from two separate emcee runs", because the
    walkers inside one ensemble are not independent; Gelman et al. 2014, p. 281-287).

    Each argument is one run's chains [nwalkers, nsamples(, ntheta)] -- e.g. the per-rank results of
    distributed.emcee_independent.  Every ensemble is reduced to its walker-mean time series split
    in two halves (split R-hat); effective sample size and thinning come from eff_samples of all
    walkers of all runs."""
    runs = []
    for th in ensembles:
        th = np.asarray(th, dtype=np.float64)
        runs.append(th[:, :, None] if th.ndim == 2 else th)
    assert len(runs) >= 2, "R-hat needs at least two independent ensembles"
    n = min(r.shape[1] for r in runs) // 2
    chains = []
    for r in runs:                       # chain = all samples of one half of one run, [walkers*n, ntheta]
        chains.append(r[:, :n].reshape(-1, r.shape[2]))
        chains.append(r[:, n:2 * n].reshape(-1, r.shape[2]))
    m, L = len(chains), chains[0].shape[0]
    means = np.stack([ch.mean(0) for ch in chains])
    W = np.stack([ch.var(0, ddof=1) for ch in chains]).mean(0)
    B = L * means.var(0, ddof=1)
    rhat = np.sqrt(((L - 1) / L * W + B / L) / W)
    neff, nthin, _, _, _, _ = eff_samples(np.concatenate([r[:, :2 * n] for r in runs], axis=0), c=c)
    return rhat, neff, nthin
