"""Generates tests/golden/*.npz from the CPU oracle (oracle/kmc_oracle.c).

The reference is Julia and cannot run in this image, and its tests hold no golden vectors, so
these fixtures pin OUR restatement: seeded inputs, the full draw trace (partner, z, u), every
accept decision and the resulting chains.  CPU tests check the oracle still reproduces them;
GPU tests check the CUDA path against them without importing the oracle.

    python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import oracle as o  # noqa: E402
from tests import cases  # noqa: E402

OUT = Path(__file__).resolve().parent
CONFIGS = [  # (case, nw, niter_walker, nburnin_walker, nthin, a_scale, seed)
    ("exponential", 100, 40, 20, 1, 2.0, 11),
    ("exponential3", 16, 30, 10, 3, 2.5, 12),
    ("rosenbrock", 64, 60, 30, 2, 2.0, 13),
    ("normal", 10, 50, 0, 1, 2.0, 14),
    ("mvn2", 32, 40, 15, 5, 3.5, 15),
    ("mvn10", 24, 30, 10, 1, 2.0, 16),
    ("lognormal", 20, 40, 20, 1, 2.0, 17),
]


def main():
    specs = cases.plugin_specs()
    for case, nw, nit, nb, nthin, a, seed in CONFIGS:
        name, d, params, th0, rad = specs[case]
        dens = o.Density(name, d, params)
        x0 = np.abs(cases.ball(th0, rad, nw, seed)) if name in ("exponential", "lognormal") else cases.ball(th0, rad, nw, seed)
        r = o.emcee(dens, x0, nit, nb, nthin, a, seed=seed, trace=True)
        np.savez_compressed(
            OUT / f"{case}.npz", name=name, d=d, params=np.asarray(params, dtype=np.float64), theta0s=x0,
            niter_walker=nit, nburnin_walker=nb, nthin=nthin, a_scale=a, seed=seed,
            partner=r["trace"][0], z=r["trace"][1], u=r["trace"][2], accept=r["trace"][3],
            chain_x=r["chain_x"], chain_lp=r["chain_lp"], accept_ratio=r["accept_ratio"], naccept=r["naccept"],
            final_x=r["x"], final_lp=r["lp"], min_margin=r["min_margin"])
        print(case, r["chain_x"].shape, "accept", r["accept_ratio"].mean().round(3), "margin", r["min_margin"])


if __name__ == "__main__":
    main()
