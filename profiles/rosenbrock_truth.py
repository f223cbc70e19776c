"""Free-running statistical check against the reference's own long-run numbers: the Rosenbrock/20 posterior
moments quoted in /root/reference/test/runtests.jl:70-72 ("from running emcee with niter=10^9"):
mean [0.98, 10.3], std [3.1, 13.8].  Runs ~2.6e10 walker-steps on the GPU (26x the reference's run) and
reduces the chain moments on the device (kmc_emcee_chain_moments)."""
import json, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import kissmcmc_b200 as km

nw, nitw, nthin = 1 << 16, 400_000, 200
ld = km.rosenbrock()
x0 = km.make_theta0s(np.array([0.0, 0.0]), 0.1, ld, nw, seed=1)
t0 = time.time()
s = km.Sampler(ld, x0, nitw, nitw // 4, nthin, 2.0, seed=2024)
s.run(-1)
mean, var, n = s.chain_moments()
th, lp, ar = None, None, None
ms, _ = s.last_run_ms()
_, _, na = s.state()
s.close()
out = {"walker_steps": nw * nitw, "samples": n, "mean": mean.tolist(), "std": np.sqrt(var).tolist(),
       "reference_mean": [0.98, 10.3], "reference_std": [3.1, 13.8],
       "accept_ratio_mean": float(na.mean() / (nitw - nitw // 4)), "kernel_ms": ms, "wall_s": time.time() - t0}
print(json.dumps(out))
