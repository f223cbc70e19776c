"""Accuracy and speed of one build of the tcgen05 logistic kernel (KMC_LIB selects the build).
    KMC_LIB=build/variants/libkmc_p4.so python profiles/k3_variants.py
Prints max |logp_tc - logp_fp64| and the same for differences between neighbouring points at N = 10^6 (the
bench workload), then the per-half-step time of the bench workload."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
import kissmcmc_b200 as km  # noqa: E402

wl = dict(bench.WORKLOADS["logistic32d"])
params, x0 = bench.make_inputs(wl, 1)
ld = km.LogDensity(wl["plugin"], wl["d"], params, data=wl.get("_data"))
pts = x0[:512]
ld.set_option("tensor_cores", 1)      # opt-in since round 2
got = ld.eval(pts)
ld.set_option("tensor_cores", 0)
want = ld.eval(pts)
ld.set_option("tensor_cores", 1)
err = got - want
print(f"lib {km.LIB_PATH}: N=1e6, 512 points: max|err| {np.abs(err).max():.3e} mean err {err.mean():+.3e} "
      f"max|err of neighbour differences| {np.abs(np.diff(err)).max():.3e} rms {np.diff(err).std():.3e}")
for rep in range(3):
    s = km.Sampler(ld, x0, 4, 2, 1, 2.0, seed=rep)
    s.run(-1)
    ms, n = s.last_run_ms()
    print(f"rep {rep}: {ms * 1e3 / 8:.1f} us per half-step, {wl['nw'] * 4 / (ms * 1e-3):.3e} walker-steps/s")
    s.close()
