"""Small configurations of the kernels changed in round 2, for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python profiles/sanitizer_run.py [k1|k2g|push|huge|all]"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import kissmcmc_b200 as km  # noqa: E402
from tests import cases  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"


def run(ld, x0, nitw, **kw):
    s = km.Sampler(ld, x0, nitw, nitw // 2, 2, 2.0, 7, **kw)
    s.run(-1)
    th, lp, ar = s.results()
    s.close()
    return float(ar.mean())


if what in ("k1", "all"):        # emcee_smem_kernel: 16384 walkers, several CTAs, grid barrier with the polling warp
    name, d, params, th0, rad = cases.plugin_specs()["rosenbrock"]
    print("k1", run(km.LogDensity(name, d, params), cases.ball(th0, rad, 16384, 1), 20))
if what in ("k2g", "all"):       # gaussian_fused2_kernel, tensor cores on
    d = 100
    ld = km.gaussian(np.linspace(-1, 1, d), cases.spd_cov(d, 2))
    ld.set_option("tensor_cores", 1)
    print("k2g", run(ld, cases.ball(np.zeros(d), 0.5, 4096, 2), 6))
if what in ("huge", "all"):      # gaussian_huge_logp_kernel through the batched half-step
    d = 200
    ld = km.gaussian(np.zeros(d), cases.spd_cov(d, 3))
    print("huge", run(ld, cases.ball(np.zeros(d), 0.5, 404, 3), 4))
if what in ("push", "all"):      # emcee_push_kernel: two ranks on one GPU
    name, d, params, th0, rad = cases.plugin_specs()["mvn10"]
    x0 = cases.ball(th0, rad, 8192, 4)
    m = km.MultiSampler(km.LogDensity(name, d, params), x0, 8, 4, 2, 2.0, 5, devices=[0, 0], sharded=True)
    m.run(-1)
    print("push", float(m.results()[2].mean()))
    m.close()
