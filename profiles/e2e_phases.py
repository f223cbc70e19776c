"""Times the phases of one end-to-end emcee job (create / run / results / close) with host buffers."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import bench, kissmcmc_b200 as km

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "rosenbrock2d"]
d, nw, nitw, nthin = wl["d"], wl["nw"], wl["niter_walker"], wl["nthin"]
nbw = nitw // 2; ns = (nitw - nbw) // nthin
params, x0 = bench.make_inputs(wl, 1)
ld = km.LogDensity(wl["plugin"], d, params)
pin = len(sys.argv) > 2 and sys.argv[2] == "pin"
x0p = torch.from_numpy(x0).pin_memory() if pin else torch.from_numpy(x0)
th = torch.empty((nw, ns, d), dtype=torch.float64); lp = torch.empty((nw, ns), dtype=torch.float64); ar = torch.empty(nw, dtype=torch.float64)
if pin: th, lp, ar = th.pin_memory(), lp.pin_memory(), ar.pin_memory()
for rep in range(3):
    t0 = time.perf_counter(); s = km.Sampler(ld, x0p.numpy(), nitw, nbw, nthin, 2.0, seed=rep)
    t1 = time.perf_counter(); s.run(-1)
    t2 = time.perf_counter(); s.results(th.numpy(), lp.numpy(), ar.numpy())
    t3 = time.perf_counter(); s.close()
    t4 = time.perf_counter()
    print(f"rep {rep} pin={pin}: create {1e3*(t1-t0):.1f} ms, run {1e3*(t2-t1):.1f} ms, results {1e3*(t3-t2):.1f} ms, close {1e3*(t4-t3):.1f} ms")
