"""Small driver for ncu captures: runs a few iterations of a bench workload.
    python profiles/prof_run.py [workload] [iters] [launch_mode] [nwalkers]"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
import kissmcmc_b200 as km  # noqa: E402

wl = dict(bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "rosenbrock2d"])
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 100
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0
if len(sys.argv) > 4:
    wl = dict(wl, nw=int(sys.argv[4]))
params, x0 = bench.make_inputs(wl, 1)
ld = km.LogDensity(wl["plugin"], wl["d"], params, data=wl.get("_data"))
import os
if os.environ.get("KMC_TC") is not None:
    ld.set_option("tensor_cores", int(os.environ["KMC_TC"]))
if os.environ.get("KMC_FUSED_VARIANT") is not None:
    ld.set_option("fused_variant", int(os.environ["KMC_FUSED_VARIANT"]))
print("tensor_cores:", ld.info("tensor_cores"))
for rep in range(3):
    s = km.Sampler(ld, x0, iters, iters // 2, max(1, iters // 4), 2.0, seed=rep, launch_mode=mode)
    s.run(-1)
    ms, n = s.last_run_ms()
    print(f"rep {rep}: {iters} iterations, {n} launches, {ms:.3f} ms, {ms * 1e3 / (2 * iters):.2f} us per half-step, "
          f"{wl['nw'] * iters / (ms * 1e-3):.3e} walker-steps/s")
    s.close()
