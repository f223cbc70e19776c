"""K1 (`emcee_smem_kernel`, 2^20-walker Rosenbrock) variant builds: parity against the oracle + time per half-step.
    python profiles/k1_variants.py build/variants/k1_*.so
Each library runs in its own process (KMC_LIB); one JSON line per variant."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent

if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, str(ROOT))
    import numpy as np
    import kissmcmc_b200 as km
    from oracle import oracle as orc
    from tests import cases
    name, d, params, th0, rad = cases.plugin_specs()["rosenbrock"]
    ok = True
    for nw in (4096, 20000):
        x0 = cases.ball(th0, rad, nw, 3)
        want = orc.emcee(orc.Density(name, d, params), x0, 24, 8, 3, 2.0, seed=5, nthreads=8)
        s = km.Sampler(km.LogDensity(name, d, params), x0, 24, 8, 3, 2.0, 5)
        s.run(7)
        s.run(-1)
        th, lp, ar = s.results()
        s.close()
        ok = ok and np.array_equal(th, want["chain_x"]) and np.array_equal(lp, want["chain_lp"]) and \
            np.array_equal(ar, want["accept_ratio"])
    nw, iters = 1 << 20, 1000
    x0 = cases.ball(th0, rad, nw, 1)
    s = km.Sampler(km.LogDensity(name, d, params), x0, iters + 100, 0, 10**6, 2.0, 7)
    s.run(100)
    best = 1e9
    for _ in range(3):
        s2 = km.Sampler(km.LogDensity(name, d, params), x0, iters + 100, 0, 10**6, 2.0, 7)
        s2.run(100)
        s2.run(iters)
        ms, _ = s2.last_run_ms()
        s2.close()
        best = min(best, ms)
    s.close()
    print(json.dumps({"lib": os.path.basename(os.environ.get("KMC_LIB", "in-tree")), "parity": bool(ok),
                      "us_per_halfstep": round(best * 1e3 / (2 * iters), 4),
                      "walker_steps_per_s": float("%.4g" % (nw * iters / (best * 1e-3)))}), flush=True)
    sys.exit(0)

for lib in sys.argv[1:] or [""]:
    env = dict(os.environ)
    if lib:
        env["KMC_LIB"] = str(Path(lib).resolve())
    r = subprocess.run([sys.executable, __file__, "--child"], env=env, capture_output=True, text=True, timeout=300)
    print(r.stdout.strip() or ("FAILED " + lib + "\n" + r.stderr[-1500:]), flush=True)
