"""K1b (emcee_bulk_kernel, HBM-resident 2^24-walker 10-D Gaussian): does the L2 fetch granularity change the cost of
the uniformly random 80-byte partner-row gathers?  cudaLimitMaxL2FetchGranularity in {32, 64, 128} bytes, set through
the CUDA runtime before the library touches the device, then 3 timed launches each.
    python profiles/l2_fetch_experiment.py            (1 GPU; prints one JSON line per setting)"""
import ctypes
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

if len(sys.argv) > 1:            # child: one granularity per process (the limit must be set before the context's first use)
    gran = int(sys.argv[1])
    import torch
    rt = torch.cuda.cudart()
    torch.cuda.init()
    cudaLimitMaxL2FetchGranularity = 0x05
    before = rt.cudaDeviceGetLimit(cudaLimitMaxL2FetchGranularity) if hasattr(rt, "cudaDeviceGetLimit") else None
    lib = ctypes.CDLL("libcudart.so.12")
    rc = lib.cudaDeviceSetLimit(ctypes.c_int(cudaLimitMaxL2FetchGranularity), ctypes.c_size_t(gran))
    val = ctypes.c_size_t(0)
    lib.cudaDeviceGetLimit(ctypes.byref(val), ctypes.c_int(cudaLimitMaxL2FetchGranularity))
    import bench
    import kissmcmc_b200 as km
    wl = dict(bench.WORKLOADS["gaussian10d"])
    params, x0 = bench.make_inputs(wl, 1)
    ld = km.LogDensity("gaussian", 10, params)
    ms = []
    for rep in range(3):
        s = km.Sampler(ld, x0, 24, 0, 10**6, 2.0, rep)
        s.run(4)
        s.run(20)
        ms.append(s.last_run_ms()[0] / 40)
        s.close()
    print(json.dumps({"l2_fetch_granularity_requested": gran, "set_rc": rc, "in_effect": val.value,
                      "ms_per_halfstep": [round(m, 4) for m in ms]}))
else:
    for gran in (32, 64, 128):
        r = subprocess.run([sys.executable, __file__, str(gran)], capture_output=True, text=True)
        print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else "failed: " + r.stderr[-300:])
