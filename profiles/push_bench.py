"""Sharded-ensemble throughput with the push exchange, one process driving all GPUs (kmc_emcee_create_multi).
    python profiles/push_bench.py <log2 nwalkers> <iters> <devices, e.g. 0,1,2,3> [chunk] [cap] [lag]
Prints one JSON line per run: the unsharded single-GPU anchor (devices of length 1 uses the plain sampler) or the
sharded run.  Device time = max over the devices' kernels (CUDA events on each launch stream)."""
import json
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np

import bench
import kissmcmc_b200 as km

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 24
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
devarg = sys.argv[3] if len(sys.argv) > 3 else "0"
push1 = devarg.endswith("p")          # "0p": the push kernel with ONE rank (its task loop alone) on device 0
devices = [int(v) for v in devarg.rstrip("p").split(",")]
chunk, cap, lag = (int(sys.argv[i]) if len(sys.argv) > i else 0 for i in (4, 5, 6))
wl = dict(bench.WORKLOADS["gaussian10d"], nw=1 << lg)
params, x0 = bench.make_inputs(wl, 1)
nw, d = wl["nw"], wl["d"]
ld = km.LogDensity("gaussian", d, params, device=devices[0])
warm = 4
if len(devices) == 1:
    kw = dict(shard=(0, nw // 2), exchange=km.EXCHANGE_PUSH, push_chunk=chunk, push_lag=lag) if push1 else {}
    s = km.Sampler(ld, x0, iters + warm, 0, 10**6, 2.0, 7, device=devices[0], **kw)
    s.run(warm)
    s.run(iters)
    ms, _ = s.last_run_ms()
    s.close()
    mode = "push-1-rank" if push1 else "unsharded"
else:
    m = km.MultiSampler(ld, x0, iters + warm, 0, 10**6, 2.0, 7, devices=devices, sharded=True, push_chunk=chunk,
                        push_cap=cap, push_lag=lag)
    m.run(warm)
    t0 = time.perf_counter()
    m.run(iters)
    wall = time.perf_counter() - t0
    ms = m.last_run_ms()
    m.close()
    mode = "sharded-push"
G = len(devices)
S = nw // 2 // G
print(json.dumps({"mode": mode, "devices": devices, "nwalkers": nw, "d": d, "iters": iters, "chunk": chunk, "cap": cap,
                  "lag": lag, "ms_per_halfstep": ms / (2 * iters), "walker_steps_per_s": nw * iters / (ms * 1e-3),
                  "nvlink_bytes_per_gpu_per_halfstep": S * 8 * d * (G - 1) // G if len(set(devices)) > 1 else 0}))
