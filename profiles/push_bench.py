"""Sharded-ensemble throughput with the push exchange, one process driving all GPUs (kmc_emcee_create_multi).
    python profiles/push_bench.py <log2 nwalkers> <iters> <devices, e.g. 0,1,2,3 | 0p> [chunk,cap,lag ...]
One JSON line per configuration (0 = library default; lag > 0: ordered hand-out with that lag, -1: adaptive hand-out).
(The batch / age columns of the round-2 logs were environment knobs of an experiment build; the library now fixes them.)  A single device runs the unsharded sampler (the
1-GPU anchor); "0p" runs the push kernel with ONE rank (its task loop alone).  Device time = max over the devices'
kernels (CUDA events on each launch stream)."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench
import kissmcmc_b200 as km

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 24
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
devarg = sys.argv[3] if len(sys.argv) > 3 else "0"
push1 = devarg.endswith("p")
devices = [int(v) for v in devarg.rstrip("p").split(",")]
import os
cfgs = [tuple(int(v) for v in a.split(",")) for a in sys.argv[4:]] or [(0, 0, 0)]
cfgs = [tuple(list(c) + [0] * (5 - len(c))) for c in cfgs]          # chunk, cap, lag, batch, age (0 = default)
wl = dict(bench.WORKLOADS["gaussian10d"], nw=1 << lg)
params, x0 = bench.make_inputs(wl, 1)
nw, d = wl["nw"], wl["d"]
ld = km.LogDensity("gaussian", d, params, device=devices[0])
warm, G = 4, len(devices)
S = nw // 2 // G
for chunk, cap, lag, nbatch, age in cfgs:
    if G == 1:
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
        m.run(iters)
        ms = m.last_run_ms()
        m.close()
        mode = "sharded-push"
    km.trim()
    nvl = S * 8 * d * (G - 1) // G if len(set(devices)) > 1 else 0
    print(json.dumps({"mode": mode, "devices": devices, "nwalkers": nw, "iters": iters, "chunk": chunk, "cap": cap,
                      "lag": lag, "batch": nbatch, "age": age, "ms_per_halfstep": round(ms / (2 * iters), 5),
                      "walker_steps_per_s": float("%.4g" % (nw * iters / (ms * 1e-3))),
                      "nvlink_gbs_per_gpu": round(nvl / (ms / (2 * iters) * 1e-3) / 1e9, 1)}), flush=True)
