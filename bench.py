#!/usr/bin/env python
"""bench.py -- walker-steps/s of the emcee stretch-move hot path on N B200s.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload rosenbrock2d]

One "step" is one complete emcee job of the workload: BASELINE.json configs[1], the 2-D
Rosenbrock density with 2^20 walkers and 10^4 iterations per walker (API niter = 10^4 * 2^20,
default burn-in niter/2, nthin=1000 so that 5 samples per walker are stored -- the unthinned
chain would be 126 GB).  At N>1 every rank runs its own ensemble of that size with its own
Philox key and walker-id range (independent-ensembles mode: no data-path collective, weak
scaling); value = walker-steps of all ranks / max-over-ranks device time.

value     device-timed (CUDA events on the launch stream), inputs already resident in HBM.
e2e       the same metric through the public API with HOST buffers: pinned theta0s -> H2D ->
          initial log-densities -> run -> chain transpose -> D2H into pinned result buffers.
roofline  of the dominant kernel (named in roofline.kernel): algorithmic bytes per launch
          (24d+24 per walker-step + 8d+8 per stored sample) over the
          launch duration measured with the library's own CUDA events on the launch stream.
cpu_baseline / --impl reference: the Julia reference cannot run here (no Julia in the image);
          the C restatement of its loop (oracle/kmc_oracle.c, OpenMP over the active half like
          Threads.@threads at src/samplers.jl:248) is timed on the host cores on a bounded
          sample (same ensemble, fewer iterations).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "walker_steps_per_s"
UNIT = "walker-steps/s"

WORKLOADS = {
    # name: plugin, d, nwalkers, niter_walker, nthin
    "rosenbrock2d": dict(plugin="rosenbrock", d=2, nw=1 << 20, niter_walker=10_000, nthin=1000,
                         desc="BASELINE.json configs[1]: 2-D Rosenbrock/20, 2^20 walkers, 10^4 iterations per walker"),
    "gaussian10d": dict(plugin="gaussian", d=10, nw=1 << 24, niter_walker=200, nthin=100,
                        desc="BASELINE.json configs[4] ensemble: 10-D Gaussian, 2^24 walkers (per GPU), 200 iterations"),
    "gaussian100d": dict(plugin="gaussian", d=100, nw=1 << 16, niter_walker=400, nthin=100,
                         desc="BASELINE.json configs[2]: 100-D dense correlated Gaussian, 2^16 walkers, 400 iterations"),
    "logistic32d": dict(plugin="logistic", d=32, nw=8192, niter_walker=4, nthin=1, ndata=1_000_000,
                        desc="BASELINE.json configs[3]: Bayesian logistic regression d=32, N=10^6, 8192 walkers, 4 iterations"),
    "exponential1d": dict(plugin="exponential", d=1, nw=100, niter_walker=1000, nthin=1,
                          desc="BASELINE.json configs[0]: README exponential, 100 walkers, niter=10^5"),
}


def b_step(d: int) -> int:
    """Algorithmic bytes per walker-step."""
    return 24 * d + 24


def dominant_kernel(wl, tensor, launch_mode):
    """Name of the kernel the roofline object describes (selection logic: csrc/kmc_api.cu, kmc_emcee_create/run)."""
    d = wl["d"]
    if d > 16:
        if tensor:
            return "tc::gaussian_fused2_kernel" if launch_mode == 0 else \
                "propose_split_kernel + tc::gaussian_tc_kernel + accept_kernel"
        return "propose_kernel + gaussian_wide_logp_kernel + accept_kernel"
    if launch_mode == 0 and wl["nw"] * (8 * d + 12) <= 148 * 200 * 1024:
        return "emcee_smem_kernel"
    if launch_mode == 0 and d >= 6 and d % 2 == 0:
        return "emcee_bulk_kernel"
    return "emcee_run_kernel"


def make_inputs(wl, seed):
    rng = np.random.default_rng(seed)
    d, nw = wl["d"], wl["nw"]
    if wl["plugin"] == "rosenbrock":
        params = [1.0, 100.0, 20.0]
        x0 = 0.1 * rng.standard_normal((nw, d))
    elif wl["plugin"] == "gaussian":
        from tests import cases
        params = cases.gaussian_params(np.linspace(-1, 1, d), cases.spd_cov(d, 1))
        x0 = 0.1 * rng.standard_normal((nw, d))
    elif wl["plugin"] == "logistic":
        from tests import cases
        X, y, tstar = cases.logistic_problem(N=wl["ndata"], d=d, seed=seed)
        wl["_data"] = np.concatenate([X.ravel(), y])
        params = [10.0]
        x0 = tstar + 1e-3 * rng.standard_normal((nw, d))
    else:
        params = []
        x0 = np.abs(0.5 + 0.1 * rng.standard_normal((nw, d)))
    return params, x0


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.path = gpu_index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.idx)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [v.strip() for v in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def run_reference(args, wl):
    """--impl reference: the C restatement of the reference loop on all host cores, bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle
    nthreads = os.cpu_count() or 1
    # each step ~4 s of CPU work (the whole run stays within a few minutes)
    dens, x0, nw, iters = _cpu_sample(wl, 4.0, nthreads)
    iters, _ = _cpu_run(wl, dens, x0, iters, 4.0, nthreads, 99)      # settles the sample size (untimed)
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        oracle.emcee(dens, x0, iters, iters // 2, max(1, iters // 5), 2.0, seed=i, store=True, nthreads=nthreads,
                     native=True)
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    value = nw * iters * args.steps / total
    sample = (f"{wl['plugin']} d={wl['d']}, {nw} walkers (of {wl['nw']}) x {iters} iterations per step "
              f"(of {wl['niter_walker']})")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "description": wl["desc"], "nwalkers": nw, "d": wl["d"],
                   "note": "C restatement of the reference loop (oracle/kmc_oracle.c, -O3 -march=native -fopenmp, "
                           "-ffp-contract=off); the Julia reference cannot run in this image"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nthreads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def _cpu_sample(wl, seconds, nthreads):
    """Bounded sample of the workload for the CPU leg: (walkers, iterations) worth ~`seconds` of host time.
    Calibrated in two steps (single-thread cost on the smallest legal ensemble, then all threads on a sub-ensemble
    sized from it) so that even the calibration of an expensive density stays short; the sample keeps the full
    ensemble unless two iterations of it would already exceed the budget."""
    from oracle import oracle
    params, x0 = make_inputs(wl, 1)
    dens = oracle.Density(wl["plugin"], wl["d"], params, data=wl.get("_data"))
    nw = wl["nw"]
    # 1) one thread, the smallest legal ensemble: a reliable per-walker-step cost without OpenMP noise
    n1 = min(nw, 2 * ((wl["d"] + 3) // 2))
    oracle.emcee(dens, x0[:n1], 1, 0, 1, 2.0, seed=1, store=False, nthreads=1, native=True)   # library load
    t0 = time.perf_counter()
    oracle.emcee(dens, x0[:n1], 1, 0, 1, 2.0, seed=1, store=False, nthreads=1, native=True)
    per_ws1 = max((time.perf_counter() - t0) / n1, 1e-9)
    # 2) all threads on a sub-ensemble sized for ~0.5 s if the threads scaled perfectly (best of 2)
    ncal = int(min(nw, max(n1, 0.5 * nthreads / (2 * per_ws1))))
    ncal -= ncal % 2
    dt = float("inf")
    for _ in range(2):
        t0 = time.perf_counter()
        oracle.emcee(dens, x0[:ncal], 2, 1, 1, 2.0, seed=1, store=False, nthreads=nthreads, native=True)
        dt = min(dt, time.perf_counter() - t0)
    per_ws = max(dt / (2 * ncal), 1e-9)
    nw_s = nw
    if 2 * nw * per_ws > 1.5 * seconds:
        nw_s = int(max(ncal, min(nw, seconds / (2 * per_ws))))
        nw_s -= nw_s % 2
    iters = int(max(2, min(wl["niter_walker"], seconds / (per_ws * nw_s))))
    return dens, x0[:nw_s], nw_s, iters


def _cpu_run(wl, dens, x0, iters, seconds, nthreads, seed):
    """One timed CPU sample; if the calibration was pessimistic (a noisy host makes small OpenMP regions look slow)
    and the run ended in under a third of the budget, lengthen it and time again (at most twice)."""
    from oracle import oracle
    for _ in range(3):
        t0 = time.perf_counter()
        oracle.emcee(dens, x0, iters, iters // 2, max(1, iters // 5), 2.0, seed=seed, store=True, nthreads=nthreads,
                     native=True)
        dt = time.perf_counter() - t0
        if dt >= seconds / 3 or iters >= wl["niter_walker"]:
            break
        iters = int(min(wl["niter_walker"], max(iters + 1, iters * 0.8 * seconds / max(dt, 1e-6))))
    return iters, dt


def cpu_baseline(wl, seconds=15.0):
    nthreads = os.cpu_count() or 1
    dens, x0, nw_s, iters = _cpu_sample(wl, seconds, nthreads)
    iters, dt = _cpu_run(wl, dens, x0, iters, seconds, nthreads, 2)
    return {"value": nw_s * iters / dt, "unit": UNIT, "cores": nthreads, "kind": "port",
            "sample": f"{wl['plugin']} d={wl['d']}, {nw_s} walkers (of {wl['nw']}) x {iters} iterations "
                      f"(of {wl['niter_walker']}), {dt:.1f} s, C restatement of the reference loop with OpenMP"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="rosenbrock2d", choices=sorted(WORKLOADS))
    ap.add_argument("--launch-mode", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--tensor-cores", type=int, default=None, choices=[0, 1],
                    help="force the tcgen05 (1) or FP64 (0) log-density kernel of the dense plugins")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, wl)

    import torch
    import torch.distributed as dist

    import kissmcmc_b200 as km

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    d, nw, nitw, nthin = wl["d"], wl["nw"], wl["niter_walker"], wl["nthin"]
    nbw = nitw // 2
    ns = (nitw - nbw) // nthin
    params, x0 = make_inputs(wl, 1000 + rank)
    ld = km.LogDensity(wl["plugin"], d, params, data=wl.get("_data"), device=local)
    if args.tensor_cores is not None:
        ld.set_option("tensor_cores", args.tensor_cores)
    elif wl["plugin"] == "gaussian" and d > 16:
        ld.set_option("tensor_cores", 1)          # bench default for configs[2]: the tcgen05 Mahalanobis GEMM
    tensor = ld.info("tensor_cores") == 1.0
    stream = torch.cuda.Stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def new_sampler(step):
        # independent ensembles: rank-distinct Philox key and walker-id range
        s = km.Sampler(ld, x0, nitw, nbw, nthin, 2.0, seed=(step << 8) | rank, walker_id_base=rank * nw,
                       launch_mode=args.launch_mode)
        s.set_stream(stream.cuda_stream)
        return s

    walker_steps_per_step = nw * nitw
    alg_bytes_per_step = b_step(d) * nw * nitw + (8 * d + 8) * nw * ns

    # ---- value: device-timed, inputs resident in HBM -------------------------------------
    total = args.warmup + args.steps
    samplers = [new_sampler(i) for i in range(total)]
    with torch.cuda.stream(stream):
        for i in range(args.warmup):
            flush.fill_(i & 0xFF)
            samplers[i].run(-1, sync=False)
    barrier()
    clocks = ClockSampler(local)
    clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for i in range(args.warmup, total):
            flush.fill_(i & 0xFF)                     # L2 flush between steps (inside the timed region)
            samplers[i].run(-1, sync=False)
        ev1.record(stream)
    barrier()
    clk = clocks.stop()
    dev_ms = ev0.elapsed_time(ev1)
    kern_ms, launches = [], 0
    for i in range(args.warmup, total):
        ms, n = samplers[i].last_run_ms()
        kern_ms.append(ms)
        launches += n
    for s in samplers:
        s.close()

    # ---- e2e: public API, host buffers, H2D + D2H inside the timed region ------------------
    x0_pinned = torch.from_numpy(x0).pin_memory()
    out_th = torch.empty((nw, ns, d), dtype=torch.float64).pin_memory()
    out_lp = torch.empty((nw, ns), dtype=torch.float64).pin_memory()
    out_ar = torch.empty((nw,), dtype=torch.float64).pin_memory()

    def e2e_step(step):
        ta = time.perf_counter()
        s = km.Sampler(ld, x0_pinned.numpy(), nitw, nbw, nthin, 2.0, seed=(step << 8) | rank,
                       walker_id_base=rank * nw, launch_mode=args.launch_mode)
        tb = time.perf_counter()
        s.run(-1)
        tc = time.perf_counter()
        s.results(out_th.numpy(), out_lp.numpy(), out_ar.numpy())
        td = time.perf_counter()
        s.close()
        if os.environ.get("KMC_BENCH_VERBOSE"):
            print(f"e2e step {step}: create {1e3 * (tb - ta):.1f} run {1e3 * (tc - tb):.1f} results "
                  f"{1e3 * (td - tc):.1f} close {1e3 * (time.perf_counter() - td):.1f} ms", file=sys.stderr)

    e2e_steps = max(1, min(args.steps, 3))
    e2e_step(0)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step(100 + i)
    barrier()
    e2e_s = time.perf_counter() - t0

    t = torch.tensor([dev_ms, e2e_s * 1e3, max(kern_ms), float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = t.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = t.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        dev_ms, e2e_ms, launches = mx[0].item(), mx[1].item(), int(sm[3].item())
    else:
        e2e_ms = e2e_s * 1e3

    if rank == 0:
        peaks_path = ROOT / "MEASURED_PEAKS.json"
        if peaks_path.exists():
            peak, peak_src = json.loads(peaks_path.read_text())["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        k_ms = statistics.mean(kern_ms)
        achieved = alg_bytes_per_step / (k_ms * 1e-3) / 1e9
        traffic = None
        tpath = ROOT / "profiles" / "traffic.json"
        if tpath.exists():
            traffic = json.loads(tpath.read_text()).get(args.workload)
        if wl["plugin"] == "logistic":      # tensor-bound nominally: 2*d*N flops per walker-step
            peaks = json.loads(peaks_path.read_text()) if peaks_path.exists() else {}
            tpeak = peaks.get("bf16_tflops_sustained", 1400.0)
            tflops = 2.0 * d * wl["ndata"] * walker_steps_per_step / (k_ms * 1e-3) / 1e12
            roof = {"bound": "tensor", "achieved": tflops, "peak": tpeak, "unit": "TFLOP/s", "frac": tflops / tpeak,
                    "traffic": traffic, "kernel": "tc::logistic_tc_kernel" if tensor else "logistic_logp_kernel",
                    "peak_source": "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback",
                    "algorithmic_flops_per_step": 2.0 * d * wl["ndata"] * walker_steps_per_step,
                    "kernel_ms_per_step": k_ms,
                    "note": "algorithmic flops 2*d*N per walker-step; the tcgen05 kernel issues 3x that (theta split "
                            "into 3 bf16 pieces) and is bounded by the MUFU softplus epilogue (profiles/r1_summary.md)"}
        else:
            roof = None
        line = {
            "metric": METRIC, "value": world * walker_steps_per_step * args.steps / (dev_ms * 1e-3), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16x3 split operands, f32 accumulate (tcgen05); f64 state" if tensor else "f64",
            "data": "synthetic",
            "config": {
                "workload": args.workload, "tensor_cores": tensor, "description": wl["desc"], "plugin": wl["plugin"], "d": d,
                "nwalkers_per_gpu": nw, "niter_walker": nitw, "nburnin_walker": nbw, "nthin": nthin,
                "samples_per_walker": ns, "a_scale": 2.0, "rng": "philox4x32-10",
                "parallelism": "1 ensemble" if world == 1 else f"{world} independent ensembles (no collective)",
                "launch_mode": "persistent kernel, grid barrier per half-step" if args.launch_mode == 0
                else "one launch per half-step",
                "l2": "ensemble state is L2-resident by construction across the dependent half-steps of one step; "
                      "L2 flushed (256 MiB fill) between steps, inside the timed region",
            },
            "roofline": roof or {
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src,
                "kernel": dominant_kernel(wl, tensor, args.launch_mode),
                "algorithmic_bytes_per_launch": alg_bytes_per_step, "kernel_ms_per_launch": k_ms,
                "note": "algorithmic bytes = (24d+24) per walker-step + (8d+8) per stored sample; a state that fits "
                        "L2 / shared memory makes frac against the HBM copy peak able to exceed 1",
            },
            "e2e": {"value": world * walker_steps_per_step * e2e_steps / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": int(x0.nbytes),
                    "d2h_bytes_per_step": int(out_th.numel() * 8 + out_lp.numel() * 8 + nw * 4),
                    "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps},
            "gpu_launches": launches,
            "clocks": clk,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(wl)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
