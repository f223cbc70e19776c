"""CPU checks of bench.py's contract pieces that need no GPU: workloads = BASELINE.json's configs, the roofline arithmetic
(SURVEY.md section 8d), the kernel labels, the reference arm's JSON line, and the order of operations in the rank barrier."""
import inspect
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def test_workloads_are_the_baseline_configs():
    w = bench.WORKLOADS
    assert (w["exponential1d"]["nw"], w["exponential1d"]["d"]) == (100, 1)                  # configs[0]
    assert (w["rosenbrock2d"]["nw"], w["rosenbrock2d"]["d"], w["rosenbrock2d"]["niter_walker"]) == (1 << 20, 2, 10_000)
    assert (w["gaussian100d"]["nw"], w["gaussian100d"]["d"]) == (1 << 16, 100)              # configs[2]
    assert (w["logistic32d"]["nw"], w["logistic32d"]["d"], w["logistic32d"]["ndata"]) == (8192, 32, 10**6)
    assert (w["gaussian10d"]["nw"], w["gaussian10d"]["d"]) == (1 << 24, 10)                 # configs[4]
    cfgs = json.loads((ROOT / "BASELINE.json").read_text())["configs"]
    assert len(cfgs) == 5 and "2^20 walkers" in cfgs[1] and "2^24-walker 10-D" in cfgs[4]


def test_roofline_arithmetic():
    wl = bench.WORKLOADS["rosenbrock2d"]
    assert bench.b_step(2) == 72 and bench.b_step(10) == 264 and bench.b_step(100) == 2424   # 24 d + 24
    r = bench.roofline_of(wl, "rosenbrock2d", False, 0, 118.5, None)
    steps = (1 << 20) * 10_000
    alg = 72 * steps + (8 * 2 + 8) * (1 << 20) * 5                   # + 5 stored samples per walker
    assert r["algorithmic_bytes_per_launch"] == alg == 755100549120
    assert r["bound"] == "hbm" and r["unit"] == "GB/s"
    assert r["achieved"] == pytest.approx(alg / 0.1185 / 1e9) and r["frac"] == pytest.approx(r["achieved"] / r["peak"])
    rl = bench.roofline_of(bench.WORKLOADS["logistic32d"], "logistic32d", True, 0, 8.8, None)
    assert rl["bound"] == "tensor" and rl["algorithmic_flops_per_step"] == 2.0 * 32 * 10**6 * 8192 * 4


def test_dominant_kernel_labels_follow_the_library_rules():
    w = bench.WORKLOADS
    assert bench.dominant_kernel(w["rosenbrock2d"], False, 0) == "emcee_smem_kernel"
    assert bench.dominant_kernel(w["exponential1d"], False, 0) == "emcee_smem_kernel"
    assert bench.dominant_kernel(w["gaussian10d"], False, 0) == "emcee_bulk_kernel"
    assert bench.dominant_kernel(w["gaussian100d"], True, 0) == "tc::gaussian_fused2_kernel"
    big = dict(w["rosenbrock2d"], nw=1 << 22)                         # does not fit 7 rounds of 256 per CTA any more
    assert bench.dominant_kernel(big, False, 0) == "emcee_run_kernel"


def test_rank_barrier_synchronizes_before_the_collective():
    """dist.barrier() with this rank's cooperative launches still queued cost one random rank +12 ms per step
    (profiles/r2_call34.log): the device is drained first."""
    src = inspect.getsource(bench.Env.barrier)
    code = [ln.strip() for ln in src.splitlines() if ln.strip() and not ln.strip().startswith("#")]
    assert code.index("self.torch.cuda.synchronize()") < code.index("self.dist.barrier()")
    assert code.count("self.torch.cuda.synchronize()") == 2            # and again after it


def test_reference_arm_line():
    """`bench.py --impl reference`: the oracle timed on the host cores, same metric / unit / config keys, no GPU touched."""
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == bench.UNIT
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 1e6
    assert d["config"]["workload"] == "rosenbrock2d"
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
