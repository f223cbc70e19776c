#!/bin/bash
# Round-2 validation on one B200 (all outputs under gpurun_out/r2_*): smoke, the GPU suite, the default bench line (with
# others / sharded anchor / live traffic), the gaussian10d line with its own live traffic, the reference arm, the ncu
# launch list of the bench command (main workload), and ncu --set full extracts of K1, K1b, K2G and K3.
mkdir -p gpurun_out
timeout 120 python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1; tail -1 gpurun_out/r2_smoke.log
timeout 900 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r2_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_pytest_gpu.log; tail -3 gpurun_out/r2_pytest_gpu.log
timeout 400 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -c 300 gpurun_out/r2_bench.err
timeout 300 python bench.py --workload gaussian10d --steps 3 --warmup 3 --no-sharded > gpurun_out/r2_bench_gaussian10d.json 2> gpurun_out/r2_bench_gaussian10d.err
timeout 300 python bench.py --workload gaussian100d --steps 3 --warmup 3 > gpurun_out/r2_bench_gaussian100d.json 2> gpurun_out/r2_bench_gaussian100d.err
timeout 300 python bench.py --workload logistic32d --steps 3 --warmup 3 > gpurun_out/r2_bench_logistic32d.json 2> gpurun_out/r2_bench_logistic32d.err
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> /dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-traffic --no-sharded --no-others > gpurun_out/r2_launches_bench.log 2>&1
. profiles/capture_final.sh.lib
cap r2_k1_smem emcee_smem python profiles/prof_run.py rosenbrock2d 100 0
cap r2_k1b_bulk emcee_bulk python profiles/prof_run.py gaussian10d 6 0
KMC_TC=1 cap r2_k2g_gaussian_fused2 gaussian_fused2 python profiles/prof_run.py gaussian100d 50 0
KMC_TC=1 SKIP=2 cap r2_k3_logistic_tc logistic_tc_kernel python profiles/prof_run.py logistic32d 2 0
for n in r2_k1_smem r2_k1b_bulk r2_k2g_gaussian_fused2 r2_k3_logistic_tc; do mv gpurun_out/ncu_final_$n.csv gpurun_out/${n/r2_/r2_ncu_}.csv; done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "%.3e" % d["value"], "e2e %.3e" % d["e2e"]["value"], "frac %.3f" % d.get("roofline", {}).get("frac", 0),
              "traffic", d.get("roofline", {}).get("traffic"), "cpu", d.get("cpu_baseline", {}).get("value"))
        for o in d.get("others", []):
            print("   other", o["workload"], "%.3e" % o["value"], "frac %.3f" % o["roofline"]["frac"], "e2e %.3e" % o["e2e"]["value"])
    except Exception as e:
        print(f, "unreadable", e)
PY
