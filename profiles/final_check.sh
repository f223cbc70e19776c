#!/bin/bash
# Round-end validation on one B200: smoke, GPU parity suite, bench lines of the default workload and the other
# BASELINE.json configs.  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 120 python __graft_entry__.py smoke > gpurun_out/final_smoke.log 2>&1; tail -1 gpurun_out/final_smoke.log
timeout 300 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/final_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/final_pytest.log
tail -3 gpurun_out/final_pytest.log
timeout 200 python bench.py --steps 5 --warmup 3 > gpurun_out/final_bench_default.json 2> gpurun_out/final_bench_default.err
for W in gaussian100d logistic32d gaussian10d exponential1d; do
  timeout 200 python bench.py --workload $W --steps 5 --warmup 3 > gpurun_out/final_bench_$W.json 2> gpurun_out/final_bench_$W.err
done
timeout 100 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/final_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("final_bench_")[1], "%.3e" % d["value"], "e2e %.3e" % d["e2e"]["value"], "frac", d.get("roofline", {}).get("frac"),
              "cpu", d.get("cpu_baseline", {}).get("value"))
    except Exception as e:
        print(f, "unreadable", e)
PY
