#!/bin/bash
# Round-end validation on one B200: GPU parity suite, bench lines of the default and the K2F workload,
# and an ncu --set full extract of the fused Gaussian kernel.  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 480 python -m pytest tests -m gpu -x -q --durations=12 > gpurun_out/final_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/final_pytest.log
tail -3 gpurun_out/final_pytest.log
timeout 150 python bench.py --workload gaussian100d --steps 5 --warmup 3 > gpurun_out/final_bench_gaussian100d.json 2> gpurun_out/final_bench_gaussian100d.err
timeout 200 python bench.py --steps 5 --warmup 3 > gpurun_out/final_bench_default.json 2> gpurun_out/final_bench_default.err
KMC_TC=1 timeout 60 python profiles/prof_run.py gaussian100d 200 0 > gpurun_out/final_k2f_times.log 2>&1
source profiles/capture_final.sh.lib
KMC_TC=1 cap k2f_gaussian_fused gaussian_fused python profiles/prof_run.py gaussian100d 20 0
cat gpurun_out/final_bench_gaussian100d.json gpurun_out/final_bench_default.json gpurun_out/final_k2f_times.log
