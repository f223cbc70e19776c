#!/bin/bash
mkdir -p gpurun_out
timeout 60 python -m pytest tests -m gpu -x -q > gpurun_out/final2_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/final2_pytest.log; tail -4 gpurun_out/final2_pytest.log
timeout 40 python bench.py --workload gaussian100d --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/final2_bench_gaussian100d.json 2> gpurun_out/final2_bench_gaussian100d.err; cut -c1-200 gpurun_out/final2_bench_gaussian100d.json
