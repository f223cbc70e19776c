#!/bin/bash
mkdir -p gpurun_out
./build/mb/ffma2_bench > gpurun_out/ffma2_bench.log 2>&1; cat gpurun_out/ffma2_bench.log
KMC_TC=1 KMC_LIB=$PWD/build/variants/libkmc_k2fprof.so timeout 60 python profiles/prof_run.py gaussian100d 200 0 > gpurun_out/k2f_phase_cycles.log 2>&1; tail -8 gpurun_out/k2f_phase_cycles.log
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_k3.log 2>&1; tail -3 gpurun_out/pytest_k3.log
timeout 150 python bench.py --workload logistic32d --steps 5 --warmup 3 > gpurun_out/bench_logistic32d.json 2> gpurun_out/bench_logistic32d.err; cat gpurun_out/bench_logistic32d.json
