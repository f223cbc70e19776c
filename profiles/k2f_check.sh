#!/bin/bash
mkdir -p gpurun_out
KMC_TC=1 timeout 60 python profiles/prof_run.py gaussian100d 200 0 > gpurun_out/k2f_times.log 2>&1; tail -3 gpurun_out/k2f_times.log
KMC_TC=1 KMC_LIB=$PWD/build/variants/libkmc_k2fprof.so timeout 60 python profiles/prof_run.py gaussian100d 200 0 > gpurun_out/k2f_phase_cycles.log 2>&1; grep "cta 0" gpurun_out/k2f_phase_cycles.log | tail -1
timeout 200 python -m pytest tests/test_gpu_batched.py -m gpu -x -q > gpurun_out/k2f_pytest.log 2>&1; tail -3 gpurun_out/k2f_pytest.log
