#!/bin/bash
mkdir -p gpurun_out
export KMC_LIB=$PWD/build/variants/libkmc_rowmap4.so
KMC_TC=1 timeout 25 python profiles/prof_run.py gaussian100d 200 0 > gpurun_out/k2g_rowmap4_times.log 2>&1; tail -2 gpurun_out/k2g_rowmap4_times.log
timeout 30 python -m pytest tests/test_gpu_batched.py -m gpu -x -q -k "tmem_variant" > gpurun_out/k2g_rowmap4_pytest.log 2>&1; tail -2 gpurun_out/k2g_rowmap4_pytest.log
