#!/bin/bash
mkdir -p gpurun_out
KMC_TC=1 KMC_FUSED_VARIANT=2 KMC_LIB=$PWD/build/variants/libkmc_k2fprof.so timeout 60 python profiles/prof_run.py gaussian100d 200 0 > gpurun_out/k2g_phase_cycles.log 2>&1; tail -4 gpurun_out/k2g_phase_cycles.log
