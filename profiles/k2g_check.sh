#!/bin/bash
mkdir -p gpurun_out
KMC_TC=1 timeout 50 compute-sanitizer --tool racecheck --print-limit 5 python profiles/prof_run.py gaussian100d 3 0 8192 > gpurun_out/k2g_racecheck.log 2>&1; tail -6 gpurun_out/k2g_racecheck.log
