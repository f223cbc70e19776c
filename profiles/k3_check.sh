#!/bin/bash
mkdir -p gpurun_out
source profiles/capture_final.sh.lib
for V in x_e16_p4 e16_p4; do
  KMC_LIB=$PWD/build/variants/libkmc_$V.so timeout 120 python profiles/k3_variants.py > gpurun_out/k3_variant_$V.log 2>&1
  tail -4 gpurun_out/k3_variant_$V.log
done
for V in x_e16_p4; do
  export KMC_LIB=$PWD/build/variants/libkmc_$V.so
  SKIP=2 cap k3_$V logistic_tc_kernel python profiles/prof_run.py logistic32d 2 0
done
