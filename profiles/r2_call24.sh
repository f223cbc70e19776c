set -x
export KMC_TC=1
timeout 300 python profiles/prof_run.py gaussian100d 400 0 2>&1 | tail -3
KMC_LIB=$PWD/build/variants/k2g_prof.so timeout 300 python profiles/prof_run.py gaussian100d 100 0 2>&1 | grep "K2G cta" | tail -4
. profiles/capture_final.sh.lib
cap k2g_acq gaussian_fused2 python profiles/prof_run.py gaussian100d 50 0
cp /tmp/k2g_acq.ncu-rep gpurun_out/
