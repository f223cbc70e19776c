set -x
export KMC_TC=1
timeout 300 python profiles/prof_run.py gaussian100d 400 0 2>&1 | tail -2
KMC_LIB=$PWD/build/variants/k2g_half.so timeout 300 python profiles/prof_run.py gaussian100d 400 0 2>&1 | tail -2
