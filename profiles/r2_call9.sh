set -x
timeout 400 python profiles/push_bench.py 24 10 0,1 0,0,0 1024,384,0 2>&1 | tail -2
KMC_LIB=$PWD/build/variants/push_prof.so timeout 300 python profiles/push_bench.py 24 10 0,1 0,0,0 2>&1 | tail -5
KMC_LIB=$PWD/build/variants/push_prof.so timeout 300 python profiles/push_bench.py 24 10 0p 2>&1 | tail -3
