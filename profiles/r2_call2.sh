set -x
export KMC_LIB=$PWD/build/variants/push_prof.so
timeout 120 python profiles/push_bench.py 24 10 0p 2>&1 | tail -3
timeout 120 python profiles/push_bench.py 24 10 0,0 2>&1 | tail -5
timeout 120 python profiles/push_bench.py 24 10 0,0 1024 2>&1 | tail -5
timeout 120 python profiles/push_bench.py 24 10 0,0,0,0,0,0,0,0 2>&1 | tail -17
unset KMC_LIB
(time timeout 900 python bench.py) 2>&1 | tail -12
