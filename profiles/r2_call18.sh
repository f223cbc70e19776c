set -x
timeout 200 python profiles/push_bench.py 24 10 0 2>&1 | tail -1
KMC_LIB=$PWD/build/variants/store_all.so timeout 200 python profiles/push_bench.py 24 10 0 2>&1 | tail -1
timeout 200 python profiles/push_bench.py 24 10 0p 2>&1 | tail -1
timeout 400 python profiles/push_bench.py 24 10 0,1 0,0,0 0,0,2048 0,0,4096 0,0,-1 2>&1 | tail -4
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -5
