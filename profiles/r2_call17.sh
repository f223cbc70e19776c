set -x
export KMC_LIB=$PWD/build/variants/push_w32.so
timeout 900 python -m pytest tests/test_gpu_push.py -m gpu -q 2>&1 | tail -8
timeout 200 python profiles/push_bench.py 24 10 0p 2>&1 | tail -1
timeout 400 python profiles/push_bench.py 24 10 0,1 0,0,0 0,0,-1 0,0,8000 0,0,30000 2>&1 | tail -4
KMC_LIB=$PWD/build/variants/push_w32prof.so timeout 300 python profiles/push_bench.py 24 10 0,1 0,0,0 2>&1 | grep -E "rank 0|mode" | tail -3
