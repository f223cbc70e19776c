set -x
timeout 600 python -m pytest tests/test_gpu_push.py -m gpu -q 2>&1 | tail -3
timeout 600 python profiles/push_bench.py 24 10 0,1,2,3,4,5,6,7 0,0,0 0,0,192 0,0,448 2>&1 | tail -3
KMC_LIB=$PWD/build/variants/push_prof.so timeout 300 python profiles/push_bench.py 24 10 0,1,2,3,4,5,6,7 0,0,0 2>&1 | grep -E "rank 0|mode" | tail -3
timeout 600 python profiles/push_bench.py 24 10 0,1,2,3 0,0,0 2>&1 | tail -1
timeout 600 python profiles/push_bench.py 24 10 0,1 0,0,0 2>&1 | tail -1
