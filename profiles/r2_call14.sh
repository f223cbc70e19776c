set -x
timeout 900 python -m pytest tests/test_gpu_push.py -m gpu -q -x 2>&1 | tail -3
timeout 900 python profiles/push_bench.py 24 10 0,1,2,3,4,5,6,7 0,0,0 0,0,128 0,0,256 0,0,384 0,0,512 0,0,768 0,0,1024 2>&1 | tail -7
KMC_LIB=$PWD/build/variants/push_prof.so timeout 300 python profiles/push_bench.py 24 10 0,1,2,3,4,5,6,7 0,0,384 2>&1 | grep -E "rank 0|mode" | tail -3
timeout 600 python profiles/push_bench.py 24 10 0,1,2,3 0,0,0 0,0,512 0,0,1024 0,0,2048 2>&1 | tail -4
timeout 600 python profiles/push_bench.py 24 10 0,1 0,0,0 0,0,2048 0,0,4096 0,0,8192 2>&1 | tail -4
