set -x
timeout 900 python -m pytest tests/test_gpu_push.py -m gpu -q -x 2>&1 | tail -3
timeout 900 python profiles/push_bench.py 24 10 0,1,2,3,4,5,6,7 0,0,0,1,3 0,0,0,3,1 0,0,0,2,2 0,0,64,2,2 0,0,128,2,2 0,0,256,2,2 0,0,512,2,2 0,0,256,1,1 0,0,256,3,3 2>&1 | tail -9
KMC_LIB=$PWD/build/variants/push_prof.so timeout 300 python profiles/push_bench.py 24 10 0,1,2,3,4,5,6,7 0,0,256,2,2 2>&1 | grep -E "rank 0|mode" | tail -3
timeout 600 python profiles/push_bench.py 24 10 0,1,2,3 0,0,0,2,2 0,0,128,2,2 0,0,512,2,2 2>&1 | tail -3
timeout 600 python profiles/push_bench.py 24 10 0,1 0,0,0,2,2 0,0,512,2,2 0,0,2048,2,2 2>&1 | tail -3
