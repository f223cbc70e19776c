set -x
nvidia-smi -L | wc -l
timeout 300 python profiles/push_bench.py 24 10 0 2>&1 | tail -1
timeout 300 python profiles/push_bench.py 24 10 0,1,2,3 0,0,0 2>&1 | tail -1
timeout 300 python profiles/push_bench.py 24 10 0,1,2,3,4,5,6,7 0,0,0 512,0,0 2>&1 | tail -2
KMC_LIB=$PWD/build/variants/push_prof.so timeout 300 python profiles/push_bench.py 24 10 0,1,2,3,4,5,6,7 0,0,0 2>&1 | grep -E "rank 0|rank 5|mode" | tail -5
timeout 600 python -m pytest tests/test_gpu_push.py -m gpu -q -x 2>&1 | tail -3
