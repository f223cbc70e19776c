set -x
timeout 600 python -m pytest tests/test_gpu_push.py -x -q 2>&1 | tail -5
for lib in push_prof push_prof128; do
export KMC_LIB=$PWD/build/variants/$lib.so
timeout 120 python profiles/push_bench.py 24 10 0p 2>&1 | tail -2
timeout 120 python profiles/push_bench.py 24 10 0,0 2>&1 | tail -3
timeout 120 python profiles/push_bench.py 24 10 0,0 1024 2>&1 | tail -3
timeout 120 python profiles/push_bench.py 24 10 0,0,0,0,0,0,0,0 2>&1 | tail -4
done
export KMC_LIB=$PWD/build/variants/push_t128.so
timeout 600 python -m pytest tests/test_gpu_push.py -x -q 2>&1 | tail -5
unset KMC_LIB
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
