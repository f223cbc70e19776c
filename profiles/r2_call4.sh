set -x
nvidia-smi -L
nvidia-smi topo -m | head -8
timeout 900 python -m pytest tests/test_gpu_push.py tests/test_distributed.py -m gpu -v 2>&1 | tail -40 | tee gpurun_out/r2_pytest_2gpu.log
timeout 200 python profiles/push_bench.py 24 10 0 2>&1 | tail -1
timeout 400 python profiles/push_bench.py 24 10 0,1 0,0,0 0,0,200 0,0,1000 0,0,4000 0,0,16384 512,256,0 1024,256,0 2>&1 | tail -7
KMC_LIB=$PWD/build/variants/push_b4.so timeout 300 python profiles/push_bench.py 24 10 0,1 0,0,0 0,0,4000 2>&1 | tail -2
KMC_LIB=$PWD/build/variants/push_t128.so timeout 300 python profiles/push_bench.py 24 10 0,1 0,0,0 0,0,4000 2>&1 | tail -2
