set -x
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_push.py -x -q 2>&1 | tail -15
timeout 120 python profiles/push_bench.py 24 10 0 2>&1 | tail -1
timeout 120 python profiles/push_bench.py 24 10 0p 2>&1 | tail -1
timeout 120 python profiles/push_bench.py 24 10 0p 1024 2>&1 | tail -1
timeout 120 python profiles/push_bench.py 24 10 0,0 2>&1 | tail -1
timeout 120 python profiles/push_bench.py 24 10 0,0,0,0 2>&1 | tail -1
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
