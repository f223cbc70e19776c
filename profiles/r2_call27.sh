set -x
timeout 900 python -m pytest tests/test_gpu_batched.py -m gpu -q -x -k logistic 2>&1 | tail -8
