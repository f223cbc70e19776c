set -x
timeout 900 python -m pytest tests/test_gpu_push.py -m gpu -q -x 2>&1 | tail -3
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 3 --warmup 3) > gpurun_out/r2_bench_n2.log 2>&1; tail -4 gpurun_out/r2_bench_n2.log | cut -c1-3000
(time timeout 900 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1) 2>&1 | tail -5 | cut -c1-600
