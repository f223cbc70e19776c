set -x
timeout 600 python profiles/push_bench.py 24 10 0,1,2,3,4,5,6,7 0,0,0 0,0,256 0,0,448 2>&1 | tail -3
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 8 --steps 3 --warmup 3) > gpurun_out/r2_bench_n8.log 2>&1; tail -4 gpurun_out/r2_bench_n8.log | cut -c1-4000
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus 4 --steps 3 --warmup 3) > gpurun_out/r2_bench_n4.log 2>&1; tail -4 gpurun_out/r2_bench_n4.log | cut -c1-600
timeout 600 python -m pytest tests/test_gpu_push.py -m gpu -q 2>&1 | tail -3
