set -x
timeout 900 python -m pytest tests/test_gpu_push.py tests/test_distributed.py -m gpu -q -rs 2>&1 | tail -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r2_bench_torchrun_n2.log; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_torchrun_n2.log").read().strip().splitlines()[-1])
print("N=2 value %.4e" % d["value"], "sharded", {k:v for k,v in d["sharded"].items() if k.startswith(("value","eff"))})
PY
