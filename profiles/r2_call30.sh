set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 3 --warmup 3 2>gpurun_out/r2_bench_torchrun_n8.err | tail -1 > gpurun_out/r2_bench_torchrun_n8.log; tail -3 gpurun_out/r2_bench_torchrun_n8.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_torchrun_n8.log").read().strip().splitlines()[-1])
print("N=8 value %.4e" % d["value"], "ms", d["ms_per_step"], "sharded", {k:v for k,v in d["sharded"].items() if k.startswith(("value","eff","weak"))})
PY
