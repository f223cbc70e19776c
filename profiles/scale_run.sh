#!/bin/bash
# Multi-GPU measurements on one node: independent ensembles (bench.py, weak scaling) and the
# sharded 2^24-walker 10-D Gaussian ensemble (strong scaling) in all-gather and peer mode.
out=gpurun_out/scale_r1.jsonl; : > $out
port=29600
for n in "$@"; do
  port=$((port+1))
  if [ "$n" = "1" ]; then
    python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 >> $out
    python profiles/sharded_bench.py 24 10 peer 2>/dev/null | tail -1 >> $out
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 3 --warmup 3 2>/dev/null | tail -1 >> $out
    port=$((port+1))
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port profiles/sharded_bench.py 24 10 peer 2>/dev/null | tail -1 >> $out
    port=$((port+1))
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port profiles/sharded_bench.py 24 10 allgather 2>/dev/null | tail -1 >> $out
  fi
done
python - <<'PY'
import json
for line in open("gpurun_out/scale_r1.jsonl"):
    line=line.strip()
    if not line.startswith("{"): continue
    d=json.loads(line)
    if "mode" in d: print(d["mode"], d["n_gpus"], "%.3e walker-steps/s" % d["walker_steps_per_s"], "%.3f ms/half-step" % d["ms_per_halfstep"])
    else: print("independent", d["n_gpus"], "%.3e walker-steps/s" % d["value"], "%.1f ms/step" % d["ms_per_step"], "e2e %.3e" % d["e2e"]["value"])
PY
