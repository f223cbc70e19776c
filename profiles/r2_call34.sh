set -x
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,power.limit,temperature.gpu --format=csv
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 3 --warmup 3 --no-sharded 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.4e ms %.2f per_rank %s' % (d['value'], d['ms_per_step'], d['per_rank']))"; }
run 29521
CUDA_VISIBLE_DEVICES=1,0 run 29522
for g in 0 1; do CUDA_VISIBLE_DEVICES=$g python bench.py --steps 3 --warmup 3 --no-sharded --no-others --no-traffic --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('gpu $g alone: ms %.2f' % d['ms_per_step'])"; done
