set -x
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 3 --warmup 3 $2 2>/dev/null | tail -1 > gpurun_out/r2_n2_$1.json; python -c "
import json,sys
d=json.loads(open('gpurun_out/r2_n2_$1.json').read()); print('value %.4e ms %.2f per_rank %s sharded %s' % (d['value'], d['ms_per_step'], d['per_rank'], {k:v for k,v in d.get('sharded',{}).items() if k.startswith(('value_p','eff'))}))"; }
run 29521 --no-sharded
run 29522 --no-sharded
run 29523 --no-sharded
run 29524
