timeout 300 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -2 gpurun_out/r2_bench.err; python -c "
import json
d=json.loads(open('gpurun_out/r2_bench.json').read().strip().splitlines()[-1]); print('value %.4e ms %.2f frac %.3f e2e %.4e traffic %s others %d' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['roofline']['traffic'], len(d['others'])))"
