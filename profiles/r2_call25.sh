set -x
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for w in exponential1d rosenbrock2d; do timeout 300 python bench.py --workload $w --no-others --no-sharded --no-traffic --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['workload'], 'value %.4e' % d['value'], 'ms', d['ms_per_step'], 'e2e %.4e' % d['e2e']['value'], 'frac %.3f' % d['roofline']['frac'])"; done
