python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C2full', round(d['value']), round(d['ms_per_step'],1), {k:round(v,1) for k,v in d['detail']['kernel_ms_per_step_rank0'].items()}, 'frac', round(d['roofline']['frac'],3))"
