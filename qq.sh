run() {
python bench.py --config $1 --steps 2 --warmup 1 --positions $2 --no-cpu $3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 $2', round(d['value']), 'proc/s', round(d['detail']['site_alleles_processed_per_s']), round(d['ms_per_step'],1), 'e2e', d['e2e'] and round(d['e2e']['ms_per_step'],1), {k:round(v,1) for k,v in d['detail']['kernel_ms_per_step_rank0'].items()}, 'heavy', d['detail']['heavy_units_per_step_rank0'], 'fitted', d['detail']['fitted_per_step_rank0'])"
}
run C5 4000000
run C1 10000 --no-e2e
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
