make -s -C oracle
python -m pytest tests -q -m gpu -x 2>&1 | tail -2
for c in "C2 300000" "C5 1000000" "C1 300000"; do set -- $c
python bench.py --config $1 --steps 3 --warmup 2 --positions $2 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value']), round(d['ms_per_step'],1), 'gate ms', round(d['detail']['kernel_ms_per_step_rank0']['gate_ms'],3), 'gate GB/s', round(d['roofline_gate']['achieved']), 'frac', round(d['roofline_gate']['frac'],3))"
done
