make -s -C oracle
python -m pytest tests -q -m gpu -x 2>&1 | tail -3
python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -2 gpurun_out/bench_full.err
python bench.py --config C5 --positions 1000000 --steps 2 --warmup 1 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C5', round(d['value']), round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value']), {k:round(v,1) for k,v in d['detail']['kernel_ms_per_step_rank0'].items()}, 'proc/s', round(d['detail']['site_alleles_processed_per_s']), 'gate GB/s', round(d['roofline_gate']['achieved']))"
