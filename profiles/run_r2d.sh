#!/bin/bash
# round-2 fourth GPU pass: full tests, evidence capture (bench line, ncu launch list, traffic, full-set), driver profile
set -u
O=gpurun_out
python -m pytest tests -m gpu -q > $O/r2d_pytest.log 2>&1; echo "pytest rc $?" >> $O/r2d_pytest.log
tail -4 $O/r2d_pytest.log
bash profiles/capture.sh r2 > $O/r2d_capture.log 2>&1; tail -3 $O/r2d_capture.log
python profiles/driver_profile.py > $O/r2d_driver_profile.log 2>&1; head -12 $O/r2d_driver_profile.log
python profiles/driver_case.py 120000 0 > $O/r2d_driver_case.log 2>&1; tail -2 $O/r2d_driver_case.log
python bench.py --scaling strong --gpus 1 --positions 300000 --steps 2 --warmup 1 > $O/r2d_strong1.json 2> $O/r2d_strong1.err; tail -c 600 $O/r2d_strong1.json; tail -3 $O/r2d_strong1.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2d_smoke.log 2>&1; tail -2 $O/r2d_smoke.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_full.json').read().strip().splitlines()[-1])
k=d['detail']['kernel_ms_per_step_rank0']
print('C2', round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],2), 'gate_frac', round(d['roofline_gate']['frac'],3), 'fit_frac', round(d['roofline']['frac'],3))
print(' text', d['detail'].get('text'))
print(' parity', {k:v for k,v in (d.get('parity_check') or {}).items() if k in ('units','fitted','calls','mismatches','nonconverged','nonconverged_compared','max_rel_sigma','max_abs_sigma')})
print(' cpu', d['cpu_baseline']['value'])
PY
