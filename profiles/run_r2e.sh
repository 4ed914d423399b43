#!/bin/bash
# round-2 fifth GPU pass: chains per lane in the warp kernel (2 vs 4), equal chunks, driver from a file
set -u
O=gpurun_out
: > $O/r2e_other.jsonl
run() { tag=$1; shift; "$@" 2>>$O/r2e_other.err | tail -1 | sed "s/^{/{\"tag\": \"$tag\", /" >> $O/r2e_other.jsonl; }
NS_FIT_CHAINS=2 run c2_ch2 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e
NS_FIT_CHAINS=4 run c2_ch4 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e
NS_FIT_CHAINS=2 run c4_ch2 python bench.py --config C4 --positions 300000 --steps 2 --warmup 2 --no-cpu --no-e2e
NS_FIT_CHAINS=4 run c4_ch4 python bench.py --config C4 --positions 300000 --steps 2 --warmup 2 --no-cpu --no-e2e
NS_FIT_CHAINS=2 run n128_ch2 python bench.py --samples 128 --positions 600000 --steps 2 --warmup 2 --no-cpu --no-e2e
NS_FIT_CHAINS=4 run n128_ch4 python bench.py --samples 128 --positions 600000 --steps 2 --warmup 2 --no-cpu --no-e2e
NS_FIT_CHAINS=4 run c5_ch4 python bench.py --config C5 --positions 4000000 --steps 3 --warmup 3 --no-cpu --no-e2e
NS_FIT_CHAINS=2 run c5_ch2 python bench.py --config C5 --positions 4000000 --steps 3 --warmup 3 --no-cpu --no-e2e
python -m pytest tests/test_gpu_golden.py tests/test_gpu_parity.py tests/test_gpu_gate_paths.py -m gpu -q > $O/r2e_pytest.log 2>&1; echo "pytest rc $?" >> $O/r2e_pytest.log; tail -3 $O/r2e_pytest.log
python profiles/driver_profile.py > $O/r2e_driver_profile.log 2>&1; head -10 $O/r2e_driver_profile.log
python - <<'PY'
import json
for ln in open('gpurun_out/r2e_other.jsonl'):
    try:
        d=json.loads(ln); k=d['detail']['kernel_ms_per_step_rank0']
        print(d['tag'], d['config']['workload'][:28], 'val', round(d['value']), 'ms', round(d['ms_per_step'],2), 'gate', round(k['gate_ms'],3), 'fit', round(k['fit_ms'],1), 'fit_frac', round(d['roofline']['frac'],3), 'tail', round(k['pool_tail_ms'],2))
    except Exception as e: print('ERR', e, ln[:200])
PY
