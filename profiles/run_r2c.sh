#!/bin/bash
# round-2 third GPU pass: full tests, gate forms, rows-only C5, driver profile, shared-cudart experiment
set -u
O=gpurun_out
python -m pytest tests -m gpu -q > $O/r2c_pytest.log 2>&1; echo "pytest rc $?" >> $O/r2c_pytest.log
tail -6 $O/r2c_pytest.log
: > $O/r2c_other.jsonl
run() { tag=$1; shift; "$@" 2>>$O/r2c_other.err | tail -1 | sed "s/^{/{\"tag\": \"$tag\", /" >> $O/r2c_other.jsonl; }
run c2 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e
NS_GATE_STAGED=1 run c2_staged python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e
run c5 python bench.py --config C5 --positions 4000000 --steps 3 --warmup 3 --no-cpu --no-e2e
run c5_rows_only python bench.py --config C5 --positions 4000000 --steps 3 --warmup 3 --no-cpu --no-e2e --rows-only
NS_GATE_STAGED=1 run c5_staged python bench.py --config C5 --positions 4000000 --steps 3 --warmup 3 --no-cpu --no-e2e
run c2_n64 python bench.py --samples 64 --positions 1000000 --steps 3 --warmup 3 --no-cpu --no-e2e
run c2_n128 python bench.py --samples 128 --positions 600000 --steps 3 --warmup 3 --no-cpu --no-e2e
run c1 python bench.py --config C1 --positions 10000 --steps 3 --warmup 3 --no-cpu --no-e2e
python profiles/driver_profile.py > $O/r2c_driver_profile.log 2>&1; head -12 $O/r2c_driver_profile.log
# shared CUDA runtime: does the library work next to torch's own libcudart?
cp needlestack_b200/lib/libns_b200.so /tmp/libns_static.so
NS_CUDART=shared python -m needlestack_b200.build --force > $O/r2c_cudart_shared.log 2>&1
ldd needlestack_b200/lib/libns_b200.so | grep -i cudart >> $O/r2c_cudart_shared.log
python -m pytest tests/test_gpu_golden.py tests/test_gpu_pool_multi.py -m gpu -q >> $O/r2c_cudart_shared.log 2>&1; echo "shared-cudart pytest rc $?" >> $O/r2c_cudart_shared.log
python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e 2>>$O/r2c_cudart_shared.log | tail -1 | cut -c1-200 >> $O/r2c_cudart_shared.log
tail -4 $O/r2c_cudart_shared.log
cp /tmp/libns_static.so needlestack_b200/lib/libns_b200.so
python - <<'PY'
import json
for ln in open('gpurun_out/r2c_other.jsonl'):
    try:
        d=json.loads(ln); k=d['detail']['kernel_ms_per_step_rank0']
        print(d['tag'], d['config']['workload'][:28], 'val', round(d['value']), 'ms', round(d['ms_per_step'],2), 'gate', round(k['gate_ms'],3), 'gate_frac', round(d['roofline_gate']['frac'],3), 'fit', round(k['fit_ms'],1), 'pack', round(k['pack_ms'],2), 'd2h', round(k['d2h_ms'],2), 'tail', round(k['pool_tail_ms'],2))
    except Exception as e: print('ERR', e, ln[:200])
PY
