#!/bin/bash
# round-2 second GPU pass: full test suite, C2 headline line, staged vs register-load gate, other shapes, parity sweep,
# driver throughput (writes under gpurun_out/)
set -u
O=gpurun_out
python -m pytest tests -m gpu -q > $O/r2b_pytest.log 2>&1; echo "pytest rc $?" >> $O/r2b_pytest.log
tail -15 $O/r2b_pytest.log
python bench.py --steps 5 --warmup 3 > $O/r2b_bench_c2.json 2> $O/r2b_bench_c2.err
: > $O/r2b_other.jsonl
run() { tag=$1; shift; "$@" 2>>$O/r2b_other.err | tail -1 | sed "s/^{/{\"tag\": \"$tag\", /" >> $O/r2b_other.jsonl; }
NS_GATE_STAGED=0 run c2_gate_reg python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e
run c2_gate_staged python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e
NS_GATE_STAGES=2 run c2_gate_staged_s2 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e
NS_GATE_STAGED=0 run c5_gate_reg python bench.py --config C5 --positions 4000000 --steps 3 --warmup 3 --no-cpu
run c5 python bench.py --config C5 --positions 4000000 --steps 3 --warmup 3 --no-cpu
run c5_rows_only python bench.py --config C5 --positions 4000000 --steps 3 --warmup 3 --no-cpu --rows-only
run c1 python bench.py --config C1 --positions 10000 --steps 3 --warmup 3 --no-cpu
run c4 python bench.py --config C4 --positions 300000 --steps 3 --warmup 3 --no-cpu
run c3 python bench.py --config C3 --positions 4000 --steps 3 --warmup 3 --no-cpu --no-e2e
run c2_n64 python bench.py --samples 64 --positions 1000000 --steps 3 --warmup 3 --no-cpu --no-e2e
python tests/parity_at_scale.py $O/parity_r2b.json > $O/parity_r2b.log 2>&1
tail -2 $O/parity_r2b.log
python profiles/driver_case.py 120000 0 > $O/r2b_driver.log 2>&1; tail -3 $O/r2b_driver.log
python - <<'PY'
import json
def show(d, tag=""):
    k=d['detail']['kernel_ms_per_step_rank0']
    print(tag, d['config']['workload'][:30], 'val', round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', d['e2e'] and round(d['e2e']['value']), d['e2e'] and round(d['e2e']['ms_per_step'],2),
          'gate', round(k['gate_ms'],3), 'fit', round(k['fit_ms'],1), 'post', round(k['post_ms'],2), 'pack', round(k['pack_ms'],2), 'd2h', round(k['d2h_ms'],2), 'tail', round(k['pool_tail_ms'],2),
          'gate_frac', round(d['roofline_gate']['frac'],3), 'fit_frac', round(d['roofline']['frac'],3))
try:
    d=json.loads(open('gpurun_out/r2b_bench_c2.json').read().strip().splitlines()[-1]); show(d,'C2 headline'); print(' parity', d.get('parity_check')); print(' text', d['detail'].get('text')); print(' cpu', d.get('cpu_baseline'))
except Exception as e: print('ERR c2', e)
for ln in open('gpurun_out/r2b_other.jsonl'):
    try:
        d=json.loads(ln); show(d, d['tag'])
    except Exception as e: print('ERR', e, ln[:200])
PY
