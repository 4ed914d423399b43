#!/bin/bash
# round-2 first GPU pass: tests, pool on/off on every BASELINE shape, parity sweep (writes under gpurun_out/)
set -u
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r2a_pytest.log 2>&1; echo "pytest rc $?" >> $O/r2a_pytest.log
tail -5 $O/r2a_pytest.log
python bench.py --steps 5 --warmup 3 > $O/r2a_bench_c2_pool.json 2> $O/r2a_bench_c2_pool.err
NS_POOL=0 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > $O/r2a_bench_c2_nopool.json 2> $O/r2a_bench_c2_nopool.err
: > $O/r2a_other.jsonl
for c in "C1 10000" "C5 4000000" "C4 300000" "C3 4000"; do set -- $c
  for pool in 1 0; do
    NS_POOL=$pool python bench.py --config $1 --positions $2 --steps 3 --warmup 3 --no-cpu --no-e2e 2>/dev/null | tail -1 | sed "s/^{/{\"pool\": $pool, /" >> $O/r2a_other.jsonl
  done
done
python tests/parity_at_scale.py $O/parity_r2a.json > $O/parity_r2a.log 2>&1
tail -3 $O/parity_r2a.log
python - <<'PY'
import json
for f in ("gpurun_out/r2a_bench_c2_pool.json","gpurun_out/r2a_bench_c2_nopool.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value']), round(d['ms_per_step'],1), d['e2e'] and round(d['e2e']['value']), d.get('parity_check'), d['detail']['kernel_ms_per_step_rank0'], d['detail'].get('pool'))
    except Exception as e: print(f, 'ERR', e)
for ln in open('gpurun_out/r2a_other.jsonl'):
    try:
        d=json.loads(ln); print(d['config']['workload'][:34], 'pool', d['pool'], round(d['value']), round(d['ms_per_step'],2), d['detail']['kernel_ms_per_step_rank0'], d['detail'].get('pool'))
    except Exception as e: print('ERR', e)
PY
