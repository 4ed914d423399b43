#!/bin/bash
# round-2 last 1-GPU pass on the final build (cluster kernel with shared dnbinom_mu / digamma copies): tests, evidence capture,
# other shapes, parity sweep, and a full-set capture of the cluster kernel's throughput form on a C3-shaped batch
set -u
O=gpurun_out
python -m pytest tests -m gpu -q > $O/r2m_pytest.log 2>&1; echo "pytest rc $?" >> $O/r2m_pytest.log
tail -3 $O/r2m_pytest.log
bash profiles/capture.sh r2 > $O/r2m_capture.log 2>&1; tail -2 $O/r2m_capture.log
: > $O/bench_r2_other_configs.jsonl
for c in "C1 10000" "C3 4000 --no-e2e" "C4 300000" "C5 4000000" "C5 4000000 --rows-only"; do set -- $c
python bench.py --config $1 --positions $2 --steps 3 --warmup 3 --no-cpu ${3:-} 2>/dev/null | tail -1 >> $O/bench_r2_other_configs.jsonl
done
python tests/parity_at_scale.py $O/parity_r2.json > $O/parity_r2.log 2>&1; tail -1 $O/parity_r2.log
python profiles/profile_case.py 1000 1e-4 C3 > $O/prof_r2_c3_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k 'regex:^k_fit_heavy_tp$' -c 1 -f -o $O/prof_r2_c3_heavy_tp \
    python profiles/profile_case.py 1000 1e-4 C3 > $O/ncu_c3_r2.log 2>&1; tail -2 $O/ncu_c3_r2.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_full.json').read().strip().splitlines()[-1])
print('C2', round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'gate_frac', round(d['roofline_gate']['frac'],3), 'fit_frac', round(d['roofline']['frac'],3), 'text', d['detail']['text'] and round(d['detail']['text'].get('driver_site_alleles_per_s',0)), 'parity', d['parity_check'] and d['parity_check'].get('mismatches'))
for ln in open('gpurun_out/bench_r2_other_configs.jsonl'):
    d=json.loads(ln); k=d['detail']['kernel_ms_per_step_rank0']
    print(d['config']['workload'][:30], round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', d['e2e'] and round(d['e2e']['value']), 'gate', round(k['gate_ms'],2), 'fit', round(k['fit_ms'],1), 'tail', round(k['pool_tail_ms'],1), 'd2h', round(k['d2h_ms'],2))
PY
