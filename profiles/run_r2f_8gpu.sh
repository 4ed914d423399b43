#!/bin/bash
# round-2 multi-GPU pass (one box, 8 x B200): ONE workload over N devices from ONE host process (ns_multi_call_snv)
set -u
O=gpurun_out
nvidia-smi -L | wc -l > $O/r2f_ngpu.txt
python -m pytest tests/test_gpu_pool_multi.py -m gpu -q > $O/r2f_pytest.log 2>&1; echo "pytest rc $?" >> $O/r2f_pytest.log; tail -3 $O/r2f_pytest.log
python bench.py --scaling strong --gpus 8 --gpus-list 1,2,4,8 --steps 3 --warmup 1 > $O/r2f_strong_c2.jsonl 2> $O/r2f_strong_c2.err; echo "c2 rc $?"
python bench.py --scaling strong --config C4 --positions 2000000 --gpus 8 --gpus-list 1,2,4,8 --steps 1 --warmup 1 > $O/r2f_strong_c4.jsonl 2> $O/r2f_strong_c4.err; echo "c4 rc $?"
python bench.py --scaling strong --config C5 --positions 16000000 --gpus 8 --gpus-list 1,2,4,8 --steps 2 --warmup 1 > $O/r2f_strong_c5.jsonl 2> $O/r2f_strong_c5.err; echo "c5 rc $?"
python - <<'PY'
import json
for f in ("c2","c4","c5"):
    try:
        for ln in open(f"gpurun_out/r2f_strong_{f}.jsonl"):
            d=json.loads(ln)
            print(f, d['n_gpus'], 'ms', round(d['ms_per_step'],1), 'fits/s', round(d['value']), 'speedup', round(d['speedup_vs_first']['speedup'],2), 'same', d['detail']['identical_to_single_device_call'],
                  'per-dev chunks', [x['chunks'] for x in d['per_device']], 'wall', [round(x['wall_ms']) for x in d['per_device']], 'tail', [round(x['pool_tail_ms']) for x in d['per_device']])
    except Exception as e: print(f,'ERR',e)
PY
tail -3 $O/r2f_strong_c4.err $O/r2f_strong_c5.err
