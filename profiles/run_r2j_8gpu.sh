#!/bin/bash
# round-2 fourth multi-GPU pass (8 x B200): even deal of the chunks when a device gets one or two; 1 against 2 chunks per device
set -u
O=gpurun_out
python bench.py --scaling strong --gpus 8 --gpus-list 1,2,4,8 --steps 3 --warmup 1 > $O/r2j_strong_c2.jsonl 2> $O/r2j_strong_c2.err; echo "c2 rc $?"
NS_MULTI_PIECES=2 python bench.py --scaling strong --gpus 8 --gpus-list 4,8 --steps 3 --warmup 1 > $O/r2j_strong_c2_2perdev.jsonl 2> $O/r2j_strong_c2_2perdev.err; echo "c2 2/dev rc $?"
NS_MULTI_PIECES=1 python bench.py --scaling strong --gpus 8 --gpus-list 2,4 --steps 3 --warmup 1 > $O/r2j_strong_c2_1perdev.jsonl 2> $O/r2j_strong_c2_1perdev.err; echo "c2 1/dev rc $?"
python -m pytest tests/test_gpu_pool_multi.py -m gpu -q 2>&1 | tail -2
python - <<'PY'
import json
for f in ("c2","c2_2perdev","c2_1perdev"):
    try:
        for ln in open(f"gpurun_out/r2j_strong_{f}.jsonl"):
            d=json.loads(ln)
            print(f, d['n_gpus'], 'ms', round(d['ms_per_step'],1), 'fits/s', round(d['value']), 'speedup', round(d['speedup_vs_first']['speedup'],2), 'same', d['detail']['identical_to_single_device_call'],
                  'chunks', [x['chunks'] for x in d['per_device']], 'wall', [round(x['wall_ms']) for x in d['per_device']], 'tail', [round(x['pool_tail_ms']) for x in d['per_device']])
    except Exception as e: print(f,'ERR',e)
PY
