#!/bin/bash
# round-2 last multi-GPU pass (8 x B200) on the final build: C2 strong scaling with the default chunking (2 chunks per device, dealt evenly)
set -u
O=gpurun_out
python bench.py --scaling strong --gpus 8 --gpus-list 1,2,4,8 --steps 3 --warmup 1 > $O/r2l_strong_c2.jsonl 2> $O/r2l_strong_c2.err; echo "c2 rc $?"
python - <<'PY'
import json
for ln in open("gpurun_out/r2l_strong_c2.jsonl"):
    d=json.loads(ln)
    print(d['n_gpus'], 'ms', round(d['ms_per_step'],1), 'fits/s', round(d['value']), 'speedup', round(d['speedup_vs_first']['speedup'],2), 'same', d['detail']['identical_to_single_device_call'],
          'chunks', [x['chunks'] for x in d['per_device']], 'wall', [round(x['wall_ms']) for x in d['per_device']], 'tail', [round(x['pool_tail_ms']) for x in d['per_device']])
PY
