#!/bin/bash
# round-2 second multi-GPU pass (8 x B200) on the final build (bounded residency wait): strong scaling, one host process
set -u
O=gpurun_out
python bench.py --scaling strong --gpus 8 --gpus-list 1,2,4,8 --steps 3 --warmup 1 > $O/r2h_strong_c2.jsonl 2> $O/r2h_strong_c2.err; echo "c2 rc $?"
python bench.py --scaling strong --config C4 --positions 2000000 --gpus 8 --gpus-list 1,2,4,8 --steps 1 --warmup 1 > $O/r2h_strong_c4.jsonl 2> $O/r2h_strong_c4.err; echo "c4 rc $?"
NS_MULTI_CHUNK_POSITIONS=41728 python bench.py --scaling strong --config C4 --positions 2000000 --gpus 8 --gpus-list 8 --steps 1 --warmup 1 > $O/r2h_strong_c4_fine.jsonl 2> $O/r2h_strong_c4_fine.err; echo "c4 fine rc $?"
NS_MULTI_CHUNK_POSITIONS=15680 python bench.py --scaling strong --gpus 8 --gpus-list 8 --steps 3 --warmup 1 > $O/r2h_strong_c2_fine.jsonl 2> $O/r2h_strong_c2_fine.err; echo "c2 fine rc $?"
python - <<'PY'
import json
for f in ("c2","c4","c4_fine","c2_fine"):
    try:
        for ln in open(f"gpurun_out/r2h_strong_{f}.jsonl"):
            d=json.loads(ln)
            print(f, d['n_gpus'], 'ms', round(d['ms_per_step'],1), 'fits/s', round(d['value']), 'speedup', round(d['speedup_vs_first']['speedup'],2), 'same', d['detail']['identical_to_single_device_call'],
                  'chunks', [x['chunks'] for x in d['per_device']], 'wall', [round(x['wall_ms']) for x in d['per_device']], 'tail', [round(x['pool_tail_ms']) for x in d['per_device']])
    except Exception as e: print(f,'ERR',e)
PY
