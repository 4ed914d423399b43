#!/bin/bash
# round-2 third multi-GPU pass (8 x B200): share-dependent chunk count per device (1 / 2 / 4), A/B against the fixed 4 per device
set -u
O=gpurun_out
python bench.py --scaling strong --gpus 8 --gpus-list 1,2,4,8 --steps 3 --warmup 1 > $O/r2i_strong_c2.jsonl 2> $O/r2i_strong_c2.err; echo "c2 rc $?"
NS_MULTI_CHUNK_POSITIONS=31250 python bench.py --scaling strong --gpus 8 --gpus-list 8 --steps 3 --warmup 1 > $O/r2i_strong_c2_4perdev.jsonl 2> $O/r2i_strong_c2_4perdev.err; echo "c2 4/dev rc $?"
NS_MULTI_CHUNK_POSITIONS=62500 python bench.py --scaling strong --gpus 8 --gpus-list 8 --steps 3 --warmup 1 > $O/r2i_strong_c2_2perdev.jsonl 2> $O/r2i_strong_c2_2perdev.err; echo "c2 2/dev rc $?"
python bench.py --scaling strong --config C4 --positions 2000000 --gpus 8 --gpus-list 1,4,8 --steps 1 --warmup 1 > $O/r2i_strong_c4.jsonl 2> $O/r2i_strong_c4.err; echo "c4 rc $?"
python -m pytest tests/test_gpu_pool_multi.py -m gpu -q 2>&1 | tail -2
python - <<'PY'
import json
for f in ("c2","c2_4perdev","c2_2perdev","c4"):
    try:
        for ln in open(f"gpurun_out/r2i_strong_{f}.jsonl"):
            d=json.loads(ln)
            print(f, d['n_gpus'], 'ms', round(d['ms_per_step'],1), 'fits/s', round(d['value']), 'speedup', round(d['speedup_vs_first']['speedup'],2), 'same', d['detail']['identical_to_single_device_call'],
                  'chunks', [x['chunks'] for x in d['per_device']], 'wall', [round(x['wall_ms']) for x in d['per_device']], 'tail', [round(x['pool_tail_ms']) for x in d['per_device']])
    except Exception as e: print(f,'ERR',e)
PY
