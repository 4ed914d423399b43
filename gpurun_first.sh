set -x
make -s -C oracle
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -30
python __graft_entry__.py --smoke
python bench.py --steps 2 --warmup 1 --positions 100000 --cpu-seconds 5
