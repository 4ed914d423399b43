set -x
make -s -C oracle
python -m pytest tests/test_gpu_parity.py -q -m gpu 2>&1 | tail -15
python bench.py --steps 2 --warmup 1 --positions 100000 --no-cpu --no-e2e --dump-cycles gpurun_out/cyc_g.npz
