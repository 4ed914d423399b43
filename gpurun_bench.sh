set -x
python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
tail -3 gpurun_out/bench_full.err
python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log
