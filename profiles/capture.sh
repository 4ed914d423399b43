#!/bin/bash
# Evidence capture on the GPU box (run through gpurun from the repo root):  bash profiles/capture.sh <tag>
# Every ncu pass follows a plain run of the same command that exited 0; numbers printed under ncu are never bench values.
set -u
TAG=${1:-r1}
O=gpurun_out
mkdir -p $O
# 1. the default bench line (C2: 100 samples x 1 M positions, 5 steps; device-resident, end-to-end, CPU oracle)
python bench.py > $O/bench_${TAG}_full.json 2> $O/bench_${TAG}_full.err || exit 1
# 2. launch list of a short bench run (plain run first)
python bench.py --steps 2 --warmup 1 --no-cpu > $O/bench_${TAG}_short.json 2> $O/bench_${TAG}_short.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > $O/ncu_launch_${TAG}.log 2>&1
# 3. DRAM traffic of the two roofline kernels inside the bench workload (per launch = per chunk of 262144 positions)
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k 'regex:^k_(fit|gate|gate_v4|gate_v2)$' -c 16 --csv --log-file $O/traffic_${TAG}.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > $O/ncu_traffic_${TAG}.log 2>&1
# 4. full-set captures of our kernels on the fixed small case
python profiles/profile_case.py 60000 1e-4 > $O/prof_${TAG}_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k 'regex:^k_(fit|gate|gate_v4|gate_v2|post|fit_heavy|pack)$' -c 5 -f \
    -o $O/prof_${TAG}_full python profiles/profile_case.py 60000 1e-4 > $O/ncu_full_${TAG}.log 2>&1
tail -c 400 $O/bench_${TAG}_full.json; echo; tail -2 $O/ncu_full_${TAG}.log
