#!/bin/bash
# one GPU session: tests, bench (both arms), launch list, full ncu capture of the rollout kernel
set -x
python bench.py > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err; tail -c 600 gpurun_out/bench_$1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$1.json 2>> gpurun_out/bench_$1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$1.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_$1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rollout -s 3 -c 1 -f -o gpurun_out/prof_rollout_$1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_rollout_$1.log 2>&1
ncu -i gpurun_out/prof_rollout_$1.ncu-rep --page raw --csv > gpurun_out/prof_rollout_$1.raw.csv 2>/dev/null
ncu -i gpurun_out/prof_rollout_$1.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/prof_rollout_$1.cuda.csv 2>/dev/null
bash tools/ncu_trajgen.sh $1
