#!/bin/bash
# minimal evidence run: bench (both arms), launch list, ncu capture of the rollout kernel at 65 536 envs
T=$1
python bench.py > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err
python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_ref_$T.json 2>> gpurun_out/bench_$T.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$T.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_$T.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rollout -s 3 -c 1 -f -o gpurun_out/prof_rollout_$T python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_rollout_$T.log 2>&1
ncu -i gpurun_out/prof_rollout_$T.ncu-rep --page raw --csv > gpurun_out/prof_rollout_$T.raw.csv 2>/dev/null
ncu -i gpurun_out/prof_rollout_$T.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/prof_rollout_$T.cuda.csv 2>/dev/null
tail -c 300 gpurun_out/bench_$T.json
