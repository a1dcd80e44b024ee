#!/bin/bash
# one `ncu --set full` capture + its csv pages:  bash tools/gpu_cap_one.sh <tag> <name> <kernel regex> <skip> <command...>
T=$1; name=$2; rx=$3; skip=$4; shift 4
ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o gpurun_out/prof_${name}_$T "$@" > gpurun_out/ncu_${name}_$T.log 2>&1
ncu -i gpurun_out/prof_${name}_$T.ncu-rep --page raw --csv > gpurun_out/prof_${name}_$T.raw.csv 2>/dev/null
ncu -i gpurun_out/prof_${name}_$T.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/prof_${name}_$T.cuda.csv 2>/dev/null
ncu -i gpurun_out/prof_${name}_$T.ncu-rep --page source --print-source sass --csv > gpurun_out/prof_${name}_$T.sass.csv 2>/dev/null
rm -f gpurun_out/prof_${name}_$T.ncu-rep
