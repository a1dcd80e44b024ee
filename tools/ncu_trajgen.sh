#!/bin/bash
# one --set full capture of the trajectory-only kernel (the probe launches it ~25 times; capture launch #5)
ncu --set full --clock-control none --import-source on -k regex:k_trajgen -s 5 -c 1 -f -o gpurun_out/prof_trajgen_$1 python tools/probe_trajgen.py > gpurun_out/ncu_trajgen_$1.log 2>&1
ncu -i gpurun_out/prof_trajgen_$1.ncu-rep --page raw --csv > gpurun_out/prof_trajgen_$1.raw.csv 2>/dev/null
ncu -i gpurun_out/prof_trajgen_$1.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/prof_trajgen_$1.cuda.csv 2>/dev/null
tail -3 gpurun_out/ncu_trajgen_$1.log
