#!/bin/bash
# ncu --set full captures of the two covariance paths (CUDA cores / tcgen05), one launch each
for k in k_cov_simt k_cov_umma; do
ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/prof_${k}_$1 python tools/probe_cov.py > gpurun_out/ncu_${k}_$1.log 2>&1
ncu -i gpurun_out/prof_${k}_$1.ncu-rep --page raw --csv > gpurun_out/prof_${k}_$1.raw.csv 2>/dev/null
done
tail -2 gpurun_out/ncu_k_cov_umma_$1.log
