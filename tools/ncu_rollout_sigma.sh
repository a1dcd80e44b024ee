ncu --set full --clock-control none --import-source on -k regex:k_rollout -s 3 -c 1 -f -o gpurun_out/prof_rollout_sigma1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --sigma 1.0 > gpurun_out/ncu_rollout_sigma1.log 2>&1
ncu -i gpurun_out/prof_rollout_sigma1.ncu-rep --page raw --csv > gpurun_out/prof_rollout_sigma1.raw.csv 2>/dev/null
ncu -i gpurun_out/prof_rollout_sigma1.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/prof_rollout_sigma1.cuda.csv 2>/dev/null
