T=r2c
cap () {
  local name=$1 rx=$2 skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o gpurun_out/prof_${name}_$T "$@" > gpurun_out/ncu_${name}_$T.log 2>&1
  ncu -i gpurun_out/prof_${name}_$T.ncu-rep --page raw --csv > gpurun_out/prof_${name}_$T.raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_${name}_$T.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/prof_${name}_$T.cuda.csv 2>/dev/null
  ncu -i gpurun_out/prof_${name}_$T.ncu-rep --page source --print-source sass --csv > gpurun_out/prof_${name}_$T.sass.csv 2>/dev/null
  rm -f gpurun_out/prof_${name}_$T.ncu-rep
}
cap rollout_sigma1 k_rollout 3 python tools/run_kernel.py rollout_sigma1
cap rollout_1m_sigma1 k_rollout 2 python tools/run_kernel.py rollout_1m_sigma1
cap rollout k_rollout 3 python bench.py --steps 3 --warmup 3 --no-cpu-baseline
ls gpurun_out | grep _$T
