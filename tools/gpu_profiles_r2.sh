#!/bin/bash
# Round 2 evidence in one GPU session: bench (both arms), launch list of the bench command, one `ncu --set full` capture per
# kernel (summaries + hot lines + SASS opcode histogram).  Usage: bash tools/gpu_profiles_r2.sh <tag>; then tools/update_profiles_r2.sh <tag>
T=$1
cap () {   # cap <name> <kernel regex> <skip> <command...>
  local name=$1 rx=$2 skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o gpurun_out/prof_${name}_$T "$@" > gpurun_out/ncu_${name}_$T.log 2>&1
  ncu -i gpurun_out/prof_${name}_$T.ncu-rep --page raw --csv > gpurun_out/prof_${name}_$T.raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_${name}_$T.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/prof_${name}_$T.cuda.csv 2>/dev/null
  ncu -i gpurun_out/prof_${name}_$T.ncu-rep --page source --print-source sass --csv > gpurun_out/prof_${name}_$T.sass.csv 2>/dev/null
  rm -f gpurun_out/prof_${name}_$T.ncu-rep
}
python bench.py > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; tail -c 400 gpurun_out/bench_$T.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$T.json 2>> gpurun_out/bench_$T.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$T.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_$T.log 2>&1
cap rollout k_rollout 3 python bench.py --steps 3 --warmup 3 --no-cpu-baseline
cap rollout_sigma1 k_rollout 3 python tools/run_kernel.py rollout_sigma1
cap rollout_1m k_rollout 2 python tools/run_kernel.py rollout_1m
cap rollout_viapoint_dmp k_rollout 2 python tools/run_kernel.py rollout_config3
cap rollout_simple_prodmp_plans k_rollout 2 python tools/run_kernel.py rollout_config4_plans
cap rollout_learned_tau k_rollout 2 python tools/run_kernel.py rollout_learned_tau
cap rollout_1m_sigma1 k_rollout 2 python tools/run_kernel.py rollout_1m_sigma1
cap trajgen_promp k_trajgen_closed 3 python tools/run_kernel.py trajgen_promp
cap trajgen_prodmp k_trajgen_closed 3 python tools/run_kernel.py trajgen_prodmp
cap trajgen_dmp k_trajgen_dmp 3 python tools/run_kernel.py trajgen_dmp
cap trajgen_phase_promp k_trajgen_phase 3 python tools/run_kernel.py trajgen_phase_promp
cap dmp_integrate_phase k_dmp_integrate_phase 3 python tools/run_kernel.py trajgen_phase_dmp
cap reset k_reset 3 python tools/run_kernel.py reset
cap cov_simt k_cov_simt 4 python tools/probe_cov.py
cap cov_umma k_cov_umma 4 python tools/probe_cov.py
ls gpurun_out | grep _$T | wc -l
