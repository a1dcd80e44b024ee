#!/bin/bash
# copies the evidence of gpurun_out/*_$1.* into profiles/ (tracked) under the round's names and rewrites profiles/ncu_numbers.json
T=$1
for k in rollout rollout_sigma1 rollout_1m rollout_1m_sigma1 rollout_learned_tau rollout_viapoint_dmp rollout_simple_prodmp_plans trajgen_promp trajgen_prodmp trajgen_dmp trajgen_phase_promp dmp_integrate_phase reset cov_simt cov_umma; do
  out=$k; [ $k = rollout_1m ] && out=rollout_1m_envs
  [ -s gpurun_out/prof_${k}_$T.raw.csv ] && python profiles/ncu_summary.py gpurun_out/prof_${k}_$T.raw.csv > profiles/r2_${out}_ncu_summary.txt
  [ -s gpurun_out/prof_${k}_$T.cuda.csv ] && python profiles/src_hot.py gpurun_out/prof_${k}_$T.cuda.csv 25 > profiles/r2_${out}_hot_lines.txt
  [ -s gpurun_out/prof_${k}_$T.sass.csv ] && python profiles/sass_hist.py gpurun_out/prof_${k}_$T.sass.csv > profiles/r2_${out}_sass.txt
  case $k in rollout*) [ -s gpurun_out/prof_${k}_$T.cuda.csv ] && python profiles/lane_loss.py gpurun_out/prof_${k}_$T.cuda.csv 20 > profiles/r2_${out}_lane_loss.txt;; esac
done
cp gpurun_out/launches_$T.csv profiles/r2_launches.csv
cp gpurun_out/bench_$T.json profiles/r2_bench.json
cp gpurun_out/bench_ref_$T.json profiles/r2_bench_reference_arm.json
python tools/ncu_to_json.py
ls profiles | head -80
