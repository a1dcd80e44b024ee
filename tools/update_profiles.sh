#!/bin/bash
# copies the evidence of gpurun_out/*_$1.* into profiles/ (tracked) under the round's names
T=$1
python profiles/ncu_summary.py gpurun_out/prof_rollout_$T.raw.csv > profiles/r1_rollout_ncu_summary.txt
python profiles/ncu_summary.py gpurun_out/prof_rollout1m_$T.raw.csv > profiles/r1_rollout_1m_envs_ncu_summary.txt
python profiles/ncu_summary.py gpurun_out/prof_trajgen_$T.raw.csv > profiles/r1_trajgen_ncu_summary.txt
python profiles/ncu_summary.py gpurun_out/prof_k_cov_simt_$T.raw.csv > profiles/r1_cov_simt_ncu_summary.txt
python profiles/ncu_summary.py gpurun_out/prof_k_cov_umma_$T.raw.csv > profiles/r1_cov_umma_ncu_summary.txt
python profiles/src_hot.py gpurun_out/prof_rollout_$T.cuda.csv 25 > profiles/r1_rollout_hot_lines.txt
python profiles/src_hot.py gpurun_out/prof_trajgen_$T.cuda.csv 20 > profiles/r1_trajgen_hot_lines.txt
cp gpurun_out/launches_$T.csv profiles/r1_launches.csv
cp gpurun_out/bench_$T.json profiles/r1_bench.json
cp gpurun_out/bench_ref_$T.json profiles/r1_bench_reference_arm.json
cp gpurun_out/bench_configs.json profiles/r1_bench_configs.json
ls -la profiles
