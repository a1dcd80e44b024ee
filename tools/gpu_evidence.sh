#!/bin/bash
# evidence run: full GPU test suite, memcheck of the per-env-phase / ragged / pipeline paths, then tools/gpu_final.sh
T=$1
python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/gputests_$T.log; tail -2 gpurun_out/gputests_$T.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q -x -k "per_env or ragged or pipeline or sub_traj" > gpurun_out/memcheck_$T.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/memcheck_$T.log
bash tools/gpu_final.sh $T
