#!/usr/bin/env bash
# round-2 visit 8: where the small-channel wgrad launch spends its time; deferred-fetch tests; plan-order launch list
set -u
mkdir -p gpurun_out
TAG="${1:-s8}"
( timeout 300 python tools/timeline_small_wgrad.py 2>&1 | tail -40 ) > gpurun_out/timeline_small_wgrad_${TAG}.txt
cat gpurun_out/timeline_small_wgrad_${TAG}.txt
( timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_deferred.py -m gpu -x -q 2>&1 | tail -30 | cut -c1-260 ) > gpurun_out/pytest_kern_${TAG}.log
tail -4 gpurun_out/pytest_kern_${TAG}.log
( timeout 300 python tools/time_small.py 2>&1 | grep "bn R" ) > gpurun_out/time_small_${TAG}.txt
cat gpurun_out/time_small_${TAG}.txt
GG_CUDA_GRAPH=0 timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_eager_${TAG}.csv python tools/profile_step.py 2>&1 | tail -1
( timeout 200 python bench.py --quick --steps 40 --warmup 5 2>&1 | tail -1 | cut -c1-300 ) > gpurun_out/quick_${TAG}.txt
cat gpurun_out/quick_${TAG}.txt
