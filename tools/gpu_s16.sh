#!/usr/bin/env bash
# round-2 visit 16: thin GEMM dispatch fix; the step's graph-structure floor (empty kernels)
set -u
mkdir -p gpurun_out
TAG="${1:-s16}"
( timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "gemm" 2>&1 | tail -5 | cut -c1-260 ) > gpurun_out/pytest_kern_${TAG}.log
tail -3 gpurun_out/pytest_kern_${TAG}.log
( timeout 300 python tools/exp_null_step.py 2>&1 | tail -4 ) > gpurun_out/null_step_${TAG}.txt
cat gpurun_out/null_step_${TAG}.txt
( GG_STREAMS=1 timeout 300 python tools/exp_null_step.py 2>&1 | tail -3 ) > gpurun_out/null_step_1stream_${TAG}.txt
cat gpurun_out/null_step_1stream_${TAG}.txt
( timeout 200 python bench.py --quick --steps 40 --warmup 5 2>&1 | tail -1 | cut -c1-300 ) > gpurun_out/quick_${TAG}.txt
cat gpurun_out/quick_${TAG}.txt
