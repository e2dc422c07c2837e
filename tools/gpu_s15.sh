#!/usr/bin/env bash
# round-2 visit 15: activation gradient fused into the dgrad write-out
set -u
mkdir -p gpurun_out
TAG="${1:-s15}"
( timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_gmgan_step.py tests/test_gpu_production_shapes.py -m gpu -x -q 2>&1 | tail -30 | cut -c1-260 ) > gpurun_out/pytest_kern_${TAG}.log
tail -5 gpurun_out/pytest_kern_${TAG}.log
: > gpurun_out/quick_${TAG}.txt
for v in "GG_X=0" "GG_FUSE_ACTGRAD=0"; do
  echo "== $v" >> gpurun_out/quick_${TAG}.txt
  ( env $v timeout 200 python bench.py --quick --steps 40 --warmup 5 2>&1 | tail -1 | cut -c1-300 ) >> gpurun_out/quick_${TAG}.txt
done
cat gpurun_out/quick_${TAG}.txt
( timeout 300 python tools/time_conv.py dom 2>&1 | tail -4 ) > gpurun_out/time_conv_${TAG}.txt
cat gpurun_out/time_conv_${TAG}.txt
