#!/usr/bin/env bash
# round-2 visit 13: persistent small-channel forward
set -u
mkdir -p gpurun_out
TAG="${1:-s13}"
( timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_production_shapes.py tests/test_gpu_golden.py -m gpu -x -q 2>&1 | tail -30 | cut -c1-260 ) > gpurun_out/pytest_kern_${TAG}.log
tail -4 gpurun_out/pytest_kern_${TAG}.log
( timeout 300 python tools/time_conv.py batched 2>&1 | grep "3->64" ) > gpurun_out/time_conv_${TAG}.txt
cat gpurun_out/time_conv_${TAG}.txt
cat gpurun_out/time_conv_nopersist_${TAG}.txt
: > gpurun_out/quick_${TAG}.txt
  echo "== $v" >> gpurun_out/quick_${TAG}.txt
  ( env $v timeout 200 python bench.py --quick --steps 40 --warmup 5 2>&1 | tail -1 | cut -c1-300 ) >> gpurun_out/quick_${TAG}.txt
done
cat gpurun_out/quick_${TAG}.txt
