#!/usr/bin/env bash
# round-2 visit 21: ring depth of un-split launches = deepest that keeps two CTAs per SM
set -u
mkdir -p gpurun_out
TAG="${1:-s21}"
( timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_production_shapes.py -m gpu -x -q 2>&1 | tail -5 | cut -c1-260 ) > gpurun_out/pytest_kern_${TAG}.log
tail -3 gpurun_out/pytest_kern_${TAG}.log
: > gpurun_out/quick_${TAG}.txt
for cfg in cifar face ssgan; do
  echo "== $cfg" >> gpurun_out/quick_${TAG}.txt
  ( timeout 300 python bench.py --config $cfg --quick --steps 30 --warmup 5 2>&1 | tail -1 | cut -c1-300 ) >> gpurun_out/quick_${TAG}.txt
done
for v in "GG_STREAMS=8" "GG_TC_STAGES=4" "GG_TC_STAGES=0"; do
  echo "== cifar $v" >> gpurun_out/quick_${TAG}.txt
  ( env $v timeout 300 python bench.py --quick --steps 30 --warmup 5 2>&1 | tail -1 | cut -c1-300 ) >> gpurun_out/quick_${TAG}.txt
done
cat gpurun_out/quick_${TAG}.txt
