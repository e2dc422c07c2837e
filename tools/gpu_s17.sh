#!/usr/bin/env bash
# round-2 visit 17: N-tower sibling batching (SSGAN)
set -u
mkdir -p gpurun_out
TAG="${1:-s17}"
( timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_objectives.py tests/test_gpu_gmgan_step.py -m gpu -x -q 2>&1 | tail -8 | cut -c1-260 ) > gpurun_out/pytest_${TAG}.log
tail -4 gpurun_out/pytest_${TAG}.log
: > gpurun_out/quick_${TAG}.txt
for v in "GG_X=0" "GG_BATCH_GROUPS=0"; do
  echo "== ssgan $v" >> gpurun_out/quick_${TAG}.txt
  ( env $v timeout 300 python bench.py --config ssgan --quick --steps 30 --warmup 5 2>&1 | tail -1 | cut -c1-300 ) >> gpurun_out/quick_${TAG}.txt
done
echo "== cifar" >> gpurun_out/quick_${TAG}.txt
( timeout 200 python bench.py --quick --steps 40 --warmup 5 2>&1 | tail -1 | cut -c1-300 ) >> gpurun_out/quick_${TAG}.txt
cat gpurun_out/quick_${TAG}.txt
