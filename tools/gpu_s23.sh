#!/usr/bin/env bash
# round-2 visit 23: programmatic dependent launch again, now that dependents can co-reside (named-barrier fix)
set -u
mkdir -p gpurun_out
TAG="${1:-s23}"
: > gpurun_out/quick_${TAG}.txt
for v in "GG_X=0" "GG_PDL=1" "GG_PDL=1 GG_STREAMS=6"; do
  echo "== cifar $v" >> gpurun_out/quick_${TAG}.txt
  ( env $v timeout 300 python bench.py --quick --steps 30 --warmup 5 2>&1 | tail -1 | cut -c1-300 ) >> gpurun_out/quick_${TAG}.txt
done
echo "== face GG_PDL=1" >> gpurun_out/quick_${TAG}.txt
( GG_PDL=1 timeout 300 python bench.py --config face --quick --steps 30 --warmup 5 2>&1 | tail -1 | cut -c1-300 ) >> gpurun_out/quick_${TAG}.txt
cat gpurun_out/quick_${TAG}.txt
( GG_PDL=1 timeout 600 python -m pytest tests/test_gpu_gmgan_step.py tests/test_gpu_deferred.py -m gpu -x -q 2>&1 | tail -3 ) > gpurun_out/pytest_pdl_${TAG}.log
cat gpurun_out/pytest_pdl_${TAG}.log
