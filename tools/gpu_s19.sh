#!/usr/bin/env bash
# round-2 visit 19: named-barrier fix (two CTAs per SM for the un-split tensor-core launches)
set -u
mkdir -p gpurun_out
TAG="${1:-s19}"
( timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_production_shapes.py tests/test_gpu_golden.py tests/test_gpu_gmgan_step.py -m gpu -x -q 2>&1 | tail -8 | cut -c1-260 ) > gpurun_out/pytest_kern_${TAG}.log
tail -4 gpurun_out/pytest_kern_${TAG}.log
TL=graphical-gan_b200/lib/libgg_b200_tl.so
( GG_LIB=$TL timeout 200 python tools/timeline_conv.py "face D.2 fwd" "face D.2 dgrad" 2>&1 | tail -12 ) > gpurun_out/timeline_face_${TAG}.txt
cat gpurun_out/timeline_face_${TAG}.txt | cut -c1-250
( timeout 300 python tools/time_conv.py batched 2>&1 | tail -16 ) > gpurun_out/time_conv_${TAG}.txt
cat gpurun_out/time_conv_${TAG}.txt
: > gpurun_out/quick_${TAG}.txt
for cfg in cifar face ssgan; do
  echo "== $cfg" >> gpurun_out/quick_${TAG}.txt
  ( timeout 300 python bench.py --config $cfg --quick --steps 30 --warmup 5 2>&1 | tail -1 | cut -c1-300 ) >> gpurun_out/quick_${TAG}.txt
done
echo "== cifar GG_TC_STAGES=4" >> gpurun_out/quick_${TAG}.txt
( GG_TC_STAGES=4 timeout 300 python bench.py --quick --steps 30 --warmup 5 2>&1 | tail -1 | cut -c1-300 ) >> gpurun_out/quick_${TAG}.txt
cat gpurun_out/quick_${TAG}.txt
