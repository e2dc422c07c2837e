#!/usr/bin/env bash
# round-2 visit 4: conv kernel with deep ring in cluster mode + specialised reduce; ring-depth variants for the step;
# trajectory / family tests after the fixes
set -u
mkdir -p gpurun_out
TAG="${1:-s4}"
TL=graphical-gan_b200/lib/libgg_b200_tl.so
( timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_production_shapes.py tests/test_gpu_golden.py -m gpu -x -q 2>&1 | tail -30 | cut -c1-260 ) > gpurun_out/pytest_kern_${TAG}.log
tail -4 gpurun_out/pytest_kern_${TAG}.log
( GG_LIB=$TL timeout 120 python tools/timeline_conv.py 2>&1 | tail -40 ) > gpurun_out/timeline_${TAG}.txt
head -5 gpurun_out/timeline_${TAG}.txt | cut -c1-230
( timeout 300 python tools/time_conv.py batched 2>&1 | tail -30 ) > gpurun_out/time_conv_${TAG}.txt
cat gpurun_out/time_conv_${TAG}.txt
: > gpurun_out/quick_${TAG}.txt
for v in "GG_X=0" "GG_TC_STAGES=0" "GG_TC_STAGES=4" "GG_STREAMS=4" "GG_STREAMS=8"; do
  echo "== $v" >> gpurun_out/quick_${TAG}.txt
  ( env $v timeout 200 python bench.py --quick --steps 40 --warmup 5 2>&1 | tail -1 | cut -c1-300 ) >> gpurun_out/quick_${TAG}.txt
done
cat gpurun_out/quick_${TAG}.txt
( timeout 900 python -m pytest tests/test_gpu_objectives.py tests/test_gpu_trajectory100.py -m gpu -q --no-header -s -k "gmgan_inference_face or iteration_100" 2>&1 | tail -150 | cut -c1-300 ) > gpurun_out/pytest_fail_${TAG}.log
grep -n "kernels vs fp64" -A10 gpurun_out/pytest_fail_${TAG}.log | cut -c1-200
tail -5 gpurun_out/pytest_fail_${TAG}.log
python tools/profile_timeline.py gen > gpurun_out/timeline_gen_${TAG}.txt 2>&1
python tools/profile_timeline.py disc > gpurun_out/timeline_disc_${TAG}.txt 2>&1
head -4 gpurun_out/timeline_gen_${TAG}.txt; head -4 gpurun_out/timeline_disc_${TAG}.txt
