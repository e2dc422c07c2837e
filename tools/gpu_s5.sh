#!/usr/bin/env bash
# round-2 visit 5: dual TMA producers + per-chunk push; kernel tests, timeline, timings, step throughput, smoke
set -u
mkdir -p gpurun_out
TAG="${1:-s5}"
TL=graphical-gan_b200/lib/libgg_b200_tl.so
( timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_production_shapes.py tests/test_gpu_golden.py -m gpu -x -q 2>&1 | tail -30 | cut -c1-260 ) > gpurun_out/pytest_kern_${TAG}.log
tail -4 gpurun_out/pytest_kern_${TAG}.log
( GG_LIB=$TL timeout 120 python tools/timeline_conv.py 2>&1 | tail -40 ) > gpurun_out/timeline_${TAG}.txt
head -5 gpurun_out/timeline_${TAG}.txt | cut -c1-230
( timeout 300 python tools/time_conv.py batched 2>&1 | tail -30 ) > gpurun_out/time_conv_${TAG}.txt
cat gpurun_out/time_conv_${TAG}.txt
( timeout 200 python bench.py --quick --steps 40 --warmup 5 2>&1 | tail -1 | cut -c1-300 ) > gpurun_out/quick_${TAG}.txt
cat gpurun_out/quick_${TAG}.txt
( timeout 600 python -m pytest tests -m gpu -q --no-header -x 2>&1 | tail -12 | cut -c1-300 ) > gpurun_out/pytest_all_${TAG}.log
tail -6 gpurun_out/pytest_all_${TAG}.log
