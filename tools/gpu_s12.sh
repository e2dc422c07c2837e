#!/usr/bin/env bash
# round-2 visit 12: GPU tests of the N2 / N4 additions (VEGAN no-discriminator modes, Conv3D, SSGAN ALI critics)
set -u
mkdir -p gpurun_out
TAG="${1:-s12}"
( timeout 900 python -m pytest tests/test_gpu_conv3d.py tests/test_gpu_objectives.py -m gpu -q --no-header 2>&1 | tail -30 | cut -c1-300 ) > gpurun_out/pytest_${TAG}.log
tail -25 gpurun_out/pytest_${TAG}.log
( timeout 200 python bench.py --quick --steps 40 --warmup 5 2>&1 | tail -1 | cut -c1-300 ) > gpurun_out/quick_${TAG}.txt
cat gpurun_out/quick_${TAG}.txt
