#!/usr/bin/env bash
# round-2 visit 31: element-wise program kernels without 64-bit divisions — bit-identity suite + step time
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_fusion.py -m gpu -q --no-header 2>&1 | grep -v "^  File" | tail -6 | cut -c1-250 ) > gpurun_out/pytest_s31.log
cat gpurun_out/pytest_s31.log
for cfg in cifar ssgan; do
  echo "== $cfg" >> gpurun_out/quick_s31.txt
  ( timeout 200 python bench.py --quick --config $cfg --steps 30 --warmup 5 2>&1 | tail -1 | cut -c1-200 ) >> gpurun_out/quick_s31.txt
done
cat gpurun_out/quick_s31.txt
