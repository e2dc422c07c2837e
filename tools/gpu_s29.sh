#!/usr/bin/env bash
# round-2 visit 29: batch-norm statistics in the epilogue of the producing tensor-core launch — parity + step time
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_fusion.py -m gpu -q --no-header -x -k "batchnorm or statistics or bit_identical" 2>&1 | grep -v "^  File" | tail -40 | cut -c1-300 ) > gpurun_out/pytest_s29.log
tail -40 gpurun_out/pytest_s29.log
for cfg in cifar face ssgan; do
  for v in 1 0; do
    echo "== $cfg GG_BN_CONV_STATS=$v" >> gpurun_out/quick_s29.txt
    ( env GG_BN_CONV_STATS=$v timeout 200 python bench.py --quick --config $cfg --steps 30 --warmup 5 2>&1 | tail -1 | cut -c1-200 ) >> gpurun_out/quick_s29.txt
  done
done
cat gpurun_out/quick_s29.txt
