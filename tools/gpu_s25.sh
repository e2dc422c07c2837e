#!/usr/bin/env bash
# round-2 visit 25: launch-list peepholes — bit-identity tests, the objective / model / WGAN-GP suites (rank-1 rewrite),
# step time per knob
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q --no-header -x -k "fusion or objectives or wgan or models or gmgan_step or deferred" 2>&1 | grep -v "^  File" | tail -60 | cut -c1-300 ) > gpurun_out/pytest_s25b.log
tail -30 gpurun_out/pytest_s25b.log
for v in "GG_X=0" "GG_FUSE_EW=0" "GG_FUSE_TRANSPOSE=0" "GG_TICK_FIRST=1" "GG_FUSE_EW=0 GG_FUSE_TRANSPOSE=0 GG_FUSE_ACTGRAD_DENSE=0 GG_RANK1_MUL=0 GG_GATHER=0 GG_TICK_FIRST=1"; do
  echo "== cifar $v" >> gpurun_out/quick_s25b.txt
  ( env $v timeout 200 python bench.py --quick --steps 40 --warmup 5 2>&1 | tail -1 | cut -c1-200 ) >> gpurun_out/quick_s25b.txt
done
for cfg in face ssgan; do
  echo "== $cfg" >> gpurun_out/quick_s25b.txt
  ( timeout 200 python bench.py --quick --config $cfg --steps 20 --warmup 5 2>&1 | tail -1 | cut -c1-200 ) >> gpurun_out/quick_s25b.txt
done
cat gpurun_out/quick_s25b.txt
