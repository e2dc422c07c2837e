#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q --no-header --durations=8 2>&1 | tail -60 | cut -c1-220 ) > gpurun_out/pytest_s2b.log
tail -15 gpurun_out/pytest_s2b.log
for v in "GG_STREAMS=8" "GG_STREAMS=12" "GG_STREAMS=8 GG_TC_MAX_CTAS=48" "GG_STREAMS=6 GG_TC_MAX_CTAS=100"; do
  echo "== $v" >> gpurun_out/quick_s2b.txt
  ( env $v timeout 200 python bench.py --quick --steps 40 --warmup 5 2>&1 | tail -1 | cut -c1-300 ) >> gpurun_out/quick_s2b.txt
done
cat gpurun_out/quick_s2b.txt
