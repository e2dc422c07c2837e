#!/usr/bin/env bash
# usage: tools/gpu_quick.sh "<pytest -k expr or empty for all>" [tag] ["ENV=.. ENV=.." variants separated by ';']
set -u
K="${1:-}"; TAG="${2:-q}"; VARS="${3:-GG_X=0}"
mkdir -p gpurun_out
if [ -n "$K" ]; then
  ( timeout 600 python -m pytest tests -m gpu -q --no-header -k "$K" 2>&1 | tail -40 | cut -c1-250 ) > gpurun_out/pytest_${TAG}.log
else
  ( timeout 900 python -m pytest tests -m gpu -q --no-header 2>&1 | tail -40 | cut -c1-250 ) > gpurun_out/pytest_${TAG}.log
fi
tail -25 gpurun_out/pytest_${TAG}.log
IFS=';' read -ra VS <<< "$VARS"
for v in "${VS[@]}"; do
  echo "== $v" >> gpurun_out/quick_${TAG}.txt
  ( env $v timeout 200 python bench.py --quick --steps 40 --warmup 5 2>&1 | tail -1 | cut -c1-300 ) >> gpurun_out/quick_${TAG}.txt
done
cat gpurun_out/quick_${TAG}.txt
