#!/usr/bin/env bash
# second 8-GPU visit (8x charge): exchange variants of the weak-scaling step at N=8
set -u
mkdir -p gpurun_out
TAG="${1:-dp8b}"
N=8
trun() { timeout "$1" python -m torch.distributed.run --nnodes=1 --nproc-per-node "$2" --master-addr 127.0.0.1 --master-port "$3" "${@:4}"; }
: > gpurun_out/quick_${TAG}.txt
port=29600
for v in "GG_X=0" "GG_DP_LATE_BUCKET=1" "GG_DP_BUCKETS=2" "GG_DP_BUCKETS=8" "GG_SYNC_BN=0" "GG_DP_LATE_BUCKET=1 GG_DP_BUCKETS=2"; do
  port=$((port+1))
  echo "== N=$N $v" >> gpurun_out/quick_${TAG}.txt
  ( export $v; trun 90 $N $port bench.py --gpus $N --quick --steps 40 --warmup 5 2>&1 | grep -a "quick\|Error\|error" | cut -c1-230 | tail -2 ) >> gpurun_out/quick_${TAG}.txt
done
cat gpurun_out/quick_${TAG}.txt
