#!/usr/bin/env bash
# 8-GPU visit (expensive: 8x charge): DP correctness + the weak-scaling step under 3 variants
set -u
mkdir -p gpurun_out
TAG="${1:-dp8}"
N="${2:-8}"
run() { timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
run 29541 tools/dp_check.py --out /tmp/dpN.npz 2>&1 | grep -a "dp_check\|Error\|error" | tail -3
: > gpurun_out/quick_${TAG}.txt
port=29551
for v in "GG_X=0" "GG_DP_BUCKETS=8" "GG_SYNC_BN=0"; do
  echo "== N=$N $v" >> gpurun_out/quick_${TAG}.txt
  port=$((port+1))
  ( env $v timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --quick --steps 40 --warmup 5 2>&1 | grep -a "quick\|Error\|error" | cut -c1-220 | tail -3 ) >> gpurun_out/quick_${TAG}.txt
done
cat gpurun_out/quick_${TAG}.txt
