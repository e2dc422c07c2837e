#!/usr/bin/env bash
# 2-GPU visit: dedicated collective stream, SyncBN kernel microbenchmark, DP timelines
set -u
mkdir -p gpurun_out
TAG="${1:-dp4}"
N="${2:-2}"
run() { timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
( run 29531 tools/time_bn_dp.py 2>&1 | grep -a "^bn" ) | tee gpurun_out/time_bn_dp_${TAG}.txt
: > gpurun_out/quick_${TAG}.txt
port=29551
for v in "GG_X=0" "GG_DP_BUCKETS=1" "GG_DP_BUCKETS=3" "GG_SYNC_BN=0"; do
  echo "== N=$N $v" >> gpurun_out/quick_${TAG}.txt
  port=$((port+1))
  ( env $v timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --quick --steps 40 --warmup 5 2>&1 | grep -a "quick\|Error\|error" | cut -c1-220 | tail -3 ) >> gpurun_out/quick_${TAG}.txt
done
cat gpurun_out/quick_${TAG}.txt
( time run 29561 tools/profile_timeline.py gen > gpurun_out/timeline_gen_${TAG}.txt 2>&1 ) 2>&1 | grep real
( time run 29562 tools/profile_timeline.py disc > gpurun_out/timeline_disc_${TAG}.txt 2>&1 ) 2>&1 | grep real
grep -a "step:" gpurun_out/timeline_gen_${TAG}.txt gpurun_out/timeline_disc_${TAG}.txt
grep -a "nccl\|adam_multi" gpurun_out/timeline_disc_${TAG}.txt | cut -c1-120 | tail -6
