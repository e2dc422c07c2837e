#!/usr/bin/env bash
# 2-GPU visit: unordered one-launch SyncBN, DP step timelines (rank 0), exchange variants, clean exit of dp_check
set -u
mkdir -p gpurun_out
TAG="${1:-dp3}"
N="${2:-2}"
run() { timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
python tools/dp_check.py --out /tmp/single.npz 2>&1 | tail -1
( time run 29541 tools/dp_check.py --out /tmp/dpN.npz ) 2>&1 | grep -v Warning | tail -6
python - <<'PY'
import numpy as np
a,b=np.load('/tmp/single.npz'),np.load('/tmp/dpN.npz')
print('costs',a['costs'],b['costs'])
worst=max((np.abs(a[k]-b[k]).max(),k) for k in a.files if k!='costs')
print('worst param diff',worst)
PY
: > gpurun_out/quick_${TAG}.txt
port=29551
for v in "GG_X=0" "GG_BN_DP_ORDER=1" "GG_SYNC_BN=0" "GG_DP_BUCKETS=4" "GG_STREAMS=8"; do
  echo "== N=$N $v" >> gpurun_out/quick_${TAG}.txt
  port=$((port+1))
  ( env $v timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --quick --steps 40 --warmup 5 2>&1 | grep -a "quick\|Error\|error" | cut -c1-220 | tail -3 ) >> gpurun_out/quick_${TAG}.txt
done
cat gpurun_out/quick_${TAG}.txt
run 29561 tools/profile_timeline.py gen > gpurun_out/timeline_gen_${TAG}.txt 2>&1
run 29562 tools/profile_timeline.py disc > gpurun_out/timeline_disc_${TAG}.txt 2>&1
grep -a "step:" gpurun_out/timeline_gen_${TAG}.txt gpurun_out/timeline_disc_${TAG}.txt
