#!/usr/bin/env bash
# N-GPU visit: LL-protocol SyncBN exchange — microbenchmark, correctness (dp_check), step variants
set -u
mkdir -p gpurun_out
TAG="${1:-dp5}"
N="${2:-2}"
run() { timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
( run 29531 tools/time_bn_dp.py 2>&1 | grep -a "^bn" ) | tee gpurun_out/time_bn_dp_${TAG}.txt
python tools/dp_check.py --out /tmp/single.npz 2>&1 | tail -1
run 29541 tools/dp_check.py --out /tmp/dpN.npz 2>&1 | grep -a "dp_check\|Error\|error" | tail -3
python - <<'PY'
import numpy as np
a,b=np.load('/tmp/single.npz'),np.load('/tmp/dpN.npz')
print('costs',a['costs'],b['costs'])
worst=max((np.abs(a[k]-b[k]).max(),k) for k in a.files if k!='costs')
print('worst param diff',worst)
PY
: > gpurun_out/quick_${TAG}.txt
port=29551
for v in "GG_X=0" "GG_DP_BUCKETS=3" "GG_DP_BUCKETS=4" "GG_TC_MAX_CTAS=132" "GG_TC_MAX_CTAS=132 GG_DP_BUCKETS=3" "GG_SYNC_BN=0"; do
  echo "== N=$N $v" >> gpurun_out/quick_${TAG}.txt
  port=$((port+1))
  ( env $v timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --quick --steps 40 --warmup 5 2>&1 | grep -a "quick\|Error\|error" | cut -c1-220 | tail -3 ) >> gpurun_out/quick_${TAG}.txt
done
cat gpurun_out/quick_${TAG}.txt
