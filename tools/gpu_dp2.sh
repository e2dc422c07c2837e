#!/usr/bin/env bash
# 2-GPU visit (round 2): DP correctness of the new exchange (one-launch SyncBN over the peer arena, in-graph bucketed NCCL)
# and the weak-scaling step under the exchange variants
set -u
mkdir -p gpurun_out
TAG="${1:-dp2}"
N="${2:-2}"
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
python tools/dp_check.py --out /tmp/single.npz 2>&1 | tail -1
run 29541 tools/dp_check.py --out /tmp/dpN.npz 2>&1 | grep -v Warning | tail -4
python - <<'PY'
import numpy as np
a,b=np.load('/tmp/single.npz'),np.load('/tmp/dpN.npz')
print('costs',a['costs'],b['costs'])
worst=max((np.abs(a[k]-b[k]).max(),k) for k in a.files if k!='costs')
print('worst param diff',worst)
PY
( timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --no-header 2>&1 | tail -5 ) | tee gpurun_out/pytest_multi_${TAG}.log
: > gpurun_out/quick_${TAG}.txt
echo "== single GPU" >> gpurun_out/quick_${TAG}.txt
( timeout 200 python bench.py --quick --steps 40 --warmup 5 2>&1 | grep quick | cut -c1-220 ) >> gpurun_out/quick_${TAG}.txt
port=29551
for v in "GG_X=0" "GG_BN_DP=0" "GG_DP_BUCKETS=1" "GG_DP_BUCKETS=3" "GG_NCCL_IN_GRAPH=0" "GG_SYNC_BN=0" "GG_DP_DIRECT=0"; do
  echo "== N=$N $v" >> gpurun_out/quick_${TAG}.txt
  port=$((port+1))
  ( env $v timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --quick --steps 40 --warmup 5 2>&1 | grep -a "quick\|Error\|error" | cut -c1-220 | tail -3 ) >> gpurun_out/quick_${TAG}.txt
done
cat gpurun_out/quick_${TAG}.txt
( run 29571 bench.py --gpus $N --steps 50 --warmup 5 2>&1 | grep -a '"metric"' | tail -1 ) > gpurun_out/bench_${TAG}.json
python -c "
import json;d=json.load(open('gpurun_out/bench_${TAG}.json'));print('bench N',d['n_gpus'],d['value'],d['ms_per_step'],'e2e',d['e2e']['value'],'checksum',d['config']['replica_checksum'])"
