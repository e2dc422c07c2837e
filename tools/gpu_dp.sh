#!/usr/bin/env bash
# 2-GPU visit: DP correctness (tools/dp_check.py single vs 2 ranks) and the weak-scaling bench, eager NCCL vs NCCL-in-graph
set -u
mkdir -p gpurun_out
run2() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
python tools/dp_check.py --out /tmp/single.npz 2>&1 | tail -1
GG_NCCL_IN_GRAPH=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/dp_check.py --out /tmp/dp2.npz 2>&1 | grep -v Warning | tail -3
python - <<'PY'
import numpy as np
a,b=np.load('/tmp/single.npz'),np.load('/tmp/dp2.npz')
print('costs',a['costs'],b['costs'])
worst=max((np.abs(a[k]-b[k]).max(),k) for k in a.files if k!='costs')
print('worst param diff',worst)
PY
for v in "GG_NCCL_IN_GRAPH=1" "GG_X=0" "GG_NCCL_IN_GRAPH=1 GG_SMALL_ALLREDUCE=0" "GG_SYNC_BN=0 GG_NCCL_IN_GRAPH=1"; do
  echo "== $v" >> gpurun_out/quick_dp.txt
  ( env $v timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --quick --steps 40 --warmup 5 2>&1 | grep quick | cut -c1-200 ) >> gpurun_out/quick_dp.txt
done
cat gpurun_out/quick_dp.txt
