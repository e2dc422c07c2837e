#!/usr/bin/env bash
# 2-GPU check of the final defaults: DP parity test, dp_check, bench line
set -u
mkdir -p gpurun_out
TAG="${1:-dp2b}"
trun() { timeout "$1" python -m torch.distributed.run --nnodes=1 --nproc-per-node "$2" --master-addr 127.0.0.1 --master-port "$3" "${@:4}"; }
( timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -3 ) > gpurun_out/pytest_multi_${TAG}.log
cat gpurun_out/pytest_multi_${TAG}.log
trun 90 2 29571 tools/dp_check.py --out /tmp/dp2.npz 2>&1 | grep -a "dp_check\|Error\|error\|worst" | tail -3
( trun 150 2 29572 bench.py --gpus 2 --steps 60 --warmup 5 2>&1 | tail -1 ) > gpurun_out/bench_cifar_${TAG}.json
python -c "import json;d=json.load(open('gpurun_out/bench_cifar_${TAG}.json'));print('cifar N=2',d['value'],d['ms_per_step'],'e2e',d['e2e']['value'],d['config']['replica_checksum'])"
