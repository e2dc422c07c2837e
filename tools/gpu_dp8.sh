#!/usr/bin/env bash
# 8-GPU visit (expensive: 8x charge): DP correctness at N=8, the weak-scaling bench line at N=8, and the two sharded BASELINE.json
# configs (face bs=128 over 8 GPUs, SSGAN bs=32 sequences over 4)
set -u
mkdir -p gpurun_out
TAG="${1:-dp8}"
N="${2:-8}"
trun() { timeout "$1" python -m torch.distributed.run --nnodes=1 --nproc-per-node "$2" --master-addr 127.0.0.1 --master-port "$3" "${@:4}"; }
trun 90 $N 29541 tools/dp_check.py --out /tmp/dpN.npz 2>&1 | grep -a "dp_check\|Error\|error\|worst" | tail -4
( trun 120 $N 29561 bench.py --gpus $N --steps 60 --warmup 5 2>&1 | tail -1 ) > gpurun_out/bench_cifar_${TAG}.json
python -c "import json;d=json.load(open('gpurun_out/bench_cifar_${TAG}.json'));print('cifar N=$N',d['value'],d['ms_per_step'],'e2e',d['e2e']['value'],d['config']['replica_checksum'])"
( trun 120 8 29562 bench.py --config face --gpus 8 --steps 40 --warmup 5 2>&1 | tail -1 ) > gpurun_out/bench_face_dp8_${TAG}.json
python -c "import json;d=json.load(open('gpurun_out/bench_face_dp8_${TAG}.json'));print('face N=8',d['value'],d['ms_per_step'],'e2e',d['e2e']['value'],d['config']['replica_checksum'])"
( trun 120 4 29563 bench.py --config ssgan --gpus 4 --steps 40 --warmup 5 2>&1 | tail -1 ) > gpurun_out/bench_ssgan_dp4_${TAG}.json
python -c "import json;d=json.load(open('gpurun_out/bench_ssgan_dp4_${TAG}.json'));print('ssgan N=4',d['value'],d['ms_per_step'],'e2e',d['e2e']['value'],d['config']['replica_checksum'])"
( trun 80 4 29564 bench.py --gpus 4 --quick --steps 40 --warmup 5 2>&1 | grep -a "quick" | cut -c1-200 | tail -1 ) > gpurun_out/quick4_${TAG}.txt
cat gpurun_out/quick4_${TAG}.txt
