#!/usr/bin/env bash
# round-2 visit 7: register-cached BN, async-fill wgrad, deferred fetches (e2e)
set -u
mkdir -p gpurun_out
TAG="${1:-s7}"
( timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_production_shapes.py tests/test_gpu_golden.py tests/test_gpu_deferred.py -m gpu -x -q 2>&1 | tail -30 | cut -c1-260 ) > gpurun_out/pytest_kern_${TAG}.log
tail -4 gpurun_out/pytest_kern_${TAG}.log
( timeout 300 python tools/time_conv.py batched 2>&1 | tail -30 ) > gpurun_out/time_conv_${TAG}.txt
grep "3->64" gpurun_out/time_conv_${TAG}.txt
( timeout 300 python tools/time_small.py 2>&1 | tail -30 ) > gpurun_out/time_small_${TAG}.txt
cat gpurun_out/time_small_${TAG}.txt
( GG_BN_CACHED=0 timeout 300 python tools/time_small.py 2>&1 | grep -i "bn" | tail -12 ) > gpurun_out/time_small_nocache_${TAG}.txt
cat gpurun_out/time_small_nocache_${TAG}.txt
: > gpurun_out/quick_${TAG}.txt
for v in "GG_X=0" "GG_BN_CACHED=0"; do
  echo "== $v" >> gpurun_out/quick_${TAG}.txt
  ( env $v timeout 200 python bench.py --quick --steps 40 --warmup 5 2>&1 | tail -1 | cut -c1-300 ) >> gpurun_out/quick_${TAG}.txt
done
cat gpurun_out/quick_${TAG}.txt
( timeout 400 python bench.py --steps 100 --warmup 10 2>&1 | tail -1 ) > gpurun_out/bench_${TAG}.json
python -c "import json;d=json.load(open('gpurun_out/bench_${TAG}.json'));print('bench',d['value'],d['ms_per_step'],'e2e',d['e2e'],'launches/iter',d['gpu_launches']/d['steps'],'frac',d['roofline']['frac'])"
( GG_E2E_SYNC=1 timeout 400 python bench.py --steps 100 --warmup 10 2>&1 | tail -1 ) > gpurun_out/bench_sync_${TAG}.json
python -c "import json;d=json.load(open('gpurun_out/bench_sync_${TAG}.json'));print('bench sync-e2e',d['value'],'e2e',d['e2e']['value'])"
( timeout 600 python -m pytest tests -m gpu -q --no-header -x 2>&1 | tail -12 | cut -c1-300 ) > gpurun_out/pytest_all_${TAG}.log
tail -4 gpurun_out/pytest_all_${TAG}.log
