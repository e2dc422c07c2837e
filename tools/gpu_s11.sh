#!/usr/bin/env bash
# round-2 visit 11: stream priority classes (critical chain high, slack work low)
set -u
mkdir -p gpurun_out
TAG="${1:-s11}"
: > gpurun_out/quick_${TAG}.txt
for v in "GG_X=0" "GG_PRIO=0" "GG_SCHED=heft" "GG_PRIO=0 GG_SCHED=heft" "GG_PRIO_SLACK_US=15" "GG_PRIO_SLACK_US=100" "GG_STREAMS=8"; do
  echo "== $v" >> gpurun_out/quick_${TAG}.txt
  ( env $v timeout 200 python bench.py --quick --steps 40 --warmup 5 2>&1 | tail -1 | cut -c1-300 ) >> gpurun_out/quick_${TAG}.txt
done
cat gpurun_out/quick_${TAG}.txt
python tools/profile_timeline.py gen > gpurun_out/timeline_gen_${TAG}.txt 2>&1
python tools/profile_timeline.py disc > gpurun_out/timeline_disc_${TAG}.txt 2>&1
head -3 gpurun_out/timeline_gen_${TAG}.txt | tail -2; head -3 gpurun_out/timeline_disc_${TAG}.txt | tail -2
( timeout 600 python -m pytest tests/test_gpu_gmgan_step.py tests/test_gpu_deferred.py -m gpu -q --no-header -x 2>&1 | tail -5 | cut -c1-300 ) > gpurun_out/pytest_${TAG}.log
tail -3 gpurun_out/pytest_${TAG}.log
