#!/usr/bin/env bash
# round-2 visit 10: ASAP issue order, split optimiser update
set -u
mkdir -p gpurun_out
TAG="${1:-s10}"
: > gpurun_out/quick_${TAG}.txt
for v in "GG_X=0" "GG_SCHED=heft" "GG_SPLIT_UPDATE=0" "GG_SCHED=heft GG_SPLIT_UPDATE=0" "GG_STREAMS=8" "GG_STREAMS=4"; do
  echo "== $v" >> gpurun_out/quick_${TAG}.txt
  ( env $v timeout 200 python bench.py --quick --steps 40 --warmup 5 2>&1 | tail -1 | cut -c1-300 ) >> gpurun_out/quick_${TAG}.txt
done
cat gpurun_out/quick_${TAG}.txt
python tools/profile_timeline.py gen > gpurun_out/timeline_gen_${TAG}.txt 2>&1
python tools/profile_timeline.py disc > gpurun_out/timeline_disc_${TAG}.txt 2>&1
head -3 gpurun_out/timeline_gen_${TAG}.txt | tail -2; head -3 gpurun_out/timeline_disc_${TAG}.txt | tail -2
( timeout 900 python -m pytest tests -m gpu -q --no-header -x 2>&1 | tail -12 | cut -c1-300 ) > gpurun_out/pytest_all_${TAG}.log
tail -4 gpurun_out/pytest_all_${TAG}.log
