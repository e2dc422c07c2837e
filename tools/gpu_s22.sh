#!/usr/bin/env bash
# round-2 visit 22: stream-count / priority variants with two-CTA co-residency working, then the round's evidence set
set -u
mkdir -p gpurun_out
TAG="${1:-s22}"
: > gpurun_out/quick_${TAG}.txt
for v in "GG_X=0" "GG_STREAMS=10" "GG_STREAMS=12" "GG_PRIO=0" "GG_SCHED=heft GG_PRIO=0" "GG_PRIO_SLACK_US=80"; do
  echo "== cifar $v" >> gpurun_out/quick_${TAG}.txt
  ( env $v timeout 300 python bench.py --quick --steps 30 --warmup 5 2>&1 | tail -1 | cut -c1-300 ) >> gpurun_out/quick_${TAG}.txt
done
cat gpurun_out/quick_${TAG}.txt
