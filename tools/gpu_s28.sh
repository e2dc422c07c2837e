#!/usr/bin/env bash
# round-2 visit 28: GG_PRIO_TAIL=2 (only cheap glue leaves the low class) on the three workloads
mkdir -p gpurun_out
for cfg in cifar face ssgan; do
  echo "== $cfg GG_PRIO_TAIL=2" >> gpurun_out/quick_s28.txt
  ( env GG_PRIO_TAIL=2 timeout 200 python bench.py --quick --config $cfg --steps 30 --warmup 5 2>&1 | tail -1 | cut -c1-200 ) >> gpurun_out/quick_s28.txt
done
cat gpurun_out/quick_s28.txt
