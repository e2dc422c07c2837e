#!/usr/bin/env bash
# round-2 visit 27: priority classes without inversion (GG_PRIO_TAIL) — step time per variant, CUPTI timeline of the G step
mkdir -p gpurun_out
for v in "GG_PRIO_TAIL=1" "GG_PRIO_TAIL=0" "GG_PRIO_TAIL=1 GG_PRIO_SLACK_US=80" "GG_PRIO_TAIL=1 GG_STREAMS=10"; do
  echo "== cifar $v" >> gpurun_out/quick_s27.txt
  ( env $v timeout 200 python bench.py --quick --steps 40 --warmup 5 2>&1 | tail -1 | cut -c1-200 ) >> gpurun_out/quick_s27.txt
done
for cfg in face ssgan; do
  for v in "GG_PRIO_TAIL=1" "GG_PRIO_TAIL=0"; do
    echo "== $cfg $v" >> gpurun_out/quick_s27.txt
    ( env $v timeout 200 python bench.py --quick --config $cfg --steps 20 --warmup 5 2>&1 | tail -1 | cut -c1-200 ) >> gpurun_out/quick_s27.txt
  done
done
cat gpurun_out/quick_s27.txt
python tools/profile_timeline.py gen > gpurun_out/timeline_gen_s27.txt 2>&1
python tools/profile_timeline.py disc > gpurun_out/timeline_disc_s27.txt 2>&1
grep -E "step:|in flight" gpurun_out/timeline_gen_s27.txt gpurun_out/timeline_disc_s27.txt
