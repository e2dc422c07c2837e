#!/usr/bin/env bash
# GPU visit 1 of the session: validate HEAD, record a bench line, and run the split-K / concurrency experiments.
set -u
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q --no-header --durations=12 2>&1 | tail -40 | cut -c1-220 ) > gpurun_out/pytest_s2a.log
tail -4 gpurun_out/pytest_s2a.log
timeout 300 python bench.py --steps 60 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_s2a.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_s2a.json'))
print('bench', round(d['value']), 'img/s', round(d['ms_per_step'],3), 'ms; e2e', round(d['e2e']['value']), 'launches/iter', d['gpu_launches']/d['steps'], 'kernel_ms', d['roofline']['kernel_ms'], 'cpu', d.get('cpu_baseline',{}).get('value'))
PY
for v in "GG_TC_CLUSTER=1" "GG_TC_SPLITS=1" "GG_STREAMS=1" "GG_STREAMS=3" "GG_TC_MAX_CTAS=148" "GG_TC_CLUSTER=1 GG_TC_MAX_CTAS=148" "GG_IM2COL=0"; do
  echo "== $v" >> gpurun_out/quick_s2a.txt
  ( env $v timeout 200 python bench.py --quick --steps 40 --warmup 5 2>&1 | tail -1 | cut -c1-400 ) >> gpurun_out/quick_s2a.txt
done
cat gpurun_out/quick_s2a.txt
for v in "GG_X=0" "GG_TC_CLUSTER=1" "GG_TC_SPLITS=1" "GG_TC_SPLITS=2"; do
  ( env $v timeout 120 python tools/exp_concurrency.py 2>&1 | tail -4 ) >> gpurun_out/concurrency_s2a.txt
done
cat gpurun_out/concurrency_s2a.txt
( GG_LIB=$PWD/graphical-gan_b200/lib/libgg_b200_tl.so timeout 120 python tools/timeline_conv.py 2>&1 | tail -40 | cut -c1-330 ) > gpurun_out/timeline_s2a.txt
cat gpurun_out/timeline_s2a.txt
timeout 200 python tools/time_conv.py 2>&1 | tail -22 > gpurun_out/time_conv_s2a.txt
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_s2a.csv python tools/profile_step.py 2>&1 | tail -1
ls -la gpurun_out | tail -12
