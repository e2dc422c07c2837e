#!/usr/bin/env bash
# Final visit of round 2: the whole GPU suite, the bench lines of the three workloads + the reference arm, the ncu launch list of
# one iteration, a full ncu capture of the kernels added in the second half of the round, smoke(), CUPTI timelines.
set -u
tag="${1:-r2f}"
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q --no-header 2>&1 | grep -v "^  File" | tail -25 | cut -c1-250 ) > gpurun_out/pytest_${tag}.log
tail -3 gpurun_out/pytest_${tag}.log
( timeout 400 python bench.py --steps 100 --warmup 10 2>&1 | tail -1 ) > gpurun_out/bench_${tag}.json
python -c "import json;d=json.load(open('gpurun_out/bench_${tag}.json'));print('bench',d['value'],d['ms_per_step'],'e2e',d['e2e']['value'],'launches/iter',d['gpu_launches']/d['steps'],'kernel_ms',d['roofline']['kernel_ms'],'frac',d['roofline']['frac'],'cpu',d.get('cpu_baseline',{}).get('value'))"
for cfg in face ssgan; do
  ( timeout 400 python bench.py --config $cfg --steps 50 --warmup 5 2>&1 | tail -1 ) > gpurun_out/bench_${cfg}_${tag}.json
  python -c "import json;d=json.load(open('gpurun_out/bench_${cfg}_${tag}.json'));print('bench $cfg',d['value'],d['ms_per_step'],'e2e',d['e2e']['value'],'launches/iter',d['gpu_launches']/d['steps'],'frac',d['roofline']['frac'])"
done
( timeout 400 python bench.py --impl reference --steps 10 --warmup 2 2>&1 | tail -1 ) > gpurun_out/bench_ref_${tag}.json
GG_CUDA_GRAPH=1 timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${tag}.csv python tools/profile_step.py 2>&1 | tail -1
grep -c "conv_tc\|ew_program" gpurun_out/launches_${tag}.csv
GG_CUDA_GRAPH=1 timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"ew_program|transpose_b2d_ex|bn_apply|gather_rows" -c 24 -o gpurun_out/prof_glue_${tag} python tools/profile_step.py 2>&1 | tail -1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python tools/profile_timeline.py gen > gpurun_out/timeline_gen_${tag}.txt 2>&1
python tools/profile_timeline.py disc > gpurun_out/timeline_disc_${tag}.txt 2>&1
grep -E "step:|in flight" gpurun_out/timeline_gen_${tag}.txt gpurun_out/timeline_disc_${tag}.txt
ls -la gpurun_out | tail -12
