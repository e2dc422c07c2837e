#!/usr/bin/env bash
# One GPU-box visit that produces everything a round needs: test results, the bench lines (ours, all three workloads, + the
# reference arm), the ncu launch list of one iteration, a full ncu capture of the tensor-core conv kernels (fwd / dgrad /
# wgrad), CUPTI timelines of both steps.  Outputs -> gpurun_out/.
set -u
tag="${1:-r2}"
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q --no-header 2>&1 | tail -25 | cut -c1-250 ) > gpurun_out/pytest_${tag}.log
tail -3 gpurun_out/pytest_${tag}.log
( timeout 400 python bench.py --steps 100 --warmup 10 2>&1 | tail -1 ) > gpurun_out/bench_${tag}.json
python -c "import json;d=json.load(open('gpurun_out/bench_${tag}.json'));print('bench',d['value'],d['ms_per_step'],'e2e',d['e2e']['value'],'launches/iter',d['gpu_launches']/d['steps'],'kernel_ms',d['roofline']['kernel_ms'],'frac',d['roofline']['frac'],'cpu',d.get('cpu_baseline',{}).get('value'))"
for cfg in face ssgan; do
  ( timeout 400 python bench.py --config $cfg --steps 50 --warmup 5 2>&1 | tail -1 ) > gpurun_out/bench_${cfg}_${tag}.json
  python -c "import json;d=json.load(open('gpurun_out/bench_${cfg}_${tag}.json'));print('bench $cfg',d['value'],d['ms_per_step'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'])"
done
( timeout 400 python bench.py --impl reference --steps 10 --warmup 2 2>&1 | tail -1 ) > gpurun_out/bench_ref_${tag}.json
GG_CUDA_GRAPH=1 timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${tag}.csv python tools/profile_step.py 2>&1 | tail -1
grep -c conv_tc gpurun_out/launches_${tag}.csv
GG_CUDA_GRAPH=1 timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_tc -c 40 -o gpurun_out/prof_conv_${tag} python tools/profile_step.py 2>&1 | tail -1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python tools/profile_timeline.py gen > gpurun_out/timeline_gen_${tag}.txt 2>&1
python tools/profile_timeline.py disc > gpurun_out/timeline_disc_${tag}.txt 2>&1
head -3 gpurun_out/timeline_gen_${tag}.txt; head -3 gpurun_out/timeline_disc_${tag}.txt
( timeout 300 python tools/time_conv.py batched 2>&1 | tail -30 ) > gpurun_out/time_conv_${tag}.txt
ls -la gpurun_out | tail -14
