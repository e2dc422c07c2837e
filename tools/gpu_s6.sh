#!/usr/bin/env bash
# round-2 visit 6: one-launch small-channel wgrad, small-channel forward V2, CTA-cap sweep, tensor-pipe metric calibration
set -u
mkdir -p gpurun_out
TAG="${1:-s6}"
( timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_production_shapes.py tests/test_gpu_golden.py -m gpu -x -q 2>&1 | tail -30 | cut -c1-260 ) > gpurun_out/pytest_kern_${TAG}.log
tail -4 gpurun_out/pytest_kern_${TAG}.log
( GG_CONV_SMALL_V2=1 timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_golden.py -m gpu -x -q -k "conv or golden" 2>&1 | tail -8 | cut -c1-260 ) > gpurun_out/pytest_v2_${TAG}.log
tail -3 gpurun_out/pytest_v2_${TAG}.log
( timeout 300 python tools/time_conv.py batched 2>&1 | tail -30 ) > gpurun_out/time_conv_${TAG}.txt
grep "3->64" gpurun_out/time_conv_${TAG}.txt
( GG_CONV_SMALL_V2=1 timeout 300 python tools/time_conv.py batched 2>&1 | grep "3->64" ) > gpurun_out/time_conv_v2_${TAG}.txt
cat gpurun_out/time_conv_v2_${TAG}.txt
: > gpurun_out/quick_${TAG}.txt
for v in "GG_X=0" "GG_CONV_SMALL_V2=1" "GG_TC_MAX_CTAS=112" "GG_TC_MAX_CTAS=74"; do
  echo "== $v" >> gpurun_out/quick_${TAG}.txt
  ( env $v timeout 200 python bench.py --quick --steps 40 --warmup 5 2>&1 | tail -1 | cut -c1-300 ) >> gpurun_out/quick_${TAG}.txt
done
cat gpurun_out/quick_${TAG}.txt
M=sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor_subpipe_hmma.sum,sm__inst_executed_pipe_tensor.sum,sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum,sm__cycles_elapsed.max
( timeout 300 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/calib_${TAG}.csv -k regex:"gemm|nvjet|cutlass|conv_tc|sm100|xmma" python tools/ncu_calib_tensor.py 2>&1 | tail -4 ) > gpurun_out/calib_${TAG}.txt
cat gpurun_out/calib_${TAG}.txt
python tools/profile_timeline.py gen > gpurun_out/timeline_gen_${TAG}.txt 2>&1
head -3 gpurun_out/timeline_gen_${TAG}.txt | tail -2
