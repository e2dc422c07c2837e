#!/usr/bin/env bash
# round-2 visit 1: production-shape parity tests, conv timelines under split / cluster / n_tile variants, PDL variants
set -u
mkdir -p gpurun_out
TL=graphical-gan_b200/lib/libgg_b200_tl.so
( timeout 900 python -m pytest tests/test_gpu_production_shapes.py -m gpu -x -q -s 2>&1 | tail -150 | cut -c1-260 ) > gpurun_out/pytest_prod_s1.log
tail -5 gpurun_out/pytest_prod_s1.log
: > gpurun_out/timeline_s1.txt
for v in "GG_X=0" "GG_TC_CLUSTER=1" "GG_TC_NTILE=64 GG_TC_SPLITS=1" "GG_TC_SPLITS=1" "GG_TC_STAGES=6"; do
  echo "== $v" >> gpurun_out/timeline_s1.txt
  ( env $v GG_LIB=$TL timeout 120 python tools/timeline_conv.py "D.2" "E.2 fwd" "gemm 64x512x512" 2>&1 | tail -20 ) >> gpurun_out/timeline_s1.txt
done
: > gpurun_out/time_conv_s1.txt
for v in "GG_X=0" "GG_PDL=1" "GG_PDL=1 GG_PDL_COOP=1" "GG_TC_NTILE=64 GG_TC_SPLITS=1" "GG_TC_NTILE=64 GG_TC_SPLITS=1 GG_PDL=1" "GG_TC_CLUSTER=1" "GG_TC_STAGES=6"; do
  echo "== $v" >> gpurun_out/time_conv_s1.txt
  ( env $v timeout 120 python tools/time_conv.py dom 2>&1 | tail -4 ) >> gpurun_out/time_conv_s1.txt
done
cat gpurun_out/time_conv_s1.txt
: > gpurun_out/quick_s1.txt
for v in "GG_X=0" "GG_PDL=1" "GG_PDL=1 GG_PDL_COOP=1"; do
  echo "== $v" >> gpurun_out/quick_s1.txt
  ( env $v timeout 200 python bench.py --quick --steps 40 --warmup 5 2>&1 | tail -1 | cut -c1-300 ) >> gpurun_out/quick_s1.txt
done
cat gpurun_out/quick_s1.txt
( GG_PDL=1 GG_PDL_COOP=1 timeout 600 python -m pytest tests -m gpu -q --no-header -k "gmgan_step or models or wgan or golden" 2>&1 | tail -15 | cut -c1-250 ) > gpurun_out/pytest_pdl_s1.log
tail -4 gpurun_out/pytest_pdl_s1.log
