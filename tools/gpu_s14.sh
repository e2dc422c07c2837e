#!/usr/bin/env bash
# round-2 visit 14: un-split 128x64 tiles for the batched shapes, after the producer / exchange rework
set -u
mkdir -p gpurun_out
TAG="${1:-s14}"
: > gpurun_out/time_conv_${TAG}.txt
for v in "GG_X=0" "GG_TC_NTILE=64 GG_TC_SPLITS=1" "GG_TC_NTILE=64 GG_TC_SPLITS=1 GG_TC_STAGES=6" "GG_TC_NTILE=64" "GG_TC_NTILE=32 GG_TC_SPLITS=1 GG_TC_STAGES=6"; do
  echo "== $v" >> gpurun_out/time_conv_${TAG}.txt
  ( env $v timeout 300 python tools/time_conv.py batched 2>&1 | grep "64->128\|128->256" ) >> gpurun_out/time_conv_${TAG}.txt
done
cat gpurun_out/time_conv_${TAG}.txt
