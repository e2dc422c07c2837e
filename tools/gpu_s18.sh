#!/usr/bin/env bash
# round-2 visit 18: in-kernel timeline of the face-shaped (Ci = 32) tensor-core launches
set -u
mkdir -p gpurun_out
TAG="${1:-s18}"
TL=graphical-gan_b200/lib/libgg_b200_tl.so
( GG_LIB=$TL timeout 200 python tools/timeline_conv.py face "D.2 fwd batched" 2>&1 | tail -30 ) > gpurun_out/timeline_face_${TAG}.txt
cat gpurun_out/timeline_face_${TAG}.txt | cut -c1-250
( GG_LIB=$TL GG_TC_STAGES=0 timeout 200 python tools/timeline_conv.py "face D.2 fwd" 2>&1 | tail -6 ) > gpurun_out/timeline_face_deep_${TAG}.txt
cat gpurun_out/timeline_face_deep_${TAG}.txt | cut -c1-250
