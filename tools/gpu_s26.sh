#!/usr/bin/env bash
# round-2 visit 26: localise the failure of visit 25 (default knobs) launch by launch
mkdir -p gpurun_out
( GG_CUDA_GRAPH=0 timeout 300 python tools/diag_plan.py 2>&1 | grep -v "^  File" | head -60 | cut -c1-400 ) > gpurun_out/diag_s26.txt
echo "---- with graphs" >> gpurun_out/diag_s26.txt
( timeout 300 python tools/diag_plan.py 2>&1 | grep -v "^  File" | head -60 | cut -c1-400 ) >> gpurun_out/diag_s26.txt
echo "---- pytest fusion" >> gpurun_out/diag_s26.txt
( timeout 600 python -m pytest tests/test_gpu_fusion.py -m gpu -q --no-header -x 2>&1 | grep -v "^  File" | head -80 | cut -c1-300 ) >> gpurun_out/diag_s26.txt
cat gpurun_out/diag_s26.txt
