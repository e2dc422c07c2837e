#!/usr/bin/env bash
# round-2 visit 24: element-wise cluster fusion — bit-identity tests + step time with / without it
bash tools/gpu_quick.sh "fusion" s24 "GG_FUSE_EW=0;GG_FUSE_EW=1"
for cfg in face ssgan; do
  for v in 0 1; do
    echo "== $cfg GG_FUSE_EW=$v" >> gpurun_out/quick_s24.txt
    ( env GG_FUSE_EW=$v timeout 200 python bench.py --quick --config $cfg --steps 20 --warmup 5 2>&1 | tail -1 | cut -c1-300 ) >> gpurun_out/quick_s24.txt
  done
done
cat gpurun_out/quick_s24.txt
