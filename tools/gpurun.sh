#!/usr/bin/env bash
# local pre-flight for a GPU visit: rebuild libgg_b200.so (a stale library against a changed cabi.py signature is a segfault on the
# box), run the CPU symbol check, then hand the command to gpurun.   usage: tools/gpurun.sh <timeout-seconds> '<command>'
set -euo pipefail
cd "$(dirname "${BASH_SOURCE[0]}")/.."
bash graphical-gan_b200/build.sh | tail -1
python -m pytest tests/test_cpu_host.py -q -x -k "cabi" 2>&1 | tail -1
exec /usr/local/graft/bin/gpurun --timeout "$1" -- "$2"
