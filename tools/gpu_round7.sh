#!/bin/bash
# A/B of the upload order of the streamed pipeline (seam-straddling images first / last)
export TAG=${1:-r02t}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for V in 1 0 1 0; do
  P360_STRADDLERS_LAST=$V P360_PROBE_SHORT=1 P360_PROBE_NO_PAGEABLE=1 timeout 300 python tools/e2e_probe2.py cfg4 > gpurun_out/${TAG}_e2e_straddlers_last_$V.log 2>&1
  echo "== straddlers last=$V: $(grep 'windows=12' gpurun_out/${TAG}_e2e_straddlers_last_$V.log)"
done
