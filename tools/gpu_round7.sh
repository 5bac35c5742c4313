#!/bin/bash
# A/B of one vs two download streams in the streamed pipeline
export TAG=${1:-r02u}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for V in 1 2 3 1 2; do
  P360_DOWN_STREAMS=$V P360_PROBE_SHORT=1 P360_PROBE_NO_PAGEABLE=1 timeout 300 python tools/e2e_probe2.py cfg4 > gpurun_out/${TAG}_e2e_down_streams_$V.log 2>&1
  echo "== download streams=$V: $(grep 'windows=12' gpurun_out/${TAG}_e2e_down_streams_$V.log)"
done
