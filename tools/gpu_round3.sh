#!/bin/bash
# copy-rate probe + end-to-end probe + bench (one GPU)
TAG=${1:-r02l}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/copy_probe.py > gpurun_out/${TAG}_copy_probe.log 2>&1
echo "== copy probe rc=$?"; cat gpurun_out/${TAG}_copy_probe.log
timeout 600 python tools/e2e_probe2.py cfg4 > gpurun_out/${TAG}_e2e_probe.log 2>&1
echo "== e2e probe rc=$?"; grep "stitch" gpurun_out/${TAG}_e2e_probe.log
timeout 900 python bench.py --no-other-configs > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "== bench rc=$? $(python tools/show_bench.py gpurun_out/${TAG}_bench.json 2>/dev/null | cut -c1-600)"
