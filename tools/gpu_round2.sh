#!/bin/bash
# One GPU call of round 2 (second half): the new window / rectangle tests first, then the whole gpu
# tier, the end-to-end probe (window count, source rectangles, timeline) and the default bench line.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round2.sh r02k'
TAG=${1:-r02k}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "column_windows or source_rect or row_window" > gpurun_out/${TAG}_pytest_new.log 2>&1
echo "== new tests: $(tail -n 3 gpurun_out/${TAG}_pytest_new.log | tr '\n' ' ')"
timeout 600 python tools/e2e_probe2.py cfg4 > gpurun_out/${TAG}_e2e_probe.log 2>&1
echo "== e2e probe rc=$?"; grep "stitch" gpurun_out/${TAG}_e2e_probe.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "== bench rc=$? $(cut -c1-300 gpurun_out/${TAG}_bench.json)"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "== pytest -m gpu: $(tail -n 3 gpurun_out/${TAG}_pytest.log | tr '\n' ' ')"
