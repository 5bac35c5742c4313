#!/bin/bash
# the committed tree once more: smoke + the whole gpu tier
export TAG=${1:-r02x}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1
echo "== smoke rc=$? $(tail -n 3 gpurun_out/${TAG}_smoke.log | tr '\n' ' ')"
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "== pytest -m gpu: $(tail -n 2 gpurun_out/${TAG}_pytest.log | tr '\n' ' ')"
