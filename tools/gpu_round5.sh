#!/bin/bash
# short A/B of the default bench line (no other configs, no CPU baseline) + the gpu tier subset that exercises the warp
export TAG=${1:-r02r}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python bench.py --no-other-configs --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "== bench rc=$? $(python tools/show_bench.py gpurun_out/${TAG}_bench.json 2>/dev/null | head -2 | cut -c1-300)"
python - <<'P'
import json, os
d = json.loads(open(f"gpurun_out/{os.environ['TAG']}_bench.json").read().strip().splitlines()[-1])
print({k: round(v["ms_per_step"], 3) for k, v in d["kernels"].items()}, d["mosaic_checksum"], d["e2e"]["ms_per_step"],
      d.get("e2e_pageable", {}).get("ms_per_step"), d["roofline"]["frac"])
P
timeout 600 python -m pytest tests -m gpu -x -q -k "golden or warp_stage or edge or full_size_cfg4 or cfg3 or cfg2 or pole or tiny or pageable or streamed" > gpurun_out/${TAG}_pytest.log 2>&1
echo "== pytest subset: $(tail -n 2 gpurun_out/${TAG}_pytest.log | tr '\n' ' ')"
