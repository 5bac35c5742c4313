#!/bin/bash
# last check of the round: bench line with the other configs (no CPU baseline) + the plan / golden tests
export TAG=${1:-r02v}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "== bench rc=$? $(python tools/show_bench.py gpurun_out/${TAG}_bench.json 2>/dev/null | head -2 | cut -c1-300)"; tail -n 2 gpurun_out/${TAG}_bench.err
python - <<'P'
import json, os
d = json.loads(open(f"gpurun_out/{os.environ['TAG']}_bench.json").read().strip().splitlines()[-1])
print({k: round(v["ms_per_step"], 3) for k, v in d["kernels"].items()}, d["mosaic_checksum"], round(d["e2e"]["ms_per_step"], 2), d["roofline"]["frac"])
for k, v in d["other_configs"].items():
    print(k, round(v["value"]), round(v["ms_per_step"], 3), round(v["e2e"]["ms_per_step"], 2), (v.get("roofline") or {}).get("kernel"), (v.get("roofline") or {}).get("frac"))
P
timeout 600 python -m pytest tests -m gpu -x -q -k "seam_plan or golden or many_small or source_rect or column_windows or full_size_cfg4_windows" > gpurun_out/${TAG}_pytest.log 2>&1
echo "== pytest subset: $(tail -n 2 gpurun_out/${TAG}_pytest.log | tr '\n' ' ')"
