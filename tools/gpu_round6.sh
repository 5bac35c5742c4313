#!/bin/bash
# The committed state on one B200: smoke(), the whole gpu tier, the default bench line (all configs + CPU
# baseline), the reference arm, and the ncu launch list + full capture of the dominant kernel.
export TAG=${1:-r02s}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1
echo "== smoke rc=$? $(tail -n 3 gpurun_out/${TAG}_smoke.log | tr '\n' ' ')"
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/${TAG}_pytest.log 2>&1
echo "== pytest -m gpu: $(tail -n 2 gpurun_out/${TAG}_pytest.log | tr '\n' ' ')"; grep "cfg4 whole" gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "== bench rc=$? $(python tools/show_bench.py gpurun_out/${TAG}_bench.json 2>/dev/null | head -2 | cut -c1-300)"; tail -n 3 gpurun_out/${TAG}_bench.err
timeout 400 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
echo "== reference arm rc=$? $(cut -c1-300 gpurun_out/${TAG}_bench_reference.json)"
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-other-configs"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_cfg4.csv $CMD > gpurun_out/${TAG}_launches_cfg4.log 2>&1
for K in warp_tiles; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o gpurun_out/${TAG}_cfg4_$K $CMD > gpurun_out/${TAG}_cfg4_$K.log 2>&1
  echo "ncu $K rc=$?"
done
