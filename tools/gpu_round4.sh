#!/bin/bash
# Final 1-GPU call of the round: gpu test tier, the default bench line (all configs + CPU baseline),
# the end-to-end probe, the ncu launch list of the bench command and full captures of the hot kernels.
#   gpurun --timeout 2400 -- 'bash tools/gpu_round4.sh r02p'
TAG=${1:-r02p}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/${TAG}_pytest.log 2>&1
echo "== pytest -m gpu: $(tail -n 2 gpurun_out/${TAG}_pytest.log | tr '\n' ' ')"; grep "cfg4 whole" gpurun_out/${TAG}_pytest.log
timeout 300 python tools/copy_probe.py > gpurun_out/${TAG}_copy_probe.log 2>&1; grep -i "shm" gpurun_out/${TAG}_copy_probe.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "== bench rc=$? $(python tools/show_bench.py gpurun_out/${TAG}_bench.json 2>/dev/null | head -3 | cut -c1-300)"
P360_PROBE_SHORT=1 timeout 300 python tools/e2e_probe2.py cfg4 > gpurun_out/${TAG}_e2e_probe.log 2>&1
echo "== e2e probe rc=$?"; grep "stitch" gpurun_out/${TAG}_e2e_probe.log
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-other-configs"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_cfg4.csv $CMD > gpurun_out/${TAG}_launches_cfg4.log 2>&1
for K in warp_tiles pack_rgbx_batch seam_candidates pyramid_reduce_list blur_h_list blur_v_list multiband_collapse; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o gpurun_out/${TAG}_cfg4_$K $CMD > gpurun_out/${TAG}_cfg4_$K.log 2>&1
  echo "ncu $K rc=$?"
done
ls -la gpurun_out | grep ${TAG} | awk '{print $5, $9}'
