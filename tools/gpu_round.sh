#!/bin/bash
# One GPU call: gpu test tier, the default bench line, the ncu launch list of the same command and
# full captures of the hot kernels.  Results land in gpurun_out/ (copy what matters to profiles/).
#   gpurun --timeout 1800 -- 'bash tools/gpu_round.sh r02b [kernels...]'
TAG=${1:-r02}
shift
KERNELS=${@:-warp_tiles}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "== pytest -m gpu: $(tail -n 2 gpurun_out/${TAG}_pytest.log | tr '\n' ' ')"
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "== bench rc=$? $(cut -c1-400 gpurun_out/${TAG}_bench.json)"
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-other-configs"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_cfg4.csv $CMD > gpurun_out/${TAG}_launches_cfg4.log 2>&1
for K in $KERNELS; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o gpurun_out/${TAG}_cfg4_$K $CMD > gpurun_out/${TAG}_cfg4_$K.log 2>&1
done
ls -la gpurun_out | grep ${TAG}
