#!/bin/bash
# ncu evidence for profiles/: per-launch device time of bench steps + full captures of the hot kernels.
# Run on the GPU box:  bash tools/profile_ncu.sh [tag] [workload]
TAG=${1:-r01}
WL=${2:-cfg4}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --workload $WL --steps 1 --warmup 1 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_$WL.csv $CMD > $OUT/${TAG}_launches_$WL.log 2>&1
for K in warp_batch pyramid_reduce blur_h_batch blur_v_batch multiband_collapse; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o $OUT/${TAG}_${WL}_$K $CMD > $OUT/${TAG}_${WL}_$K.log 2>&1
done
ls -la $OUT | tail -14
