#!/bin/bash
# ncu evidence for profiles/: per-launch device time of one bench step + full captures of the hot kernels.
# Run on the GPU box:  bash tools/profile_ncu.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --workload cfg3 --steps 1 --warmup 1 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 2200 --csv --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_launches.log 2>&1
for K in multiband_collapse warp_patch blur_h blur_v pyramid_reduce; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 14 -c 2 -f -o $OUT/${TAG}_$K $CMD > $OUT/${TAG}_$K.log 2>&1
done
ls -la $OUT | tail -12
