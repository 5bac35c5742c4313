#!/bin/bash
# N ranks of one box: strips == single GPU byte for byte (small layouts), then the cfg4 bench line.
N=${1:-8}
TAG=${2:-r02m}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    tools/check_strips.py > gpurun_out/${TAG}_check_strips_n$N.log 2>&1
echo "check_strips rc=$?"; grep "ranks" gpurun_out/${TAG}_check_strips_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 5 --warmup 3 --no-other-configs > gpurun_out/${TAG}_scale_n$N.json 2> gpurun_out/${TAG}_scale_n$N.err
echo "rc=$? $(tail -n 1 gpurun_out/${TAG}_scale_n$N.json | cut -c1-200)"
python tools/show_bench.py gpurun_out/${TAG}_scale_n$N.json 2>/dev/null | head -12
tail -n 3 gpurun_out/${TAG}_scale_n$N.err
