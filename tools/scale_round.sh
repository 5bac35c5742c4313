#!/bin/bash
# bench.py at N ranks of one box (N = number of visible GPUs), the way the driver launches it.
N=${1:-2}
TAG=${2:-r02}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/${TAG}_scale_n$N.json 2> gpurun_out/${TAG}_scale_n$N.err
echo "rc=$? $(tail -n 1 gpurun_out/${TAG}_scale_n$N.json | cut -c1-300)"
tail -n 3 gpurun_out/${TAG}_scale_n$N.err
