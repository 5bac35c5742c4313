#!/bin/bash
# bench.py at N ranks exactly as the driver launches it (default flags: the cfg5 replicas run too)
N=${1:-2}
TAG=${2:-r02w}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/${TAG}_default_n$N.json 2> gpurun_out/${TAG}_default_n$N.err
echo "rc=$? $(tail -n 1 gpurun_out/${TAG}_default_n$N.json | cut -c1-200)"
python - <<P
import json
d = json.loads(open("gpurun_out/${TAG}_default_n$N.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["mosaic_checksum"], d["e2e"]["ms_per_step"], list(d.get("other_configs", {})))
c = d["other_configs"]["cfg5"]; print("cfg5", c["value"], c["ms_per_step"], c["roofline"]["kernel"], c["roofline"]["frac"], c["n_gpus"])
P
tail -n 3 gpurun_out/${TAG}_default_n$N.err
