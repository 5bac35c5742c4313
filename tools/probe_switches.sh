#!/bin/bash
# One GPU call: A/B of the seam plan (direct tiles) against the float-everything pipeline on the
# benchmark composite, byte / tolerance checks, and the gpu test tier.  Results land in gpurun_out/.
#
#   gpurun --timeout 900 -- 'bash tools/probe_switches.sh'
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {   # name, timeout, command...
    local name=$1 limit=$2; shift 2
    echo "== $name"
    timeout "$limit" "$@" > "gpurun_out/probe_$name.json" 2> "gpurun_out/probe_$name.err"
    echo "rc=$? $(tail -n 1 "gpurun_out/probe_$name.json" | cut -c1-900)"
}
run cfg4_direct 180 python tools/maps_probe.py cfg4 --direct --h-rows 4
run cfg3_direct 180 python tools/maps_probe.py cfg3 --direct --h-rows 4
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/probe_pytest.log 2>&1
echo "== pytest -m gpu: $(tail -n 3 gpurun_out/probe_pytest.log)"
