#!/bin/bash
# One GPU call that times every opt-in switch against the default on the benchmark composite and
# checks that the bytes stay identical.  Results land in gpurun_out/ (copy what matters to profiles/).
#
#   gpurun --timeout 600 -- 'bash tools/probe_switches.sh'
#
# Each probe is a separate short process with its own timeout: a switch that misbehaves on the GPU
# costs its own slot, not the call.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {   # name, timeout, command...
    local name=$1 limit=$2; shift 2
    echo "== $name"
    timeout "$limit" "$@" > "gpurun_out/probe_$name.json" 2> "gpurun_out/probe_$name.err"
    echo "rc=$? $(tail -n 1 "gpurun_out/probe_$name.json" | cut -c1-600)"
}
run maps            120 python tools/maps_probe.py cfg4
run maps_hrows4     120 python tools/maps_probe.py cfg4 --h-rows 4
run maps_gate       120 python tools/maps_probe.py cfg4 --gate
run maps_gate_h4    120 python tools/maps_probe.py cfg4 --gate --h-rows 4
run cfg3_maps_gate  120 python tools/maps_probe.py cfg3 --gate
# the gpu-tier tests that so far only ran on the host build of the kernels
P360_WARP_GATE=1 P360_BLUR_H_ROWS=4 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
    -k "golden or seam_band or row_window or cut_anywhere or oracle_seeded or ring12" > gpurun_out/probe_pytest_switches.log 2>&1
echo "== pytest with gate + 64-cell blur: $(tail -n 1 gpurun_out/probe_pytest_switches.log)"
# end to end: default pipeline vs streamed row windows (both PCIe directions busy)
run e2e_default     240 python bench.py --steps 5 --warmup 3 --no-cpu-baseline
P360_STREAM_WINDOWS=3 run e2e_streamed 240 python bench.py --steps 5 --warmup 3 --no-cpu-baseline
