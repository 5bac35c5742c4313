"""Quick A/B of the seam-band maps on the GPU: device time per composite with and without,
per-kernel breakdown, and byte equality of the two mosaics.  Pixel content does not matter for
the timing (all access patterns follow the geometry), so one random image stands in for every
view: the probe starts in seconds.

    python tools/maps_probe.py [cfg4] [--scale 1] [--steps 5]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main(comp=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("workload", nargs="?", default="cfg4")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--h-rows", type=int, default=1, help="4: horizontal blur lists in 64-cell segments")
    ap.add_argument("--no-pack", action="store_true", help="sources stay u8 x 3 as uploaded (no RGBX packing)")
    ap.add_argument("--direct", action="store_true", help="maps_on = the seam plan with direct tiles (p360_seam_plan_build)")
    args = ap.parse_args()
    t0 = time.time()
    import torch
    from pano360_b200 import geometry as geo, synth
    from pano360_b200.compositor import Compositor
    wl = synth.workload(args.workload, scale=args.scale)
    regs = synth.make_views(wl, only=set())
    img = np.random.default_rng(0).integers(0, 256, (wl.height, wl.width, 3), dtype=np.uint8)
    for r in regs:
        r.img = img
    plan = geo.plan_mosaic(regs, wl.blend == "multiband", 1e9)
    comp = comp or Compositor()
    comp.blur_h_rows = args.h_rows
    direct = args.direct
    src = comp.upload(regs, pack=not args.no_pack)
    out = {"workload": args.workload, "scale": args.scale, "mosaic": list(plan.shape), "setup_s": round(time.time() - t0, 1)}
    mosaics = {}
    for maps in (False, True):
        comp.seam_maps = maps
        comp.direct = maps and direct
        for _ in range(2 if args.steps else 0):
            comp.composite(regs, src, plan, wl.blend, wl.n_levels)
        torch.cuda.synchronize()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record(torch.cuda.current_stream())
        for _ in range(args.steps):
            mosaic, _ = comp.composite(regs, src, plan, wl.blend, wl.n_levels)
        end.record(torch.cuda.current_stream())
        torch.cuda.synchronize()
        ms = start.elapsed_time(end) / max(args.steps, 1)
        comp.trace = []
        mosaic, _ = comp.composite(regs, src, plan, wl.blend, wl.n_levels)
        torch.cuda.synchronize()
        kernels = {}
        for name, _, a, b in comp.trace:
            kernels[name] = round(kernels.get(name, 0.0) + a.elapsed_time(b), 3)
        comp.trace = None
        mosaics[maps] = mosaic.clone()
        out["maps_on" if maps else "maps_off"] = {"ms_per_step": round(ms, 3), "kernels_ms": kernels}
        print(json.dumps(out), flush=True)
    out["identical"] = bool(torch.equal(mosaics[False], mosaics[True]))
    diff = (mosaics[False].to(torch.int16) - mosaics[True].to(torch.int16)).abs()
    out["max_abs_diff"], out["differing_px"] = int(diff.max()), int((diff > 0).sum())
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
