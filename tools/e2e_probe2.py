"""End-to-end probe of stitcher.stitch at a BASELINE workload (default cfg4) on one B200: the
streamed column windows (count swept), with and without the source rectangles, the timeline of
the default setting, pageable inputs.   python tools/e2e_probe2.py [cfg4|cfg3]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from pano360_b200 import stitcher, synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
wl = synth.workload(name)
regs = synth.make_views(wl)
pageable = [r.img for r in regs]
for r in regs:
    t = torch.empty(r.img.shape, dtype=torch.uint8, pin_memory=True)
    t.numpy()[...] = r.img
    r._pin, r.img = t, t.numpy()
stitcher.MAX_RESOLUTION = wl.max_resolution
comp = stitcher._compositor()
first = stitcher.stitch(regs, blender=stitcher.BLENDERS[wl.blend], n_levels=wl.n_levels)
out = torch.empty(first.shape, dtype=torch.uint8, pin_memory=True)
import zlib
crc0 = zlib.crc32(first.tobytes())


def e2e(imgs=None, out_arr=out.numpy()):
    rr = regs
    if imgs is not None:
        from pano360_b200.camera import Image
        rr = [Image(i, r.rot, r.intr) for i, r in zip(imgs, regs)]
    return stitcher.stitch(rr, blender=stitcher.BLENDERS[wl.blend], n_levels=wl.n_levels, out=out_arr)


def timed(fn, n=5):
    fn(); fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


short = os.environ.get("P360_PROBE_SHORT") == "1"
for rects in ((True,) if short else (True, False)):
    comp.partial_uploads = rects
    for windows in ((0, 12) if short else (0, 6, 12, 16, 24)):
        stitcher.STREAM_WINDOWS = windows
        ms = timed(e2e)
        ok = zlib.crc32(out.numpy().tobytes()) == crc0
        print(f"stitch pinned in/out, rects={rects} windows={windows}: {ms:.2f} ms  upload {comp.last_upload_bytes / 1e6:.0f} MB  same bytes: {ok}", flush=True)
    comp.release(everything=True)
comp.partial_uploads = True
stitcher.STREAM_WINDOWS = int(os.environ.get("P360_STREAM_WINDOWS", "12"))
e2e(); e2e()
comp.timeline = []
e2e(); torch.cuda.synchronize()
t0 = comp.timeline[0][1]
for label, ev in comp.timeline:
    print(f"  {t0.elapsed_time(ev):8.2f} ms  {label}")
comp.timeline = None
if os.environ.get("P360_PROBE_NO_PAGEABLE") != "1":
    print(f"stitch pageable in, fresh out: {timed(lambda: e2e(pageable, None), 3):.1f} ms")
