"""Where does the end-to-end time of stitch() go?  (run on the GPU box)"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pano360_b200 import synth, stitcher, geometry as geo
wl = synth.workload(sys.argv[1] if len(sys.argv) > 1 else "cfg4")
regs = synth.make_views(wl)
for r in regs:
    t = torch.empty(r.img.shape, dtype=torch.uint8, pin_memory=True); t.numpy()[...] = r.img; r.img = t.numpy(); r._pin = t
print("pinned view is_pinned:", torch.from_numpy(regs[0].img).is_pinned())
stitcher.MAX_RESOLUTION = wl.max_resolution
comp = stitcher._compositor()
plan = geo.plan_mosaic(regs, wl.blend == "multiband", wl.max_resolution)
out = torch.empty(plan.shape + (3,), dtype=torch.uint8, pin_memory=True)
def sync(): torch.cuda.synchronize()
def t(fn, n=3):
    fn(); sync(); t0 = time.perf_counter()
    for _ in range(n): r = fn()
    sync(); return (time.perf_counter() - t0) / n * 1e3
nbytes = sum(r.img.nbytes for r in regs)
ms = t(lambda: [torch.from_numpy(r.img).to("cuda", non_blocking=True) for r in regs]); print(f"H2D raw {nbytes/1e6:.0f} MB: {ms:.1f} ms = {nbytes/ms/1e6:.1f} GB/s")
dev = torch.empty(plan.shape + (3,), dtype=torch.uint8, device="cuda")
ms = t(lambda: out.copy_(dev, non_blocking=True)); print(f"D2H raw {dev.numel()/1e6:.0f} MB: {ms:.1f} ms = {dev.numel()/ms/1e6:.1f} GB/s")
ms = t(lambda: comp.upload(regs)); print(f"upload (+pack): {ms:.1f} ms")
ms = t(lambda: comp.upload(regs, overlap=True)); print(f"upload overlap (+pack): {ms:.1f} ms")
src = comp.upload(regs); sync()
ms = t(lambda: comp.composite(regs, src, plan, wl.blend, wl.n_levels)); print(f"composite resident: {ms:.1f} ms")
ms = t(lambda: geo.plan_mosaic(regs, True, wl.max_resolution)); print(f"plan_mosaic host: {ms:.1f} ms")
def full(): return stitcher.stitch(regs, blender=stitcher.BLENDERS[wl.blend], n_levels=wl.n_levels, out=out.numpy())
ms = t(full); print(f"stitch e2e (pinned out): {ms:.1f} ms")
def full2(): return stitcher.stitch(regs, blender=stitcher.BLENDERS[wl.blend], n_levels=wl.n_levels)
ms = t(full2, 2); print(f"stitch e2e (fresh pageable out): {ms:.1f} ms")
