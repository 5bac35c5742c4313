"""Where does the end-to-end time of stitch() go?  (run on the GPU box)

Raw PCIe rates (each direction alone and both at once), host memcpy rates for the staging of
pageable buffers, and stitch() end to end with pinned / pageable buffers for several window counts
of the streamed pipeline, with a timeline of the last run.
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from pano360_b200 import geometry as geo, stitcher, synth

wl = synth.workload(sys.argv[1] if len(sys.argv) > 1 else "cfg4")
regs = synth.make_views(wl)
pageable = [r.img for r in regs]
for r in regs:
    t = torch.empty(r.img.shape, dtype=torch.uint8, pin_memory=True)
    t.numpy()[...] = r.img
    r.img, r._pin = t.numpy(), t
stitcher.MAX_RESOLUTION = wl.max_resolution
comp = stitcher._compositor()
plan = geo.plan_mosaic(regs, wl.blend == "multiband", wl.max_resolution)
out = torch.empty(plan.shape + (3,), dtype=torch.uint8, pin_memory=True)


def sync():
    torch.cuda.synchronize()


def timed(fn, n=3):
    fn(); sync(); t0 = time.perf_counter()
    for _ in range(n):
        fn()
    sync()
    return (time.perf_counter() - t0) / n * 1e3


nbytes = sum(r.img.nbytes for r in regs)
dev_imgs = [torch.empty(r.img.shape, dtype=torch.uint8, device="cuda") for r in regs]
dev = torch.empty(plan.shape + (3,), dtype=torch.uint8, device="cuda")
up, down = torch.cuda.Stream(), torch.cuda.Stream()


def h2d():
    with torch.cuda.stream(up):
        for d, r in zip(dev_imgs, regs):
            d.copy_(r._pin, non_blocking=True)


def d2h():
    with torch.cuda.stream(down):
        out.copy_(dev, non_blocking=True)


ms = timed(h2d); print(f"H2D pinned {nbytes/1e6:.0f} MB: {ms:.1f} ms = {nbytes/ms/1e6:.1f} GB/s")
ms = timed(d2h); print(f"D2H pinned {dev.numel()/1e6:.0f} MB: {ms:.1f} ms = {dev.numel()/ms/1e6:.1f} GB/s")
ms = timed(lambda: (h2d(), d2h())); print(f"both directions at once: {ms:.1f} ms")

# host-side staging of pageable buffers
stage = torch.empty(pageable[0].shape, dtype=torch.uint8, pin_memory=True)
print("torch threads:", torch.get_num_threads(), "cpus:", os.cpu_count())
t0 = time.perf_counter()
for p in pageable:
    stage.copy_(torch.from_numpy(p))
dt = time.perf_counter() - t0; print(f"pageable -> pinned, torch copy_: {nbytes/dt/1e9:.1f} GB/s ({dt*1e3:.0f} ms)")
t0 = time.perf_counter()
for p in pageable:
    np.copyto(stage.numpy(), p)
dt = time.perf_counter() - t0; print(f"pageable -> pinned, np.copyto 1 thread: {nbytes/dt/1e9:.1f} GB/s")
from concurrent.futures import ThreadPoolExecutor
for workers in (4, 8, 16):
    pool = ThreadPoolExecutor(workers)
    def par_copy(dst, src, pool=pool, workers=workers):
        n = dst.shape[0]
        cuts = [n * k // workers for k in range(workers + 1)]
        list(pool.map(lambda ab: np.copyto(dst[ab[0]:ab[1]], src[ab[0]:ab[1]]), zip(cuts, cuts[1:])))
    t0 = time.perf_counter()
    for p in pageable:
        par_copy(stage.numpy(), p)
    dt = time.perf_counter() - t0; print(f"pageable -> pinned, np.copyto x{workers} threads: {nbytes/dt/1e9:.1f} GB/s")
    t0 = time.perf_counter(); fresh = np.empty(out.shape, np.uint8); par_copy(fresh, out.numpy())
    dt = time.perf_counter() - t0; print(f"pinned -> fresh ndarray ({out.numel()/1e6:.0f} MB) x{workers}: {dt*1e3:.0f} ms")
t0 = time.perf_counter(); fresh = np.empty(out.shape, np.uint8); fresh[...] = out.numpy()
print(f"pinned -> fresh ndarray 1 thread: {(time.perf_counter()-t0)*1e3:.0f} ms")
t0 = time.perf_counter(); rc = torch.cuda.cudart().cudaHostRegister(pageable[0].ctypes.data, pageable[0].nbytes, 0)
print(f"cudaHostRegister 36 MB: {(time.perf_counter()-t0)*1e3:.1f} ms rc={rc}")
torch.cuda.cudart().cudaHostUnregister(pageable[0].ctypes.data)
t0 = time.perf_counter(); big = torch.empty(out.shape, dtype=torch.uint8, pin_memory=True)
print(f"pinned alloc {out.numel()/1e6:.0f} MB: {(time.perf_counter()-t0)*1e3:.0f} ms"); del big

# end to end
def e2e(imgs=None, out_arr=out.numpy()):
    rr = regs
    if imgs is not None:
        from pano360_b200.camera import Image
        rr = [Image(i, r.rot, r.intr) for i, r in zip(imgs, regs)]
    return stitcher.stitch(rr, blender=stitcher.BLENDERS[wl.blend], n_levels=wl.n_levels, out=out_arr)


for windows in (0, 2, 3, 4, 6):
    stitcher.STREAM_WINDOWS = windows
    print(f"stitch pinned in/out, windows={windows}: {timed(e2e):.1f} ms")
stitcher.STREAM_WINDOWS = int(os.environ.get("P360_STREAM_WINDOWS", "3"))
comp.timeline = []
e2e(); sync()
t0 = comp.timeline[0][1]
for label, ev in comp.timeline:
    print(f"  {t0.elapsed_time(ev):8.2f} ms  {label}")
comp.timeline = None
print(f"stitch pageable in, fresh out: {timed(lambda: e2e(pageable, None), 2):.1f} ms")
