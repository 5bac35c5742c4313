"""PCIe copy rates on one B200: contiguous vs rectangle (cudaMemcpy2DAsync through p360_copy_rect)
copies of mosaic / image sized buffers, both directions, alone and together."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from pano360_b200 import _lib  # noqa: E402

_lib.load()
H, W = 8819, 31654
dev = torch.empty((H, W, 3), dtype=torch.uint8, device="cuda")
host = torch.empty((H, W, 3), dtype=torch.uint8, pin_memory=True)
img_h = torch.empty((36, 3000, 4000, 3), dtype=torch.uint8, pin_memory=True)
img_d = torch.empty((36, 3000, 4000, 3), dtype=torch.uint8, device="cuda")
up, down = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, n=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


def d2h_rect(cols, rows_per_copy=H):
    def run():
        for xa in range(0, W, cols):
            xb = min(xa + cols, W)
            for ya in range(0, H, rows_per_copy):
                yb = min(ya + rows_per_copy, H)
                _lib.call("p360_copy_rect", host.data_ptr() + 3 * (ya * W + xa), 3 * W, dev.data_ptr() + 3 * (ya * W + xa), 3 * W,
                          3 * (xb - xa), yb - ya, down.cuda_stream)
    return run


def h2d_rect(c0, c1):
    def run():
        for i in range(36):
            _lib.call("p360_copy_rect", img_d[i].data_ptr() + 3 * c0, 12000, img_h[i].data_ptr() + 3 * c0, 12000, 3 * (c1 - c0), 3000,
                      up.cuda_stream)
    return run


def h2d_full():
    with torch.cuda.stream(up):
        img_d.copy_(img_h, non_blocking=True)


def d2h_full():
    with torch.cuda.stream(down):
        host.copy_(dev, non_blocking=True)


mb_m, mb_i = H * W * 3 / 1e6, 36 * 36.0
print(f"D2H mosaic contiguous: {timed(d2h_full):.2f} ms ({mb_m:.0f} MB)")
for cols in (1344, 2688, 5376, 10752, W):
    ms = timed(d2h_rect(cols))
    print(f"D2H mosaic in column windows of {cols}: {ms:.2f} ms = {mb_m / ms:.1f} GB/s")
ms = timed(d2h_rect(2688, 4410)); print(f"D2H 2688-column windows in 2 row bands: {ms:.2f} ms")
print(f"H2D images contiguous: {timed(h2d_full):.2f} ms ({mb_i:.0f} MB)")
for c0, c1 in ((0, 4000), (336, 3664), (1000, 3000)):
    ms = timed(h2d_rect(c0, c1))
    mb = 36 * 3000 * (c1 - c0) * 3 / 1e6
    print(f"H2D image columns {c0}-{c1}: {ms:.2f} ms = {mb / ms:.1f} GB/s ({mb:.0f} MB)")
ms = timed(lambda: (h2d_full(), d2h_full())); print(f"both contiguous at once: {ms:.2f} ms")
ms = timed(lambda: (h2d_rect(336, 3664)(), d2h_rect(2688)())); print(f"both as rectangles at once: {ms:.2f} ms")
ms = timed(lambda: (h2d_rect(336, 3664)(), d2h_full())); print(f"H2D rectangles + D2H contiguous: {ms:.2f} ms")

# the multi-GPU path downloads into a /dev/shm mapping registered with cudaHostRegister: same rate?
import numpy as np  # noqa: E402
name = f"/dev/shm/p360_probe_{os.getpid()}"
with open(name, "wb") as fid:
    fid.truncate(H * W * 3)
shm = np.memmap(name, dtype=np.uint8, mode="r+", shape=(H * W * 3,))
os.unlink(name)
shm[:] = 0                                             # touch every page
rt = torch.cuda.cudart()
t0 = time.perf_counter(); err = rt.cudaHostRegister(shm.ctypes.data, shm.nbytes, 0)
print(f"cudaHostRegister of a {shm.nbytes / 1e6:.0f} MB /dev/shm mapping: {(time.perf_counter() - t0) * 1e3:.0f} ms ({err})")
shm_t = torch.from_numpy(shm).view(H, W, 3)


def d2h_shm():
    with torch.cuda.stream(down):
        _lib.call("p360_copy_rect", shm_t.data_ptr(), 3 * W, dev.data_ptr(), 3 * W, 3 * W, H, down.cuda_stream)


def d2h_shm_cols(cols=3968):
    for xa in range(0, W, cols):
        xb = min(xa + cols, W)
        _lib.call("p360_copy_rect", shm_t.data_ptr() + 3 * xa, 3 * W, dev.data_ptr() + 3 * xa, 3 * W, 3 * (xb - xa), H, down.cuda_stream)


ms = timed(d2h_shm); print(f"D2H mosaic into the registered /dev/shm mapping: {ms:.2f} ms = {mb_m / ms:.1f} GB/s")
ms = timed(d2h_shm_cols); print(f"... in 8 column strips: {ms:.2f} ms = {mb_m / ms:.1f} GB/s")
rt.cudaHostUnregister(shm.ctypes.data)
