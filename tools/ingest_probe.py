"""Ingest measurement on one B200: cv2.resize of the -s flag on the host (the reference's
call, OpenCV's own thread pool) against p360_resize_u8 — kernel alone (CUDA events, sources resident)
and end to end (pinned images up, shrunk images back) — for 36 images of 4000 x 3000, S = 2 and 1.5."""
import os
import sys
import time

import cv2
import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from pano360_b200 import _lib, geometry as geo, ingest  # noqa: E402
from pano360_b200.compositor import Compositor  # noqa: E402

comp = Compositor()
rng = np.random.default_rng(0)
n, h, w = 36, 3000, 4000
base = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
pinned = []
for i in range(n):
    t = torch.empty((h, w, 3), dtype=torch.uint8, pin_memory=True)
    t.numpy()[...] = np.roll(base, 17 * i, axis=1)
    pinned.append(t.numpy())
dev = [torch.from_numpy(p).cuda() for p in pinned[:8]]
print(f"cv2 threads {cv2.getNumThreads()}, cpus {os.cpu_count()}")
for shrink in (2.0, 1.5):
    f = 1.0 / shrink
    t0 = time.perf_counter()
    want = [cv2.resize(p, None, fx=f, fy=f) for p in pinned]
    host_ms = (time.perf_counter() - t0) * 1e3
    ingest.resize_on_device(comp, pinned, shrink)          # (warm: the pinned landing buffer is allocated once)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    got = ingest.resize_on_device(comp, pinned, shrink)
    e2e_ms = (time.perf_counter() - t0) * 1e3
    same = all(np.array_equal(a, b) for a, b in zip(got, want))
    dh, dw = geo.resize_dsize(h, w, f)
    area2 = abs(shrink - 2.0) < 1e-12
    tabs = (None,) * 4
    if not area2:
        xo, xw = geo.resize_tables(w, dw, f, True)
        yo, yw = geo.resize_tables(h, dh, f, False)
        tabs = tuple(comp._to_device(t) for t in (xo, xw, yo, yw))
    outs = [torch.empty((dh, dw, 3), dtype=torch.uint8, device="cuda") for _ in dev]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for rep in range(3):
        torch.cuda.synchronize()
        ev[0].record()
        for d, o in zip(dev, outs):
            _lib.call("p360_resize_u8", d.data_ptr(), h, w, 3, o.data_ptr(), dh, dw, _lib.ptr(tabs[0]), _lib.ptr(tabs[1]),
                      _lib.ptr(tabs[2]), _lib.ptr(tabs[3]), int(area2), comp.stream)
        ev[1].record()
        torch.cuda.synchronize()
    k_ms = ev[0].elapsed_time(ev[1]) / len(dev)
    nbytes = 3 * (h * w + dh * dw)
    print(f"S = {shrink}: host cv2.resize {host_ms / n:.2f} ms/image; device end to end {e2e_ms / n:.2f} ms/image "
          f"({n} images {e2e_ms:.0f} ms, identical: {same}); kernel {k_ms * 1e3:.0f} us/image = {nbytes / k_ms / 1e6:.0f} GB/s "
          f"on {nbytes / 1e6:.0f} MB read + written")
