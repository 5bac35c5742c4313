"""Image ingest of the reference's CLI (stitcher.py:416-421): read every image of a directory
(``cv2.imread`` on a few threads — the decoders release the GIL) and shrink it by the ``-s``
factor **on the device** (p360_resize_u8: bit-exact with ``cv2.resize(im, None, fx=1/S, fy=1/S)``,
including OpenCV's reroute of the exact 2x shrink to the 2 x 2 area mean), through pinned staging.
The registration code of the reference (features.py / bundle_adj.py) wants host arrays, so the
resized images come back to the host; uploads, kernels and downloads of consecutive images overlap.

Measured on a B200 box (tools/ingest_probe.py, 36 images of 4000 x 3000, profiles/r02z_ingest_probe.log):
the kernel takes 41 us per image (S = 2; 1.1 TB/s on the 45 MB it reads and writes) where the host's
cv2.resize takes 1.05 ms on 16 threads — but end to end the device path is PCIe-bound (36 MB up for
9 MB of result): 1.66 ms per image against 1.05 ms (S = 1.5: 2.04 against 1.73 ms).  It pays once the
full-resolution image is on the device anyway (a device decoder); until then it is the bit-exact
device form of this stage, and P360_DEVICE_RESIZE=0 keeps the host call.
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import _lib, geometry as geo

EXTENSIONS = (".jpg", ".png", ".bmp")


def list_images(path):
    """The files the reference's ``main`` reads, in its order (stitcher.py:410-416)."""
    exts = list(EXTENSIONS) + [e.upper() for e in EXTENSIONS]
    return [f for f in os.listdir(path) if any(f.endswith(e) for e in exts)]


def resize_on_device(comp, images, shrink):
    """[cv2.resize(im, None, fx=1/shrink, fy=1/shrink) for im in images] computed on ``comp``'s GPU;
    uint8 HxW, HxWx3 or HxWx4 arrays in, arrays of the same kind out."""
    import torch
    from .compositor import parallel_copy
    f = 1.0 / shrink
    side, main = comp.copy_stream(), torch.cuda.current_stream(comp.device)
    area2 = abs(shrink - 2.0) < np.finfo(np.float64).eps
    # one pinned landing buffer for all the shrunk images (kept by the compositor, grow-only)
    sizes = []
    for img in images:
        if img.dtype != np.uint8 or img.ndim not in (2, 3) or (img.ndim == 3 and img.shape[2] not in (1, 3, 4)):
            raise TypeError("resize_on_device takes uint8 HxW, HxWx3 or HxWx4 images")
        dh, dw = geo.resize_dsize(img.shape[0], img.shape[1], f)
        if dh < 1 or dw < 1:
            raise ValueError(f"a {img.shape[0]} x {img.shape[1]} image cannot be shrunk by {shrink}")
        sizes.append(dh * dw * (1 if img.ndim == 2 else img.shape[2]))
    total = int(sum(sizes))
    stage = getattr(comp, "_ingest_stage", None)
    if stage is None or stage.numel() < total:
        stage = comp._ingest_stage = torch.empty(max(total, 1), dtype=torch.uint8, pin_memory=comp.device.type == "cuda")
    tables, out, pending, at = {}, [], [], 0
    for img, size in zip(images, sizes):
        h, w = img.shape[:2]
        c = 1 if img.ndim == 2 else img.shape[2]
        dh, dw = geo.resize_dsize(h, w, f)
        if (dh, dw) == (h, w):                           # cv2.resize copies when nothing changes
            out.append(img.copy())
            continue
        if (h, w) not in tables and not area2:
            xo, xw = geo.resize_tables(w, dw, f, True)
            yo, yw = geo.resize_tables(h, dh, f, False)
            tables[(h, w)] = tuple(comp._to_device(t) for t in (xo, xw, yo, yw))
        host = torch.from_numpy(np.ascontiguousarray(img))
        if not host.is_pinned() and host.numel() >= comp.stage_min_bytes:
            host = comp._stage_pageable(host, side)
        with torch.cuda.stream(side):
            dev = host.to(comp.device, non_blocking=host.is_pinned())
            if getattr(host, "_p360_slot", None) is not None:
                busy = torch.cuda.Event()
                busy.record(side)
                host._p360_slot[1] = busy
            small = torch.empty((dh, dw) + img.shape[2:], dtype=torch.uint8, device=comp.device)
            t = tables.get((h, w), (None,) * 4)
            _lib.call("p360_resize_u8", dev.data_ptr(), h, w, c, small.data_ptr(), dh, dw, _lib.ptr(t[0]), _lib.ptr(t[1]),
                      _lib.ptr(t[2]), _lib.ptr(t[3]), int(area2), side.cuda_stream)
            landing = stage[at:at + size].view(small.shape)
            landing.copy_(small, non_blocking=True)
            at += size
        pending.append((len(out), landing, dev, small))
        out.append(None)
    side.synchronize()
    main.wait_stream(side)
    for k, landing, _, _ in pending:                     # plain (pageable) arrays, like cv2's
        fresh = np.empty(tuple(landing.shape), np.uint8)
        parallel_copy(fresh, landing.numpy())
        out[k] = fresh
    return out


def read_images(path, shrink=1.0, comp=None, workers=8):
    """The reference's ingest (stitcher.py:416-421): every image of ``path`` decoded (BGR uint8) and,
    for ``shrink > 1``, shrunk on the GPU.  Returns the list ``main`` passes to the registration."""
    import cv2
    files = list_images(path)
    with ThreadPoolExecutor(max(1, min(workers, len(files) or 1))) as pool:
        imgs = list(pool.map(lambda name: cv2.imread(os.path.join(path, name)), files))
    if shrink > 1 and os.environ.get("P360_DEVICE_RESIZE", "1") != "1":
        return [cv2.resize(im, None, fx=1 / shrink, fy=1 / shrink) for im in imgs]     # the reference's call, on the host
    if shrink > 1:
        if comp is None:
            from .stitcher import _compositor
            comp = _compositor()
        imgs = resize_on_device(comp, imgs, shrink)
    return imgs
