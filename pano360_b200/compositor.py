"""Device-side driver of the compositing path: owns HBM buffers (torch is used
for allocation, streams and copies only) and sequences the sm_100a kernels of
``libpano360_b200.so`` through the C ABI.

A whole multiband composite is seven calls, whatever the number of images: one
packing launch over the uploaded parts of all images (K1p), the seam plan from the
geometry (K0), one tile warp (K1t: single-owner tiles straight to uint8, float
patches + owner keys only in the seam zone), one reduce, one horizontal and one
vertical coarse blur over every (patch, level) job, one output-stationary collapse
of the seam-zone tiles.  Per-patch parameters travel in small job tables.  Any
rows x columns window of the mosaic (columns on 64-pixel tile edges) can be
composited on its own, byte-identical to that part of the whole: the strips of the
multi-GPU path and the windows of the streamed end-to-end pipeline.

Data layout in HBM
------------------
* source image      u8 RGBX [h][w][4]        packed on device from the u8x3 upload (only the part the plan reads)
* sample LUT        f32 [256] per image      u8 -> float value (gain folded in)
* hat tables        f64 [h], [w]             shared by images of equal size
* ray tables        f64 [W], [W], [H]        proj2hom per mosaic column (x, z) / row (y); K*R per patch
* patch pool        f32 [ph][pw][4] RGBA + u8 [ph][pw] invalid, all patches back to back
* owner keys        u64 [H][W]               float_bits(alpha) << 32 | ~patch (seam zone only; atomicMax on the dense path)
* covered           u8  [H][W]               union of valid pixels
* coarse levels     f32 [h/f][w/f][4]        reduced (d2, d4), H-pass scratch and blurred
                                             f=2 (level 0) / f=4 (levels >= 1) images per patch
* mosaic            u8  [H][W][3]            the only mosaic-sized output

A *window* (rows ``(ya, yb)`` and / or columns ``(xa, xb)``) restricts all work to
that part of the mosaic plus a halo of the reach of the widest coarse blur and one
tile (strip sharding; streamed pipeline); patches are cropped to
the window + halo, far enough from it that neither the dropped pixels nor the
reflections at the artificial edges reach a pixel with non-zero weight.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _lib, geometry as geo


# Mosaic (or strip) size from which the seam-band maps are used.  Measured only at cfg4 (279 Mpix:
# 15.9 -> 13.2 ms); the five extra launches and the persistent grids are not free, so the 1-8 Mpix
# panoramas of cfg1 / cfg5 (0.2-0.5 ms per composite) keep the dense path until they are measured.
SEAM_MAPS_MIN_PIXELS = 1 << 24
# Rigs with an image side below this many pixels are blended at full resolution (see needs_exact).
EXACT_BELOW = 64


_copy_pool = None


def parallel_copy(dst, src, workers=16):
    """dst[...] = src for two large host arrays of equal shape, split by rows over a few threads
    (NumPy releases the GIL while it copies): the staging of pageable buffers runs at several
    times the rate of a single memcpy."""
    global _copy_pool
    n = dst.shape[0]
    if dst.nbytes < (4 << 20) or n < 2 * workers:
        np.copyto(dst, src)
        return
    if _copy_pool is None:
        from concurrent.futures import ThreadPoolExecutor
        _copy_pool = ThreadPoolExecutor(max(2, min(workers, os.cpu_count() or 2)))
    cuts = [n * k // workers for k in range(workers + 1)]
    list(_copy_pool.map(lambda ab: np.copyto(dst[ab[0]:ab[1]], src[ab[0]:ab[1]]), zip(cuts, cuts[1:])))


def band_edges(ya, yb, bands):
    """[ya, yb) cut into ``bands`` row ranges with integer arithmetic only, so
    that every rank derives identical cuts whatever its local row origin."""
    return [(ya + k * (yb - ya) // bands, ya + (k + 1) * (yb - ya) // bands) for k in range(bands)]


def _require_cuda(device):
    if not torch.cuda.is_available():
        raise RuntimeError("pano360_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError(f"pano360_b200 runs on CUDA devices only, got {dev}")
    return dev


class DevicePatch:
    """One warped image resident in HBM.  Unpacks like the reference's patch
    triple ``(warped, mask, irange)`` (stitcher.py:318-319).

    Patches produced by the warp live back to back in two pools; the tensor
    views ``rgba`` ([ph, pw, 4] float32) and ``invalid`` ([ph, pw] uint8, 1 =
    masked) are only materialised when somebody asks for them — the kernels
    work from the raw device addresses."""

    __slots__ = ("box", "index", "rgba_ptr", "invalid_ptr", "_rgba", "_invalid", "_pools", "_offset")

    def __init__(self, rgba=None, invalid=None, box=(0, 0, 0, 0), index=0, pools=None, offset=0, bases=None):
        self.box, self.index = tuple(box), index
        self._rgba, self._invalid, self._pools, self._offset = rgba, invalid, pools, offset
        if pools is None:
            self.rgba_ptr, self.invalid_ptr = rgba.data_ptr(), invalid.data_ptr()
        else:
            bases = bases or (pools[0].data_ptr(), pools[1].data_ptr())
            self.rgba_ptr = bases[0] + 16 * offset
            self.invalid_ptr = bases[1] + offset

    @property
    def shape(self):
        x0, y0, x1, y1 = self.box
        return y1 - y0, x1 - x0

    @property
    def rgba(self):
        if self._rgba is None:
            ph, pw = self.shape
            self._rgba = self._pools[0][4 * self._offset:4 * (self._offset + ph * pw)].view(ph, pw, 4)
        return self._rgba

    @property
    def invalid(self):
        if self._invalid is None:
            ph, pw = self.shape
            self._invalid = self._pools[1][self._offset:self._offset + ph * pw].view(ph, pw)
        return self._invalid

    @property
    def irange(self):
        x0, y0, x1, y1 = self.box
        return (slice(y0, y1), slice(x0, x1))

    def __iter__(self):
        return iter((self.rgba, self.invalid, self.irange))

    def to_numpy(self):
        return (self.rgba.cpu().numpy(), self.invalid.cpu().numpy().astype(bool), self.irange)


@dataclass
class DeviceSources:
    """Input images + per-image constants resident in HBM."""

    pixels: list                       # u8 [h, w, 4] tensors (None for images a rank does not need)
    luts: list                         # f32 [256] tensors
    hats: dict = field(default_factory=dict)   # (h, w) -> (hat_y, hat_x) f64 tensors
    shapes: list = field(default_factory=list)
    ready: list = None                 # per image CUDA event (uploads issued on the copy stream)
    rows: list = None                  # per image (r0, r1) or (r0, r1, c0, c1): only that part is resident / packed (None: all)
    bytes_up: int = 0                  # image bytes that crossed PCIe for this set
    issue: object = None               # upload(lazy=True): issue(count) copies the first `count` images of the order


class Compositor:
    """Runs warp / gain / blend stages on one GPU."""

    # What fresh coarse-pool memory holds, and whether it is refilled before every composite: the
    # batched blurs may stage cells no kernel of the composite wrote (skipped neighbours), which
    # is harmless exactly as long as those hold finite values.  Zeroed once in production; the
    # host build of the kernels (tests/emul) poisons with 1e30 every time.
    pool_fill = 0.0
    repoison = False

    def __init__(self, device=None):
        self.device = _require_cuda(device)
        _lib.load()
        self._pinned = {}
        self._pools = {}
        self._consts = {}
        self._packed = {}      # raw image addresses -> the buffers their RGBX copies go to
        self._images = {}      # (slot, shape) -> device buffers of uploaded images (upload(reuse=True))
        self._prepared = {}    # what a composite of one geometry over one set of buffers needs, kept ready
        self.prepared_max = 6  # ... for at most this many (geometry, window) combinations
        self._taps_key = None
        self._keep = {}
        self._copy = None      # side streams for uploads / downloads that overlap the kernels
        self._down = None
        self._download = None
        self._bands_down = []  # (y0, y1, event, x0, x1) of the banded download in flight
        self._drain = None     # (queue, thread, errors) while a copy-out thread moves landed bands to the caller's array
        self._ring = []        # pinned staging slots for pageable inputs: [tensor, busy event]
        self._ring_at = 0
        self._out_stage = None  # pinned staging for a pageable output
        self._ingest_stage = None  # pinned landing buffer of ingest.resize_on_device
        self.stage_min_bytes = 1 << 20   # pageable images from this size on go through the pinned ring
        self.trace = None      # list of (kernel, algorithmic_bytes, start_event, end_event) when enabled
        self.phases = None     # host-side phase times of stitch_strips calls when enabled (bench.py)
        self.timeline = None   # list of (label, event) across the upload / compute / download streams when enabled
        # seam-band maps (p360_tile_maps_build): reduce / blur only where two owners meet within the
        # blur reach.  Bit-identical output either way; None = on for mosaics large enough for the
        # extra launches to pay (B200, cfg4: 15.9 -> 13.2 ms), P360_SEAM_MAPS=0/1 forces it.
        self.seam_maps = {"0": False, "1": True}.get(os.environ.get("P360_SEAM_MAPS", ""))
        # horizontal blur of the block lists in 64-cell instead of 256-cell segments (4 rows per
        # warp): halves the cells run at cfg4 (tools/seam_map_stats.py).  B200, cfg4: K3 2.56 ->
        # 1.53 ms, byte-identical (profiles/r02_probe_switches.log); P360_BLUR_H_ROWS=1 for the old lists
        self.blur_h_rows = 1 if os.environ.get("P360_BLUR_H_ROWS", "4") == "1" else 4
        # seam plan (p360_seam_plan_build): ownership is geometric, so before anything is sampled the
        # mosaic tiles are split into solo tiles — written straight from the sources as uint8
        # (p360_warp_tiles) — and the seam zone, the only place where float patches, owner keys and
        # coarse levels exist.  P360_DIRECT=0: every patch warped to float, maps from the owner keys.
        self.direct = os.environ.get("P360_DIRECT", "1") == "1"
        # upload only the rectangle of every image the seam plan can sample (``source_rects``); P360_SOURCE_RECTS=0: whole images
        self.partial_uploads = os.environ.get("P360_SOURCE_RECTS", "1") == "1"
        self.last_covered = None
        self.last_upload_bytes = 0     # image bytes the last ``upload`` sent over PCIe
        self.used_after_warp = False   # did the last composite call its ``after_warp`` hook

    # -- plumbing -----------------------------------------------------------
    @property
    def stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _to_device(self, array, pinned_key=None):
        """Host ndarray -> device tensor, optionally through a reused pinned
        staging buffer (async copy)."""
        array = np.ascontiguousarray(array)
        host = torch.from_numpy(array)
        if pinned_key is not None:
            stage, busy = self._pinned.get(pinned_key, (None, None))
            if busy is not None:
                busy.synchronize()            # previous async copy out of this buffer is done
            if stage is None or stage.numel() < host.numel() or stage.dtype != host.dtype:
                stage = torch.empty(max(host.numel(), 1), dtype=host.dtype, pin_memory=True)
            view = stage[:host.numel()].view(host.shape)
            view.copy_(host)
            dev = view.to(self.device, non_blocking=True)
            busy = torch.cuda.Event()
            busy.record(torch.cuda.current_stream(self.device))
            self._pinned[pinned_key] = (stage, busy)
            return dev
        return host.to(self.device)

    def _mark(self, label, stream=None):
        """Timeline mark (tools/e2e_probe.py): a timing event on ``stream`` (default: the compute stream)."""
        if self.timeline is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(stream if stream is not None else torch.cuda.current_stream(self.device))
            self.timeline.append((label, ev))

    def copy_stream(self):
        """Side stream of the uploads (and of the NVLink pushes of strips.py)."""
        if self._copy is None:
            self._copy = torch.cuda.Stream(self.device)
        return self._copy

    def download_stream(self):
        """Side stream of the banded mosaic downloads: its own stream, so that device -> host
        copies are not queued behind uploads still in flight (PCIe is full duplex).  With
        P360_DOWN_STREAMS=2 consecutive calls alternate between two streams (two copy engines
        share the device -> host direction)."""
        if self._down is None:
            self._down = [torch.cuda.Stream(self.device) for _ in range(max(1, int(os.environ.get("P360_DOWN_STREAMS", "1"))))]
            self._down_at = 0
        self._down_at = (self._down_at + 1) % len(self._down)
        return self._down[self._down_at]

    def _table(self, records, key):
        """Structured job table -> device bytes."""
        return self._to_device(records.view(np.uint8).reshape(-1), pinned_key=key)

    def _traced(self, name, nbytes, fn, *args):
        """Run one C-ABI call; when tracing, bracket it with CUDA events on the
        launching stream (bench.py reads per-kernel time from these)."""
        if self.trace is None:
            return _lib.call(fn, *args)
        stream = torch.cuda.current_stream(self.device)
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record(stream)
        rc = _lib.call(fn, *args)
        end.record(stream)
        self.trace.append((name, nbytes, start, end))
        return rc

    # -- sources --------------------------------------------------------------
    def pack_pixels(self, dev_img, rows=None, out=None):
        """u8 x 3 -> u8 x 4 (RGBX) on the device: a bilinear tap of the warp is then one aligned
        32-bit load.  4-channel images are used as they are.  ``rows = (r0, r1)`` or ``(r0, r1, c0,
        c1)``: only those rows (that rectangle) hold data, and only they are converted."""
        h, w, c = dev_img.shape
        if c == 4:
            return dev_img
        packed = out if out is not None else torch.empty((h, w, 4), dtype=torch.uint8, device=self.device)
        r0, r1 = (0, h) if rows is None else rows[:2]
        c0, c1 = (0, w) if rows is None or len(rows) < 4 else rows[2:]
        if (c0, c1) != (0, w):
            c0 = c0 // 4 * 4
            if r1 > r0 and c1 > c0:
                self._traced("K1p_pack_rgbx", 7 * (r1 - r0) * (c1 - c0), "p360_pack_rgbx_rect", dev_img.data_ptr(), h, w,
                             r0, r1, c0, c1, packed.data_ptr(), self.stream)
            return packed
        r0 = r0 // 4 * 4                                   # keeps source and destination addresses aligned
        if r1 > r0:
            self._traced("K1p_pack_rgbx", 7 * (r1 - r0) * w, "p360_pack_rgbx", dev_img.data_ptr() + 3 * w * r0, r1 - r0, w,
                         packed.data_ptr() + 4 * w * r0, self.stream)
        return packed

    def pack_sources(self, raw):
        """A ``DeviceSources`` whose images are in the warp's RGBX layout, from one holding the
        images as uploaded (``upload(pack=False)``): the device-side part of ``_add_weights``
        (stitcher.py:257-263) that is executed once per image and stitch — ONE launch for all
        the images (p360_pack_rgbx_batch), each converting only the part that was uploaded."""
        rows = raw.rows or [None] * len(raw.pixels)
        # the packed copies of one set of resident images always land in the same buffers: stable
        # addresses let ``composite`` reuse everything it prepared for them
        key = tuple(0 if p is None else (p.data_ptr(), tuple(p.shape)) for p in raw.pixels) + (tuple(rows),)
        entry = self._packed.get(key)
        if entry is None:
            if len(self._packed) >= 8:
                self._packed.pop(next(iter(self._packed)))
            outs, jobs = [None] * len(raw.pixels), []
            for i, (p, r) in enumerate(zip(raw.pixels, rows)):
                if p is None or p.shape[2] == 4:
                    outs[i] = p
                    continue
                h, w = p.shape[:2]
                outs[i] = torch.empty((h, w, 4), dtype=torch.uint8, device=self.device)
                r0, r1 = (0, h) if r is None else r[:2]
                c0, c1 = (0, w) if r is None or len(r) < 4 else r[2:]
                jobs.append((p.data_ptr(), outs[i].data_ptr(), h, w, r0, r1, c0 // 4 * 4, c1))
            table = np.array(jobs, dtype=_lib.PACK_JOB) if jobs else np.zeros(0, dtype=_lib.PACK_JOB)
            entry = self._packed[key] = {
                "outs": outs, "n": len(jobs), "jobs": self._to_device(table.view(np.uint8).reshape(-1)) if jobs else None,
                "rows": int((table["r1"] - table["r0"]).max()) if jobs else 0,
                "cols": int((table["c1"] - table["c0"]).max()) if jobs else 0,
                "bytes": int(7 * ((table["r1"] - table["r0"]).astype(np.int64) * (table["c1"] - table["c0"])).sum()) if jobs else 0}
        if entry["n"]:
            self._traced("K1p_pack_rgbx", entry["bytes"], "p360_pack_rgbx_batch", _lib.ptr(entry["jobs"]), entry["n"],
                         entry["rows"], entry["cols"], self.stream)
        return DeviceSources(list(entry["outs"]), raw.luts, raw.hats, raw.shapes, raw.ready, raw.rows)

    def upload(self, regions, gains=None, need=None, overlap=False, pack=True, order=None, rows_of=None, reuse=False,
               rects_of=None, lazy=False):
        """H2D copy of the u8 images (+ LUT / hat tables).  Images backed by
        pinned memory are copied asynchronously.  ``need`` (a set of indices)
        restricts the copy to the images a rank's strip touches.  With
        ``overlap`` the copies (and the RGBX packing) run on a side stream and
        every image gets a ``ready`` event, so the warp of the first images
        starts while the last ones are still crossing PCIe.  ``order`` (a
        permutation of the indices) is the order in which the copies are issued.  ``rows_of``
        ({image: (r0, r1)}, ``source_rows``) uploads only the rows a composite will read;
        ``rects_of`` ({image: (r0, r1, c0, c1)}, ``source_rects``) only that rectangle.  ``bytes_up``
        of the result counts what crossed PCIe.
        ``reuse``: the images go to device buffers kept from the previous such call (same slot,
        same shape) — for callers that drop the result before they upload again (``stitch``):
        stable addresses let ``composite`` reuse what it prepared.
        ``lazy``: only the device buffers are set up (their addresses are final); the copies are
        issued by ``src.issue(count)`` — the first ``count`` images of ``order`` — so that a caller
        can interleave the staging of pageable images with the work that consumes the earlier ones."""
        n = len(regions)
        src = DeviceSources([None] * n, [None] * n)
        if rects_of is not None:
            rows_of = rects_of
        if rows_of is not None:
            src.rows = [rows_of.get(i) for i in range(n)]
        src.shapes = [reg.img.shape[:2] for reg in regions]
        src.bytes_up = 0
        lut0 = None
        main = torch.cuda.current_stream(self.device)
        side = self.copy_stream() if overlap else main
        self._mark("upload begins")
        if overlap:
            src.ready = [None] * n
            side.wait_stream(main)
        sequence = [i for i in (range(n) if order is None else order) if need is None or i in need]
        todo = {}
        for i in range(n):
            if gains is None:
                lut0 = self._constant("lut0", lambda: geo.sample_lut(None)) if lut0 is None else lut0
                src.luts[i] = lut0
            else:
                src.luts[i] = self._to_device(geo.sample_lut(gains[i]))
        # ---- phase 1: where every image goes (device buffers, the part of it that is copied)
        for i in sequence:
            img = regions[i].img
            h, w = img.shape[:2]
            if img.dtype != np.uint8 or img.ndim != 3 or img.shape[2] not in (3, 4):
                raise TypeError("region images must be uint8 HxWx3 (what the reference accepts)")
            host = torch.from_numpy(np.ascontiguousarray(img))
            part = None if rows_of is None else rows_of.get(i)
            rect, whole = False, host
            if part is not None and len(part) == 4 and (part[2] > 0 or part[3] < w):
                part = (part[0], part[1], part[2] // 4 * 4, part[3])     # (the packing converts 4 pixels per thread)
                rect = True
                host = host[part[0]:part[1], part[2]:part[3]]          # a strided view
            elif part is not None:
                part = (part[0] // 4 * 4, part[1])        # (4-row granularity: aligned addresses for the packing)
                host = host[part[0]:part[1]]
            if (h, w) not in src.hats:
                fresh = ("hat", h) not in self._consts or ("hat", w) not in self._consts
                src.hats[(h, w)] = (self._constant(("hat", h), lambda: geo.hat(h)),
                                    self._constant(("hat", w), lambda: geo.hat(w)))
                if overlap and fresh:
                    side.wait_stream(main)               # hat tables were copied on the main stream
            dev_img = None
            if reuse or rect or part is not None or lazy:
                slot = (n, i, tuple(img.shape))
                dev_img = self._images.get(slot) if reuse else None
                if dev_img is None:                      # parts outside `part` stay unwritten: nobody reads them
                    dev_img = torch.empty(img.shape, dtype=torch.uint8, device=self.device)
                    if reuse:
                        self._images[slot] = dev_img
            out = None
            if pack and img.shape[2] != 4 and (reuse or lazy):
                slot = (n, i, tuple(img.shape), "rgbx")
                out = self._images.get(slot) if reuse else None
                if out is None:
                    out = torch.empty(img.shape[:2] + (4,), dtype=torch.uint8, device=self.device)
                    if reuse:
                        self._images[slot] = out
            if lazy:                                     # final addresses before any byte has moved
                src.pixels[i] = dev_img if (not pack or img.shape[2] == 4) else out
            todo[i] = (host, whole, part, rect, dev_img, out)

        # ---- phase 2: the copies (+ RGBX packing) of one image
        def put(i):
            host, whole, part, rect, dev_img, out = todo.pop(i)
            img = regions[i].img
            w = img.shape[1]
            src.bytes_up += host.numel()
            pinned = (whole if rect else host).is_pinned()
            if not pinned and (host.numel() >= self.stage_min_bytes or rect):
                host = self._stage_pageable(host, side)      # -> a pinned slot of the ring (async copy below)
            with torch.cuda.stream(side):
                if rect:           # the rectangle lands at its place in the full-size device image
                    r0, r1, c0, c1 = part
                    ch = img.shape[2]
                    _lib.call("p360_copy_rect", dev_img.data_ptr() + ch * (r0 * w + c0), ch * w, host.data_ptr(),
                              host.stride(0), ch * (c1 - c0), r1 - r0, side.cuda_stream)
                elif dev_img is not None:
                    target = dev_img if part is None else dev_img[part[0]:part[1]]
                    target.copy_(host, non_blocking=host.is_pinned())
                else:
                    dev_img = host.to(self.device, non_blocking=host.is_pinned())
                if getattr(host, "_p360_slot", None) is not None:
                    busy = torch.cuda.Event()
                    busy.record(side)
                    host._p360_slot[1] = busy
                # pack=False keeps the uploaded u8 x 3 layout (three byte loads per tap)
                src.pixels[i] = self.pack_pixels(dev_img, part, out=out) if pack else dev_img
                if overlap:
                    src.ready[i] = torch.cuda.Event()
                    src.ready[i].record(side)
                self._mark(f"image {i} uploaded + packed", side)

        state = {"at": 0}
        positions = list(range(n) if order is None else order)

        def issue(count=None):
            """copy the images at the first ``count`` positions of ``order`` (all: None) that are still due"""
            upto = n if count is None else min(count, n)
            while state["at"] < upto:
                if positions[state["at"]] in todo:
                    put(positions[state["at"]])
                state["at"] += 1
            self.last_upload_bytes = src.bytes_up
        src.issue = issue
        if not lazy:
            issue()
        return src

    def _constant(self, key, make):
        """Small per-device tables that never change (the u8 -> float LUT without gains, the hat
        tables of an image size): uploaded once, same address ever after."""
        if key not in self._consts:
            self._consts[key] = self._to_device(make())
        return self._consts[key]

    def _stage_pageable(self, host, side, slots=3):
        """Copy a pageable host image into the next slot of a small ring of pinned buffers (a few
        threads share the memcpy) so that its upload is an asynchronous DMA: the copy of image
        k + 1 into its slot overlaps the DMA of image k."""
        if len(self._ring) < slots:
            self._ring.append([torch.empty(0, dtype=torch.uint8, pin_memory=True), None])
        slot = self._ring[self._ring_at % len(self._ring)]
        self._ring_at += 1
        if slot[1] is not None:
            slot[1].synchronize()                    # the DMA out of this slot has finished
        if slot[0].numel() < host.numel():
            slot[0] = torch.empty(host.numel(), dtype=torch.uint8, pin_memory=True)
        view = slot[0][:host.numel()].view(host.shape)
        parallel_copy(view.numpy(), host.numpy())
        view._p360_slot = slot
        return view

    def host_stage(self, shape):
        """Pinned staging buffer for a mosaic that goes to pageable host memory (kept, grow-only)."""
        n = int(np.prod(shape))
        if self._out_stage is None or self._out_stage.numel() < n:
            self._out_stage = torch.empty(n, dtype=torch.uint8, pin_memory=True)
        return self._out_stage[:n].view(tuple(shape)).numpy()

    def set_gains(self, src, gains):
        src.luts = [self._to_device(geo.sample_lut(g)) for g in gains]

    # -- K8 + host solve: exposure gains (stitcher.py:24-66) ------------------
    def pair_statistics(self, regions, src, pairs=None):
        """overlaps / sizes matrices of ``equalize_gains`` from device sums."""
        n = len(regions)
        h, w = src.shapes[0]
        overlaps, sizes = np.zeros((n, n)), np.zeros((n, n))
        todo = []
        for i in range(n):
            for j in range(i + 1, n):
                hom, behind = geo.pair_homography(regions[i], regions[j], (h, w))
                if not behind:
                    todo.append((i, j, geo.invert3x3(hom)))
        if pairs is not None:
            todo = [t for k, t in enumerate(todo) if k in pairs]
        if not todo:
            return overlaps, sizes, todo
        lut0 = self._constant("lut0", lambda: geo.sample_lut(None))
        hat_y, hat_x = src.hats[(h, w)]
        nblocks = _lib.call("p360_pair_stats_blocks", h, w)
        jobs = np.zeros(len(todo), dtype=_lib.PAIR_JOB)
        for k, (i, j, inv) in enumerate(todo):
            if src.shapes[i] != (h, w) or src.shapes[j] != (h, w):
                raise ValueError("exposure equalisation needs equally sized images (as the reference)")
            jobs[k] = (src.pixels[i].data_ptr(), src.pixels[j].data_ptr(), inv.ravel())
        dev_jobs = self._table(jobs, "pair_jobs")
        partial = torch.empty(3 * nblocks * len(todo), dtype=torch.float64, device=self.device)
        out = torch.empty((len(todo), 3), dtype=torch.float64, device=self.device)
        self._traced("K8_pair_overlap_stats", 6 * h * w * len(todo), "p360_pair_overlap_stats",
                     _lib.ptr(dev_jobs), len(todo), h, w, src.pixels[todo[0][0]].shape[2], _lib.ptr(lut0),
                     _lib.ptr(hat_y), _lib.ptr(hat_x), _lib.ptr(partial), _lib.ptr(out), self.stream)
        sums = out.cpu().numpy()
        for (i, j, _), (cnt, s_i, s_j) in zip(todo, sums):
            sizes[i, j] = sizes[j, i] = cnt
            if cnt > 0:
                overlaps[i, j] = s_i / (3.0 * cnt)
                overlaps[j, i] = s_j / (3.0 * cnt)
        return overlaps, sizes, todo

    # -- K1: warp -------------------------------------------------------------
    def plan_crops(self, regions, plan, proj=geo.SphProj, rows=None, row_align=1, split_dilate=None, cols=None,
                   col_align=1):
        """Host side of the warp: which (window-cropped, column-split) boxes get
        warped.  ``rows=(ya, yb)`` / ``cols=(xa, xb)`` crop boxes to those mosaic rows / columns; a
        cropped top (left) edge is moved up (left) to a multiple of ``row_align`` (``col_align``)
        pixels from the box's true edge so that coarse grids anchored at the crop coincide with
        those anchored at the true box.  With ``split_dilate`` (columns) the
        all-invalid middle of seam-straddling boxes is dropped
        (``geometry.active_column_runs``): such an image yields two crops with
        the same image index.
        Returns (crops, tables): crops = [(image, x0, y0, x1, y1, K*R, true y0, true y1, true x0,
        true x1)] (the true columns are those of the run), tables = the per-mosaic-column /
        per-row ray tables of the projection."""
        key = (len(regions), proj, rows, row_align, split_dilate, cols, col_align)   # (a plan belongs to one rig geometry)
        cached = plan._crops.get(key) if plan._crops is not None else None
        if cached is not None:                       # same regions, same plan, same window as last time
            return cached, plan.rays(proj)
        crops = []
        for i, (reg, box) in enumerate(zip(regions, plan.boxes)):
            x0, y0, x1, y1 = box
            ya, yb = (y0, y1) if rows is None else (max(y0, rows[0]), min(y1, rows[1]))
            if ya >= yb or x0 >= x1:
                continue
            ya = y0 + (ya - y0) // row_align * row_align
            runs = [(x0, x1)] if split_dilate is None else \
                geo.active_column_runs(i, box, plan, dilate=split_dilate)
            k_r = np.ascontiguousarray(reg.proj(), dtype=np.float64).ravel()
            for rx0, rx1 in runs:
                cx0, cx1 = (rx0, rx1) if cols is None else (max(rx0, cols[0]), min(rx1, cols[1]))
                if cx0 >= cx1:
                    continue
                cx0 = rx0 + (cx0 - rx0) // col_align * col_align
                crops.append((i, cx0, ya, cx1, yb, k_r, y0, y1, rx0, rx1))
        if plan._crops is not None:
            plan._crops[key] = crops
        return crops, plan.rays(proj)

    def source_rows(self, regions, plan, kind, n_levels=5, proj=geo.SphProj, rows=None, cols=None):
        """{image: (r0, r1)}: the source rows ``composite(..., rows=rows, cols=cols)`` can touch — for
        ``upload(rows_of=...)``.  Images the composite does not meet are absent.  (Interval
        arithmetic over the cropped boxes on the host: works for every blender; ``source_rects``
        is tighter where the seam plan applies.)"""
        key = ("rows", len(regions), proj, rows, cols, kind, n_levels)
        cached = plan._crops.get(key) if plan._crops is not None else None
        if cached is not None:
            return cached
        crops = self._window_geometry(regions, plan, kind, n_levels, proj, rows, cols)[0]
        needed = {}
        for c in crops:
            r0, r1 = geo.source_rows_needed(regions[c[0]], c[1:5], plan, proj)
            if r1 > r0:
                old = needed.get(c[0])
                needed[c[0]] = (r0, r1) if old is None else (min(r0, old[0]), max(r1, old[1]))
        if plan._crops is not None:
            plan._crops[key] = needed
        return needed

    def source_rects(self, regions, plan, kind, n_levels=5, proj=geo.SphProj, rows=None, cols=None):
        """{image: (r0, r1, c0, c1)}: the rectangle of every image that ``composite`` (of the whole
        mosaic, or of the given window) can sample — for ``upload(rects_of=...)``.  From the seam
        plan itself (K0 + K0s, before any pixel exists on the device): an image is only read where it
        is the single candidate of a tile or takes part in a seam.  The rectangles of the whole
        mosaic also cover every window of it (a window's plan is a subset of the whole plan).
        None when the composite would not go through the seam plan (other blenders, the dense
        warp, the full-resolution path): then everything may be read."""
        if not (self.direct and self.partial_uploads and kind == "multiband" and n_levels > 1) or self.needs_exact(regions):
            return None
        key = ("rects", len(regions), proj, rows, cols, kind, n_levels)
        if plan._crops is not None and key in plan._crops:
            return plan._crops[key]
        crops, tables, top, left, shape, _, _ = self._window_geometry(regions, plan, kind, n_levels, proj, rows, cols)
        rects = None
        if 0 < len(crops) <= 256:
            shapes = {c[0]: regions[c[0]].img.shape[:2] for c in crops}
            jobs, _, keep = self._warp_jobs(None, crops, tables, origin=(left, top), pools=False, shapes=shapes)
            n = len(jobs)
            pad, bands = geo.coarse_band_plan(n_levels)
            table = np.zeros(n, dtype=_lib.BAND_PATCH)
            table["w4"], table["h4"] = (jobs["pw"] + 2 * pad + 3) // 4, (jobs["ph"] + 2 * pad + 3) // 4
            maps, maps_keep = self._tile_maps(table, len(bands), shape[0], shape[1], pad, top, seam_plan=True)
            dev_jobs = self._to_device(jobs.view(np.uint8).reshape(-1))
            big, small = np.iinfo(np.int32).max, np.iinfo(np.int32).min
            dev_rects = torch.tensor([[big, big, small, small] * 2] * n, dtype=torch.int32, device=self.device)
            _lib.call("p360_seam_plan_build", _lib.ptr(dev_jobs), n, None, shape[0], shape[1], top, plan.shape[0],
                      maps.ctypes.data, self.stream)
            _lib.call("p360_source_rects", _lib.ptr(dev_jobs), n, shape[0], shape[1], top, plan.shape[0],
                      maps.ctypes.data, _lib.ptr(dev_rects), self.stream)
            found = dev_rects.cpu().numpy()
            del maps_keep, keep
            rects, used = {}, {}
            for c, (u0, v0, u1, v1, x0, y0, x1, y1) in zip(crops, found.tolist()):
                h, w = shapes[c[0]]
                if u1 <= u0 or v1 <= v0:
                    u0, v0, u1, v1 = 0, 0, min(4, w), min(4, h)       # never sampled: a token corner keeps the tables uniform
                else:
                    used.setdefault(c[0], []).append((x0 + left, y0 + top, x1 + left, y1 + top))
                old = rects.get(c[0])
                rects[c[0]] = (v0, v1, u0, u1) if old is None else (min(v0, old[0]), max(v1, old[1]), min(u0, old[2]), max(u1, old[3]))
            if plan._crops is not None:
                plan._crops[("used",) + key[1:]] = used
        if plan._crops is not None:
            plan._crops[key] = rects
        return rects

    def used_boxes(self, regions, plan, kind, n_levels=5, proj=geo.SphProj):
        """{image: [(x0, y0, x1, y1), ...]}: per column run of every image, the box (mosaic pixels)
        around the tiles it is read for (from ``source_rects``): a window of the mosaic whose
        buffer does not meet any of them does not need the image.  None without the seam plan."""
        if self.source_rects(regions, plan, kind, n_levels, proj) is None:
            return None
        return plan._crops.get(("used", len(regions), proj, None, None, kind, n_levels))

    def new_owner_state(self, shape):
        """(owner keys u64, covered u8) for a mosaic (or strip) of ``shape``."""
        h, w = shape
        return (torch.zeros((h, w), dtype=torch.int64, device=self.device),
                torch.zeros((h, w), dtype=torch.uint8, device=self.device))

    def _warp_jobs(self, src, crops, tables, origin=(0, 0), pools=True, shapes=None):
        """Host side of K1: the job table (one p360_warp_job per crop), the patch pools and the
        patches.  Returns (jobs, patches, keep) — ``keep`` holds what the launches reference.
        ``pools=False`` (with ``src=None`` and the image ``shapes``): the geometry of the jobs only,
        for the seam plan (no patch memory, no source addresses)."""
        ray_x, ray_z, ray_y = tables
        cached = self._keep.get("rays")
        if cached is not None and cached[0] is ray_x:          # same plan as last time: tables already on the device
            dev_rays = cached[1]
        else:
            dev_rays = self._to_device(np.concatenate([ray_x, ray_z, ray_y]), pinned_key="rays")
            self._keep["rays"] = (ray_x, dev_rays)
        base_x = dev_rays.data_ptr()
        base_z, base_y = base_x + 8 * len(ray_x), base_x + 8 * (len(ray_x) + len(ray_z))
        ox, oy = origin
        n = len(crops)
        sizes = np.array([(c[3] - c[1]) * (c[4] - c[2]) for c in crops], dtype=np.int64)
        offs = np.concatenate([[0], np.cumsum(sizes)])
        if pools:
            rgba_pool = torch.empty(int(offs[-1]) * 4, dtype=torch.float32, device=self.device)
            inv_pool = torch.empty(int(offs[-1]), dtype=torch.uint8, device=self.device)
            rgba_base, inv_base = rgba_pool.data_ptr(), inv_pool.data_ptr()
        else:
            rgba_pool = inv_pool = None
            rgba_base = inv_base = 0
        # the job table column by column (per-image constants looked up once per image, not per crop)
        image = np.array([c[0] for c in crops], dtype=np.int64)
        box = np.array([c[1:5] for c in crops], dtype=np.int64).reshape(n, 4)          # x0, ya, x1, yb
        true_rows = np.array([c[6:8] for c in crops], dtype=np.int64).reshape(n, 2)
        true_cols = np.array([c[8:10] for c in crops], dtype=np.int64).reshape(n, 2)
        per_image = {}
        for i in set(image.tolist()):
            if src is None:
                h, w = shapes[i]
                per_image[i] = (0, 0, 0, 0, h, w, 4)
                continue
            h, w = src.shapes[i]
            hat_y, hat_x = src.hats[(h, w)]
            pix = src.pixels[i]
            per_image[i] = (pix.data_ptr(), src.luts[i].data_ptr(), hat_y.data_ptr(), hat_x.data_ptr(), h, w, pix.shape[2])
        consts = np.array([per_image[i] for i in image.tolist()], dtype=np.uint64).reshape(n, 7)
        hs, ws = consts[:, 4].astype(np.int64), consts[:, 5].astype(np.int64)
        jobs = np.zeros(n, dtype=_lib.WARP_JOB)
        jobs["src"], jobs["lut"], jobs["hat_y"], jobs["hat_x"] = consts[:, 0], consts[:, 1], consts[:, 2], consts[:, 3]
        jobs["ray_x"], jobs["ray_z"], jobs["ray_y"] = base_x, base_z, base_y
        jobs["out"] = np.uint64(rgba_base) + np.uint64(16) * offs[:-1].astype(np.uint64)
        jobs["invalid"] = np.uint64(inv_base) + offs[:-1].astype(np.uint64)
        jobs["kr"] = np.stack([c[5] for c in crops])
        jobs["h"], jobs["w"], jobs["c"] = hs, ws, consts[:, 6]
        jobs["pw"], jobs["ph"] = box[:, 2] - box[:, 0], box[:, 3] - box[:, 1]
        jobs["x0"], jobs["y0"], jobs["col0"], jobs["row0"] = box[:, 0] - ox, box[:, 1] - oy, box[:, 0], box[:, 1]
        jobs["ty0"], jobs["ty1"] = true_rows[:, 0] - oy, true_rows[:, 1] - oy
        jobs["tx0"], jobs["tx1"] = true_cols[:, 0] - ox, true_cols[:, 1] - ox
        jobs["patch"] = np.arange(n)
        jobs["half_w"], jobs["half_h"] = (ws / 2).astype(np.float32), (hs / 2).astype(np.float32)
        jobs["max_x"], jobs["max_y"] = (ws - 1).astype(np.float32), (hs - 1).astype(np.float32)
        jobs["inv_2w"] = np.float32(1.0) / (2 * ws).astype(np.float32)
        jobs["inv_2h"] = np.float32(1.0) / (2 * hs).astype(np.float32)
        if not pools:
            return jobs, None, (dev_rays, None, None, int(offs[-1]))
        pools, bases = (rgba_pool, inv_pool), (rgba_base, inv_base)
        patches = [DevicePatch(box=(b[0] - ox, b[1] - oy, b[2] - ox, b[3] - oy), index=i, pools=pools, offset=o, bases=bases)
                   for i, b, o in zip(image.tolist(), box.tolist(), offs[:-1].tolist())]
        return jobs, patches, (dev_rays, rgba_pool, inv_pool, int(offs[-1]))

    def _launch_warp(self, src, crops, jobs, keys, covered, width, per_px, pixels):
        """K1 over the job table: one launch, or — while uploads are still in flight — one per
        group of images as they arrive."""
        n = len(jobs)
        dev_jobs = self._table(jobs, "warp_jobs")        # (constant memory is staged from this copy: no stream sync)
        base = dev_jobs.data_ptr()
        self._keep["warp_jobs"] = dev_jobs
        if src.ready is None:
            self._traced("K1_warp", per_px * pixels, "p360_warp_batch", jobs.ctypes.data, base, n,
                         _lib.ptr(keys), _lib.ptr(covered), width, self.stream)
            return
        main = torch.cuda.current_stream(self.device)
        group = max(1, -(-len(src.ready) // 8))
        a = 0
        while a < n:
            last = (crops[a][0] // group + 1) * group - 1          # last image of this group
            b = a
            while b < n and crops[b][0] <= last:
                b += 1
            for i in sorted({c[0] for c in crops[a:b]}):        # uploads may be issued in any order
                main.wait_event(src.ready[i])
            _lib.call("p360_warp_batch", jobs.ctypes.data + a * _lib.WARP_JOB.itemsize,
                      base + a * _lib.WARP_JOB.itemsize, b - a, _lib.ptr(keys), _lib.ptr(covered), width, self.stream)
            a = b

    def warp_crops(self, src, crops, tables, origin=(0, 0), owner_state=None):
        """K1 over every crop in ONE launch.  Boxes of the returned patches
        are relative to ``origin`` (x, y).  With ``owner_state = (keys,
        covered)`` (mosaic-sized, zeroed) the owner-map competition is fused
        into the warp; patch k of the returned list is known as k there."""
        if not crops:
            return []
        jobs, patches, keep = self._warp_jobs(src, crops, tables, origin)
        if owner_state is None:
            keys = covered = None
            width, per_px = 0, 17
        else:
            keys, covered = owner_state
            width, per_px = keys.shape[1], 30
        self._launch_warp(src, crops, jobs, keys, covered, width, per_px, keep[3])
        self._keep["warp"] = keep[:3] + (jobs,)
        return patches

    def warp(self, regions, src, plan, proj=geo.SphProj, rows=None, row_align=1, split_dilate=None):
        """plan_crops + warp_crops in absolute mosaic coordinates."""
        crops, tables = self.plan_crops(regions, plan, proj, rows, row_align, split_dilate)
        return self.warp_crops(src, crops, tables)

    # -- blenders (device-resident patches in, device u8 mosaic out) ---------
    @staticmethod
    def _args(p):
        x0, y0, x1, y1 = p.box
        return x1 - x0, y1 - y0, x0, y0

    def owner_state_for(self, patches, shape):
        """Owner keys + covered for patches that were not warped by us (K2)."""
        keys, covered = self.new_owner_state(shape)
        for k, p in enumerate(patches):
            pw, ph, x0, y0 = self._args(p)
            self._traced("K2_owner_update", 30 * pw * ph, "p360_owner_update", p.rgba_ptr,
                         p.invalid_ptr, pw, ph, x0, y0, k, _lib.ptr(keys), _lib.ptr(covered),
                         shape[1], self.stream)
        return keys, covered

    def owner_map(self, patches, shape, owner_state=None):
        """stitcher.py:196-204 without the H x W x N tensor: (owner int32, covered)."""
        keys, covered = owner_state if owner_state is not None else self.owner_state_for(patches, shape)
        owner = torch.empty(shape, dtype=torch.int32, device=self.device)
        _lib.call("p360_owner_decode", _lib.ptr(keys), _lib.ptr(owner), shape[0] * shape[1], self.stream)
        return owner, covered

    def blur_taps(self, rgba, taps, out=None, tmp=None):
        """Separable convolution of a device RGBA image with explicit taps."""
        out = torch.empty_like(rgba) if out is None else out
        tmp = torch.empty_like(rgba) if tmp is None else tmp
        ph, pw = rgba.shape[:2]
        taps = np.ascontiguousarray(taps, dtype=np.float32)
        self._traced("K3_gauss_blur", 32 * pw * ph, "p360_gauss_blur", _lib.ptr(rgba), _lib.ptr(out),
                     _lib.ptr(tmp), pw, ph, taps.ctypes.data_as(C.POINTER(C.c_float)), len(taps),
                     self.stream)
        return out

    def blur(self, rgba, sigma, out=None, tmp=None):
        """cv2.GaussianBlur(rgba, (0, 0), sigma) on a device image."""
        return self.blur_taps(rgba, geo.gaussian_taps(sigma), out, tmp)

    def _band_table(self, patches, pad=0, coarse=False):
        """p360_band_patch records (+ coarse-grid sizes) for a patch list."""
        n = len(patches)
        table = np.zeros(n, dtype=_lib.BAND_PATCH)
        boxes = np.array([p.box for p in patches], dtype=np.int64).reshape(n, 4)      # column-wise: no per-record Python
        table["rgba"] = [p.rgba_ptr for p in patches]
        table["invalid"] = [p.invalid_ptr for p in patches]
        table["x0"], table["y0"] = boxes[:, 0], boxes[:, 1]
        table["pw"], table["ph"] = boxes[:, 2] - boxes[:, 0], boxes[:, 3] - boxes[:, 1]
        table["pad"], table["index"] = pad, np.arange(n)
        table["own"] = (2 ** 31 - 1, 2 ** 31 - 1, -2 ** 31, -2 ** 31)                  # grown on the device (p360_owned_boxes)
        if coarse:
            table["w4"], table["h4"] = (table["pw"] + 2 * pad + 3) // 4, (table["ph"] + 2 * pad + 3) // 4
        return table

    def _tile_maps(self, table, n_blurs, h, w, pad, row_origin, seam_plan=False):
        """Device bitmaps of p360_tile_maps (one bit per patch and 64 x 32 tile, tile rows on
        absolute mosaic rows), the scratch for the compacted block lists of reduce / blur, and the
        host record that names them."""
        n = len(table)
        row0 = -(row_origin % 32)
        tiles_x, tiles_y, words = -(-w // 64), -(-(h - row0) // 32), -(-n // 32)
        cells = tiles_x * tiles_y
        w4, h4 = int(table["w4"].max()), int(table["h4"].max())
        cap = max(-(-4 * w4 // 32) * -(-h4 // 8) * n,                      # reduce blocks
                  -(-2 * w4 // 256) * -(-2 * h4 // 4) * n * n_blurs,      # horizontal blur blocks (256 x 4 cells)
                  -(-2 * w4 // 64) * -(-2 * h4 // 16) * n * n_blurs,      # ... in 64 x 16 blocks
                  -(-2 * w4 // 32) * -(-2 * h4 // 64) * n * n_blurs)      # vertical blur blocks
        bits = torch.empty(4 + 2 * cap + 4 * cells * words, dtype=torch.int32, device=self.device)
        multi = torch.empty(cells, dtype=torch.uint8, device=self.device)
        maps = np.zeros(1, dtype=_lib.TILE_MAPS)
        base = bits.data_ptr()
        # work_count: [0] the live counter, [1..3] blocks run by reduce / horizontal / vertical blur (kept
        # for the bench's "bytes of the blocks actually run"); work: 8-byte items
        maps["work_count"], maps["work"], maps["work_cap"] = base, base + 16, cap
        base += 16 + 8 * cap
        maps["present"], maps["cand"], maps["need"] = base, base + 4 * cells * words, base + 8 * cells * words
        if seam_plan:
            maps["wneed"] = base + 12 * cells * words
        maps["multi"] = multi.data_ptr()
        maps["tiles_x"], maps["tiles_y"], maps["words"], maps["row0"] = tiles_x, tiles_y, words, row0
        maps["reach_x"], maps["reach_y"] = -(-pad // 64), -(-pad // 32)
        maps["h_rows"] = self.blur_h_rows
        return maps, (bits, multi)

    def _coarse_layout(self, table, n_blurs):
        """HBM for the coarse levels of every patch, back to back: on the f = 2 grid the reduce
        output, the horizontal-pass scratch and the blurred level 0; on the f = 4 grid the reduce
        output and (scratch, blurred level) for every level >= 1.  Fills the pointer fields of the
        patch table; returns the pools and the per-level (in, tmp, out) pointer arrays."""
        cells = table["w4"].astype(np.int64) * table["h4"]                # f = 4 cells per patch
        first = np.concatenate([[0], np.cumsum(cells)]).astype(np.uint64)
        total = int(first[-1])
        pool2 = self._coarse_pool("pool2", 3 * 4 * total * 4)
        pool4 = self._coarse_pool("pool4", (1 + 2 * (n_blurs - 1)) * total * 4)
        base2, base4 = np.uint64(pool2.data_ptr()), np.uint64(pool4.data_ptr())
        plane2, plane4 = np.uint64(16 * 4 * total), np.uint64(16 * total)      # bytes per plane
        at2, at4 = np.uint64(64) * first[:-1], np.uint64(16) * first[:-1]      # byte offset of each patch
        table["d2"], table["d4"] = base2 + at2, base4 + at4
        levels = [(table["d2"], base2 + plane2 + at2, base2 + np.uint64(2) * plane2 + at2)]
        for lvl in range(1, n_blurs):
            levels.append((table["d4"], base4 + np.uint64(2 * lvl - 1) * plane4 + at4,
                           base4 + np.uint64(2 * lvl) * plane4 + at4))
        for lvl, (_, _, out) in enumerate(levels):
            table["low"][:, lvl] = out
        return {"pool2": pool2, "pool4": pool4, "levels": levels, "cells": total, "first": first}

    def _coarse_pool(self, name, numel):
        """Grow-only float32 buffer for the coarse levels, filled once with finite values and from
        then on only ever written by the reduce / blur kernels (see ``pool_fill``)."""
        pool = self._pools.get(name)
        if pool is None or pool.numel() < numel:
            self._pools.pop(name, None)
            pool = torch.empty(int(numel * 1.1) + 1024, dtype=torch.float32, device=self.device)
            pool.fill_(self.pool_fill)
            self._pools[name] = pool
        elif self.repoison:
            pool.fill_(self.pool_fill)
        return pool[:numel]

    def _blur_jobs(self, table, layout, dev_table, pad):
        """One p360_blur_job per (level, patch): level 0 on the f = 2 grid, the others on f = 4."""
        n = len(table)
        patch_ptr = np.uint64(dev_table.data_ptr()) + np.uint64(_lib.BAND_PATCH.itemsize) * np.arange(n, dtype=np.uint64)
        jobs = np.zeros(n * len(layout["levels"]), dtype=_lib.BLUR_JOB)
        for lvl, (src, tmp, out) in enumerate(layout["levels"]):
            sl = jobs[lvl * n:(lvl + 1) * n]
            scale = 2 if lvl == 0 else 1
            sl["in"], sl["tmp"], sl["out"], sl["slot"] = src, tmp, out, lvl
            sl["w"], sl["h"], sl["shift"] = scale * table["w4"], scale * table["h4"], 1 if lvl == 0 else 2
            sl["patch"], sl["pad"], sl["grow"] = patch_ptr, pad, 2 * pad + 4
        return jobs

    def _level_views(self, table, layout):
        """Per patch, tensor views of its blurred coarse levels (stage-level tests)."""
        pool2, pool4, total = layout["pool2"], layout["pool4"], layout["cells"]
        lows = []
        for k in range(len(table)):
            w4, h4, o = int(table["w4"][k]), int(table["h4"][k]), int(layout["first"][k])
            per = [pool2[(2 * 4 * total + 4 * o) * 4:][:4 * h4 * w4 * 4].view(2 * h4, 2 * w4, 4)]
            for lvl in range(1, len(layout["levels"])):
                per.append(pool4[(2 * lvl * total + o) * 4:][:h4 * w4 * 4].view(h4, w4, 4))
            lows.append(per)
        return lows

    def _set_taps(self, n_levels, plan):
        if self._taps_key == n_levels:
            return
        for slot, (_, taps) in enumerate(plan):
            _lib.call("p360_blur_set_taps", slot, taps.ctypes.data_as(C.POINTER(C.c_float)), len(taps),
                      self.stream)
        self._taps_key = n_levels

    def _collapse(self, name, nbytes, fn, head, mosaic, out_host=None, rows=None, on_band=None, bands=8,
                  row_origin=0, tail=(), cols=None, col_origin=0, out=None):
        """Launch a collapse kernel over rows ``rows`` (default: all) of the
        mosaic buffer.  With ``out_host`` (pinned host array) or ``on_band``
        (callback(y0, y1), e.g. an NVLink send) the rows are produced band by
        band so that the transfer of each finished band overlaps the
        computation of the next.  ``cols = (xa, xb)`` (buffer columns, xa a multiple of 64): only
        those columns are wanted — the multiband collapse produces just them, and just they are
        downloaded (buffer column x is mosaic column x + ``col_origin``).  ``out = (address, pitch)``:
        the kernel stores its bytes there instead of in ``mosaic`` — buffer pixel (x, y) at
        address + 3 (y pitch + x): the window's place in a larger image (``composite(out_dev=...)``)."""
        h, w = mosaic.shape[:2]
        ya, yb = (0, h) if rows is None else rows
        xa, xb = (0, w) if cols is None else cols
        xargs = (xa, xb) if fn == "p360_multiband_collapse" else ()
        optr, opitch = (_lib.ptr(mosaic), 0) if out is None else out
        if out_host is None and on_band is None:
            if fn is not None:
                self._traced(name, nbytes, fn, *head, optr, opitch, ya, yb, *xargs, row_origin, w, *tail, self.stream)
            return
        host = None if out_host is None else torch.from_numpy(out_host)
        whole_rows = host is not None and xa == 0 and xb == w and col_origin == 0 and host.shape[1] == w
        main = torch.cuda.current_stream(self.device)
        for y0, y1 in band_edges(ya, yb, bands):
            if y1 <= y0:
                continue
            if fn is not None:             # (None: the rows are already final, e.g. a blank window)
                self._traced(name, nbytes * (y1 - y0) // max(yb - ya, 1), fn, *head, optr, opitch, y0, y1, *xargs,
                             row_origin, w, *tail, self.stream)
            if on_band is not None:
                on_band(y0, y1)
            if host is not None:
                side = self.download_stream()
                done = torch.cuda.Event()
                done.record(main)
                side.wait_event(done)
                with torch.cuda.stream(side):       # buffer row y is mosaic row y + row_origin
                    if whole_rows:
                        host[y0 + row_origin:y1 + row_origin].copy_(mosaic[y0:y1], non_blocking=True)
                    else:                           # a column window: a rectangle of the host mosaic
                        pitch = 3 * host.shape[1]
                        _lib.call("p360_copy_rect", host.data_ptr() + (y0 + row_origin) * pitch + 3 * (xa + col_origin),
                                  pitch, mosaic.data_ptr() + 3 * (y0 * w + xa), 3 * w, 3 * (xb - xa), y1 - y0,
                                  side.cuda_stream)
                landed = torch.cuda.Event()
                landed.record(side)
                band = (y0 + row_origin, y1 + row_origin, landed, xa + col_origin, xb + col_origin)
                if self._drain is not None:
                    self._drain[0].put(band)           # (a copy-out thread is running: composite_streamed)
                else:
                    self._bands_down.append(band)
                self._mark(f"rows {y0 + row_origin}-{y1 + row_origin} collapsed")
                self._mark(f"rows {y0 + row_origin}-{y1 + row_origin} downloaded", side)
        if host is not None and self._down is not None:
            tail = self._down[0]
            for other in self._down[1:]:
                tail.wait_stream(other)
            self._download = torch.cuda.Event()
            self._download.record(tail)

    def _blank(self, mosaic, out_host, rows, on_band, bands, row_origin, cols=None, col_origin=0, out=None):
        """A mosaic (or window) no image touches: zeros — through the same banded path as a
        collapse, so that ``out_host`` receives its rows and ``on_band`` fires for every band
        (a strip that falls into a gap between images must still send its bands)."""
        mosaic.zero_()
        self._collapse("blank", 0, None, (), mosaic, out_host, rows, on_band, bands, row_origin, cols=cols,
                       col_origin=col_origin)
        return mosaic

    def release(self, everything=False):
        """Drop the references that keep the last composite's transient buffers alive; the memory
        goes back to torch's caching allocator.  Call only after the work that uses them has been
        waited for.  What is kept for the next composite of the same rig (prepared job tables and
        patch pools, the coarse pools, image buffers of ``upload(reuse=True)``) goes too with
        ``everything``."""
        self.last_covered = None
        if everything:
            self._ingest_stage = None
            self._prepared.clear()
            self._pools.clear()
            self._images.clear()
            self._packed.clear()
        for key in ("warp", "warp_jobs", "bands", "collapse", "streamed", "seam", "exact"):
            self._keep.pop(key, None)

    def _start_copy_out(self, copy_to, staged):
        """A thread that moves every band of the banded download from the pinned staging buffer to
        the caller's (pageable) array as soon as it has landed — while the main thread goes on
        staging images and queueing windows (the copies release the GIL)."""
        import queue
        import threading
        bands, errors = queue.Queue(), []

        def run():
            try:
                while True:
                    band = bands.get()
                    if band is None:
                        return
                    y0, y1, landed, x0, x1 = band
                    landed.synchronize()
                    parallel_copy(copy_to[y0:y1, x0:x1], staged[y0:y1, x0:x1])
            except BaseException as exc:          # handed to the caller by _join_copy_out
                errors.append(exc)
        thread = threading.Thread(target=run, name="p360-copy-out", daemon=True)
        self._drain = (bands, thread, errors)
        thread.start()

    def _join_copy_out(self):
        bands, thread, errors = self._drain
        self._drain = None
        bands.put(None)
        thread.join()
        if errors:
            raise errors[0]

    def finish_download(self, copy_to=None, staged=None):
        """Block until a banded download started by ``_collapse`` has landed.  ``copy_to`` (a
        pageable host array) receives the rows from the pinned staging buffer ``staged`` band by
        band as they land, while the later bands are still crossing PCIe."""
        if copy_to is not None:
            for y0, y1, landed, x0, x1 in self._bands_down:
                landed.synchronize()
                parallel_copy(copy_to[y0:y1, x0:x1], staged[y0:y1, x0:x1])
        self._bands_down = []
        if self._download is not None:
            self._download.synchronize()
            self._download = None

    def blend_multiband(self, patches, shape, n_levels=5, stages=None, owner_state=None, out_host=None,
                        rows=None, on_band=None, mosaic=None, bands=8, row_origin=0, use_maps=None, seam=None,
                        cols=None, col_origin=0, out=None):
        """stitcher.py:186-241.  The wide blurs are evaluated on coarse grids
        and every mosaic pixel gathers its bands from the patches covering it,
        in list order, so no mosaic-sized accumulator ever touches HBM.

        ``seam`` (from ``composite``): the patches have NOT been warped yet — the seam plan
        decides from the geometry which tiles are one patch's pixels (written straight from the
        sources as uint8) and warps float patches only in the seam zone."""
        h, w = shape
        if not 1 <= n_levels <= _lib.MAX_LEVELS:
            raise ValueError(f"n_levels must be in 1..{_lib.MAX_LEVELS}")
        if mosaic is None:
            mosaic = torch.empty((h, w, 3), dtype=torch.uint8, device=self.device)
        if not patches:      # nothing lands here: still produce (and download / hand on) every band
            return self._blank(mosaic, out_host, rows, on_band, bands, row_origin, cols, col_origin)
        pad, plan = geo.coarse_band_plan(n_levels)
        n = len(patches)
        lows, maps = [], None
        if plan:
            self._set_taps(n_levels, plan)
        if seam is not None:
            # K0: plan the tiles; K1t: one block per tile — solo tiles straight to uint8, float
            # patches + owner keys in the seam zone only.  Everything that depends on the geometry
            # and on the buffer addresses alone is prepared once (``seam["prepared"]``).
            assert plan, "the seam plan needs at least two bands"
            prep = seam["prepared"]
            jobs, crops, src = prep["jobs"], seam["crops"], seam["src"]
            if "table" not in prep:
                table = self._band_table(patches, pad, coarse=True)
                layout = self._coarse_layout(table, len(plan))
                pristine = self._to_device(table.view(np.uint8).reshape(-1))
                dev_table = torch.empty_like(pristine)
                maps, maps_keep = self._tile_maps(table, len(plan), h, w, pad, row_origin, seam_plan=True)
                blur_jobs = self._blur_jobs(table, layout, dev_table, pad)
                prep.update(table=table, layout=layout, pristine=pristine, dev_table=dev_table, maps=maps,
                            maps_keep=maps_keep, dev_wjobs=self._to_device(jobs.view(np.uint8).reshape(-1)),
                            blur_jobs=blur_jobs, dev_blur_jobs=self._to_device(blur_jobs.view(np.uint8).reshape(-1)),
                            keys=torch.empty((h, w), dtype=torch.int64, device=self.device),     # written where they are read
                            covered=torch.empty((h, w), dtype=torch.uint8, device=self.device),
                            pix=int((table["pw"].astype(np.int64) * table["ph"]).sum()),
                            pools=(layout["pool2"].data_ptr(), layout["pool4"].data_ptr()))
            table, layout, dev_table, maps, maps_keep = (prep[k] for k in ("table", "layout", "dev_table", "maps", "maps_keep"))
            dev_wjobs, keys, covered, pix = prep["dev_wjobs"], prep["keys"], prep["covered"], prep["pix"]
            dev_table.copy_(prep["pristine"])                     # (K0 grows the `own` boxes in it)
            self._traced("K0_seam_plan", 216 * n, "p360_seam_plan_build", _lib.ptr(dev_wjobs), n, _lib.ptr(dev_table),
                         h, w, row_origin, seam["mosaic_h"], maps.ctypes.data, self.stream)
            if src.ready is not None:          # a tile reads whichever images meet it: their uploads first
                main = torch.cuda.current_stream(self.device)
                for i in sorted({c[0] for c in crops} if seam.get("reads") is None else seam["reads"]):
                    main.wait_event(src.ready[i])
            ya, yb = (0, h) if rows is None else rows
            xa, xb = (0, w) if cols is None else cols
            optr, opitch = (_lib.ptr(mosaic), 0) if out is None else out
            self._traced("K1t_warp_tiles", 30 * seam["pixels"], "p360_warp_tiles", jobs.ctypes.data, _lib.ptr(dev_wjobs), n,
                         _lib.ptr(keys), _lib.ptr(covered), optr, opitch, ya, yb, xa, xb, h, w,
                         int(bool(seam.get("want_covered"))), maps.ctypes.data, self.stream)
            self._keep["seam"] = (prep,)
            if seam.get("after_warp") is not None:
                # the caller wants to move the tiles that are final after the tile warp (all but the
                # multi tiles) while reduce / blur / collapse still run: hand it the plan's multi map
                # (a function of the geometry alone: fetched once per prepared window)
                if "multi_host" not in prep:
                    torch.cuda.current_stream(self.device).synchronize()
                    prep["multi_host"] = maps_keep[1].cpu().numpy().reshape(int(maps["tiles_y"][0]), int(maps["tiles_x"][0])) != 0
                seam["after_warp"](mosaic, prep["multi_host"], int(maps["row0"][0]))
        else:
            keys, covered = owner_state if owner_state is not None else self.owner_state_for(patches, shape)
            table = self._band_table(patches, pad, coarse=True)
            layout = self._coarse_layout(table, len(plan)) if plan else None
            dev_table = self._table(table, "band_table")
            pix = int((table["pw"].astype(np.int64) * table["ph"]).sum())
        if plan:
            # where can a patch carry weight at all: seam-band bitmaps, or the box around its owned pixels
            if seam is None:
                if use_maps is None:
                    use_maps = self.seam_maps if self.seam_maps is not None else h * w >= SEAM_MAPS_MIN_PIXELS
                if use_maps:
                    maps, maps_keep = self._tile_maps(table, len(plan), h, w, pad, row_origin)
                    self._traced("K2b_tile_maps", 9 * h * w, "p360_tile_maps_build", _lib.ptr(keys), _lib.ptr(covered),
                                 _lib.ptr(dev_table), n, h, w, maps.ctypes.data, self.stream)
                else:
                    maps_keep = None
                    self._traced("K2b_owned_boxes", 8 * h * w, "p360_owned_boxes", _lib.ptr(keys), _lib.ptr(dev_table),
                                 n, h, w, self.stream)
            maps_ptr = None if maps is None else maps.ctypes.data
            if seam is not None:
                jobs, dev_jobs = seam["prepared"]["blur_jobs"], seam["prepared"]["dev_blur_jobs"]
            else:
                jobs = self._blur_jobs(table, layout, dev_table, pad)
                dev_jobs = self._table(jobs, "blur_jobs")
            if seam is not None and "w4h4" in seam["prepared"]:        # (prepared composite: computed once)
                max_w4, max_h4 = seam["prepared"]["w4h4"]
            else:
                max_w4, max_h4 = int(table["w4"].max()), int(table["h4"].max())
                if seam is not None:
                    seam["prepared"]["w4h4"] = (max_w4, max_h4)
            self._traced("K3a_pyramid_reduce", 25 * pix, "p360_pyramid_reduce_batch", _lib.ptr(dev_table), n,
                         max_w4, max_h4, _lib.ptr(keys), w, maps_ptr, self.stream)
            coarse_px = int((4 + len(plan) - 1) * layout["cells"])
            self._traced("K3_gauss_blur", 32 * coarse_px, "p360_gauss_blur_batch", _lib.ptr(dev_jobs),
                         len(jobs), 2 * max_w4, 2 * max_h4, maps_ptr, self.stream)
            if maps is not None:
                _lib.launch_count += 3          # the scan kernels that compact the block lists
            self._keep["bands"] = (layout["pool2"], layout["pool4"], dev_jobs, maps, maps_keep)
            if stages is not None:
                lows = self._level_views(table, layout)
        self._collapse("K4_multiband_collapse", 16 * pix + 12 * h * w, "p360_multiband_collapse",
                       (_lib.ptr(dev_table), n, n_levels, _lib.ptr(keys), _lib.ptr(covered)), mosaic, out_host,
                       rows, on_band, bands, row_origin, tail=(None if maps is None else maps.ctypes.data,),
                       cols=cols, col_origin=col_origin, out=out)
        self._keep["collapse"] = (dev_table, keys, covered)
        self.last_covered = covered
        if stages is not None:
            stages.update(keys=keys, covered=covered, lows=lows, maps=maps)
        return mosaic

    def blend_multiband_exact(self, patches, shape, n_levels=5, owner_state=None, mosaic=None):
        """stitcher.py:186-241 stage by stage at full resolution: every patch blurred with the
        reference's own 33-97-tap Gaussians at every level, bands accumulated per level in
        mosaic-sized float images.  Two orders of magnitude slower than ``blend_multiband`` and
        agreeing with the reference to float rounding: the device-side ground truth of the
        coarse-grid pipeline, and the path for images too small for the coarse grids.  The alpha
        channel of the patches is overwritten with the owner mask, like the reference's."""
        h, w = shape
        if not 1 <= n_levels <= _lib.MAX_LEVELS:
            raise ValueError(f"n_levels must be in 1..{_lib.MAX_LEVELS}")
        if mosaic is None:
            mosaic = torch.empty((h, w, 3), dtype=torch.uint8, device=self.device)
        if not patches:
            return mosaic.zero_()
        keys, covered = owner_state if owner_state is not None else self.owner_state_for(patches, shape)
        acc = torch.zeros((n_levels, h, w, 4), dtype=torch.float32, device=self.device)
        level_bytes = 16 * h * w
        biggest = max((p.box[2] - p.box[0]) * (p.box[3] - p.box[1]) for p in patches)
        work = torch.empty((3, biggest * 4), dtype=torch.float32, device=self.device)     # two blur outputs + scratch
        taps = [geo.gaussian_taps(geo.band_sigma(lvl)) for lvl in range(n_levels - 1)]
        for k, p in enumerate(patches):
            pw, ph, x0, y0 = self._args(p)
            if pw == 0 or ph == 0:
                continue
            _lib.call("p360_owner_to_alpha", p.rgba_ptr, pw, ph, x0, y0, k, _lib.ptr(keys), w, self.stream)
            prev = p.rgba_ptr
            for lvl in range(n_levels - 1):
                cur = work[lvl & 1].data_ptr()
                t = taps[lvl]
                _lib.call("p360_gauss_blur", p.rgba_ptr, cur, work[2].data_ptr(), pw, ph,
                          t.ctypes.data_as(C.POINTER(C.c_float)), len(t), self.stream)
                _lib.call("p360_band_accumulate", prev, cur, pw, ph, x0, y0, acc.data_ptr() + lvl * level_bytes, w, self.stream)
                prev = cur
            _lib.call("p360_band_accumulate", prev, None, pw, ph, x0, y0,
                      acc.data_ptr() + (n_levels - 1) * level_bytes, w, self.stream)
        _lib.call("p360_exact_collapse", _lib.ptr(acc), n_levels, _lib.ptr(covered), _lib.ptr(mosaic), h, w, self.stream)
        self._keep["exact"] = (acc, work, keys, covered)
        self.last_covered = covered
        return mosaic

    def _pointwise(self, fn, name, patches, shape, out_host=None, rows=None, on_band=None, mosaic=None,
                   bands=8, row_origin=0, cols=None, col_origin=0, out=None):
        h, w = shape
        if mosaic is None:
            mosaic = torch.empty((h, w, 3), dtype=torch.uint8, device=self.device)
        if not patches:
            return self._blank(mosaic, out_host, rows, on_band, bands, row_origin, cols, col_origin)
        table = self._band_table(patches)
        dev_table = self._table(table, "band_table")
        pix = int((table["pw"].astype(np.int64) * table["ph"]).sum())
        self._collapse(name, 17 * pix + 3 * h * w, fn, (_lib.ptr(dev_table), len(patches)), mosaic, out_host,
                       rows, on_band, bands, row_origin, cols=cols, col_origin=col_origin, out=out)
        self._keep["collapse"] = (dev_table,)
        return mosaic

    def blend_none(self, patches, shape, out_host=None, rows=None, on_band=None, mosaic=None, bands=8,
                   row_origin=0, cols=None, col_origin=0, out=None):
        """stitcher.py:160-168 (last valid writer wins), gather form."""
        return self._pointwise("p360_paste_collapse", "K7_paste_collapse", patches, shape, out_host, rows,
                               on_band, mosaic, bands, row_origin, cols, col_origin, out)

    def blend_linear(self, patches, shape, out_host=None, rows=None, on_band=None, mosaic=None, bands=8,
                     row_origin=0, cols=None, col_origin=0, out=None):
        """stitcher.py:171-183, gather form."""
        return self._pointwise("p360_linear_collapse", "K6_linear_collapse", patches, shape, out_host, rows,
                               on_band, mosaic, bands, row_origin, cols, col_origin, out)

    def covered_mask(self, patches, shape):
        """Area of validity for the crop stage (stitcher.py:266-271)."""
        h, w = shape
        covered = torch.zeros((h, w), dtype=torch.uint8, device=self.device)
        for p in patches:
            pw, ph, x0, y0 = self._args(p)
            _lib.call("p360_cover_update", p.invalid_ptr, pw, ph, x0, y0, _lib.ptr(covered), w,
                      self.stream)
        return covered

    def crop_rect(self, covered):
        """Largest all-valid rectangle of a device ``covered`` mask (stitcher.py:340-369):
        (y0, y1, x0, x1), the crop being mosaic[y0:y1, x0:x1]."""
        h, w = covered.shape
        covered = covered.contiguous()
        scratch = torch.empty(int(_lib.call("p360_crop_scratch_bytes", h, w)), dtype=torch.uint8, device=self.device)
        rect = torch.empty(4, dtype=torch.int32, device=self.device)
        self._traced("K9_crop_rect", 5 * h * w, "p360_crop_rect", _lib.ptr(covered), h, w, _lib.ptr(scratch), _lib.ptr(rect),
                     self.stream)
        return tuple(int(v) for v in rect.cpu().tolist())

    def blend(self, kind, patches, shape, n_levels=5, out_host=None, rows=None, on_band=None):
        if kind == "none":
            return self.blend_none(patches, shape, out_host, rows, on_band)
        if kind == "linear":
            return self.blend_linear(patches, shape, out_host, rows, on_band)
        if kind == "multiband":
            return self.blend_multiband(patches, shape, n_levels, out_host=out_host, rows=rows, on_band=on_band)
        raise ValueError(f"unknown blender {kind!r}")

    # -- whole path, device resident ------------------------------------------
    def blur_reach(self, kind, n_levels):
        """Reach of the widest coarse blur (reduce + blur + expand) in pixels, 0 for
        pointwise blenders."""
        if kind != "multiband" or n_levels < 2:
            return 0
        return geo.coarse_band_plan(n_levels)[0] + 4

    def window_halo(self, kind, n_levels):
        """Rows of context a row window needs on each side: the blur reach, plus one collapse
        tile.  The collapse decides per 64 x 32 tile (anchored at absolute rows) which patches
        carry weight and whether the tile is a single owner's pixels; a tile straddling the
        window edge must see true data on all its rows, or the decision — and with it the last
        bit of a pixel — would depend on where the window was cut."""
        reach = self.blur_reach(kind, n_levels)
        return reach + 32 if reach else 0

    def window_margin(self, kind, n_levels):
        """Upper bound of the rows beyond [ya, yb) that ``window_rows`` (plus the alignment of
        cropped patch tops) may ask for: which images a window depends on."""
        halo = self.window_halo(kind, n_levels)
        if not halo:
            return 0
        return max(halo, 32 * -(-geo.coarse_band_plan(n_levels)[0] // 32) + 31) + 3

    def window_rows(self, rows, kind, n_levels, height):
        """Rows [wa, wb) a row window [ya, yb) has to warp: the halo on both sides, widened to
        whole 32-row tiles so that every tile the seam plan consults for a tile of the window
        (the blur reach, in tiles) lies completely inside — the plan then cannot depend on where
        the window was cut."""
        ya, yb = rows
        halo = self.window_halo(kind, n_levels)
        wa, wb = ya - halo, yb + halo
        if halo:
            reach_y = -(-geo.coarse_band_plan(n_levels)[0] // 32)
            wa = min(wa, 32 * (ya // 32 - reach_y))
            wb = max(wb, 32 * ((yb - 1) // 32 + reach_y + 1))
        return max(0, wa), min(height, wb)

    def needs_exact(self, regions):
        """Images smaller than a few blur radii: the owner masks can be slivers a pixel or two wide,
        which the f = 2 / f = 4 grids cannot resolve (deviations of up to 4 grey levels were seen on
        20-30 px wide views) — such rigs are blended at full resolution (``blend_multiband_exact``)."""
        return min(min(r.img.shape[:2]) for r in regions) < EXACT_BELOW

    def window_cols(self, cols, kind, n_levels, width):
        """Columns [ca, cb) a column window [xa, xb) has to warp: the blur reach plus one 64-column
        collapse tile on both sides, widened to the whole tiles the seam plan consults (the
        column counterpart of ``window_rows``)."""
        xa, xb = cols
        reach = self.blur_reach(kind, n_levels)
        if not reach:
            return xa, xb
        halo = reach + 64
        reach_x = -(-geo.coarse_band_plan(n_levels)[0] // 64)
        ca = min(xa - halo, 64 * (xa // 64 - reach_x))
        cb = max(xb + halo, 64 * ((xb - 1) // 64 + reach_x + 1))
        return max(0, ca), min(width, cb)

    def col_margin(self, kind, n_levels):
        """Upper bound of the columns beyond [xa, xb) a column window may ask for (halo, tile
        rounding of the buffer, alignment of cropped left edges): which images it depends on."""
        reach = self.blur_reach(kind, n_levels)
        if not reach:
            return 0
        return max(reach + 64, 64 * -(-geo.coarse_band_plan(n_levels)[0] // 64) + 63) + 64 + 3

    def _window_geometry(self, regions, plan, kind, n_levels, proj, rows, cols):
        """What a window [ya, yb) x [xa, xb) of the mosaic is computed from: the crops of the images
        (window + halo), the buffer that holds them — rows [top, top + shape[0]), columns [left,
        left + shape[1]) of the mosaic, whole 64-column tiles of it — and the window in buffer
        coordinates (``local`` rows, ``local_cols`` or None for all columns)."""
        reach = self.blur_reach(kind, n_levels)
        height, width = plan.shape
        if cols is not None:
            xa, xb = cols
            if not (0 <= xa < xb <= width and xa % 64 == 0 and (xb % 64 == 0 or xb == width)):
                raise ValueError(f"column window {cols} must lie on 64-column tile edges of the {width}-column mosaic")
        if rows is None and cols is None:
            ya, yb, wa, wb = 0, height, 0, height
            crops, tables = self.plan_crops(regions, plan, proj, split_dilate=2 * reach)
        else:
            ya, yb = (0, height) if rows is None else rows
            wa, wb = (0, height) if rows is None else self.window_rows(rows, kind, n_levels, height)
            span = None if cols is None else self.window_cols(cols, kind, n_levels, width)
            crops, tables = self.plan_crops(regions, plan, proj, rows=None if rows is None else (wa, wb),
                                            row_align=4 if reach else 1, split_dilate=2 * reach,
                                            cols=span, col_align=4 if reach else 1)
        top = min([c[2] for c in crops] + [wa])                # aligned crops may start above wa
        if cols is None:
            left, right, local_cols = 0, width, None
        else:
            # the buffer spans whole tiles of the mosaic: its tiles are tiles of the whole mosaic
            left = min([c[1] for c in crops] + [span[0]]) // 64 * 64
            right = min(width, -(-span[1] // 64) * 64)
            local_cols = (xa - left, xb - left)
        return crops, tables, top, left, (wb - top, right - left), (ya - top, yb - top), local_cols

    def _images_read(self, regions, plan, kind, n_levels, proj, crops, top, left, shape):
        """The images a window's tile warp can sample, if the whole mosaic's seam plan is known
        (``source_rects`` ran for this plan): those the plan reads somewhere inside the window's
        buffer.  A patch the whole plan reads nowhere in the buffer is `present` in none of its
        tiles, so it takes no part in the window's plan either — its job stays in the table, its
        pixels are never loaded, and the window need not wait for its upload.  None: unknown,
        every image of the crops counts."""
        used = plan._crops.get(("used", len(regions), proj, None, None, kind, n_levels)) if plan._crops is not None else None
        if used is None:
            return None
        x0, y0, x1, y1 = left, top, left + shape[1], top + shape[0]
        return sorted({c[0] for c in crops
                       if any(b[0] < x1 and b[2] > x0 and b[1] < y1 and b[3] > y0 for b in used.get(c[0], ()))})

    def composite(self, regions, src, plan, kind, n_levels=5, proj=geo.SphProj, rows=None, out_host=None,
                  on_band=None, bands=8, direct=None, want_covered=False, exact=False, cols=None, out_dev=None,
                  after_warp=None):
        """warp + blend for the whole mosaic or for a window of it: rows [ya, yb) and / or columns
        [xa, xb) (xa a multiple of 64; xb too unless it is the mosaic width).  The returned strip
        has exactly (yb - ya) x (xb - xa) pixels and is bit-identical to that part of the full
        composite; only that part is collapsed.
        ``out_host`` (pinned uint8 H x W x 3: the WHOLE mosaic, also in window
        mode) receives the pixels produced through a banded download that overlaps
        the collapse; call ``finish_download`` before reading it.  ``on_band(strip_part, y0, y1)``
        is called after the collapse of mosaic rows [y0, y1) has been launched
        (``strip_part`` = that part of the device result, columns [xa, xb) only).  ``want_covered``:
        keep the union of valid pixels of the rows produced in ``last_covered`` (crop stage,
        stitcher.py:266-271).  ``out_dev = (address, width)`` of a device image of the WHOLE mosaic —
        this GPU's, or a peer's mapped over NVLink: the kernels store the window's bytes straight
        into their place there (no local result: the returned strip is None).
        ``after_warp(buffer, multi, geometry)``: called (seam-plan path only) once the tile warp has
        been launched — every tile that is not ``multi`` (bool [tiles_y, tiles_x] over the 64 x 32
        tiles of the window's buffer) is final from then on; ``geometry`` = dict(top, left, row0,
        rows, cols): where the buffer sits in the mosaic, where its tile rows start, and the
        window in buffer coordinates.  Returns with ``used_after_warp`` telling whether it fired."""
        self._mark(f"composite {rows} {cols} begins")
        self.used_after_warp = False
        if cols is not None and want_covered:
            raise ValueError("want_covered needs every column of the mosaic")
        crops, tables, top, left, shape, local, local_cols = self._window_geometry(regions, plan, kind, n_levels, proj,
                                                                                   rows, cols)
        holder = {}
        band_cb = None
        if on_band is not None:
            def band_cb(y0, y1):
                part = holder["mosaic"][y0:y1] if local_cols is None else holder["mosaic"][y0:y1, local_cols[0]:local_cols[1]]
                on_band(part, y0 + top, y1 + top)

        def result(strip):
            strip = strip[local[0]:local[1]]
            return strip if local_cols is None else strip[:, local_cols[0]:local_cols[1]]
        window = dict(rows=local, on_band=band_cb, bands=bands, row_origin=top, cols=local_cols, col_origin=left)
        if out_dev is not None:
            if exact or on_band is not None or out_host is not None:
                raise ValueError("out_dev excludes exact / on_band / out_host")
            window["out"] = (out_dev[0] + 3 * (top * out_dev[1] + left), out_dev[1])
            if not crops:             # nothing lands here: zeros in place
                view = out_dev[2][top + local[0]:top + local[1]]
                (view if local_cols is None else view[:, left + local_cols[0]:left + local_cols[1]]).zero_()
                return None, []
            result = lambda strip: None
        if exact and kind == "multiband" and n_levels > 1:      # (stitch() asks for it when needs_exact())
            # full-resolution loop nest on the whole window, then hand the rows on like a collapse
            state = self.new_owner_state(shape)
            patches = self.warp_crops(src, crops, tables, origin=(left, top), owner_state=state)
            holder["mosaic"] = self.blend_multiband_exact(patches, shape, n_levels, owner_state=state)
            self._collapse("exact", 0, None, (), holder["mosaic"], out_host, local, band_cb, bands, top,
                           cols=local_cols, col_origin=left)
            return result(holder["mosaic"]), patches
        use_plan = (self.direct if direct is None else direct) and kind == "multiband" and n_levels > 1 \
            and 0 < len(crops) <= 256                          # (the tile warp keeps its job table in constant memory)
        if use_plan:
            # job tables, pools, tile-map storage: prepared once per (geometry, window, source
            # buffers) and reused as long as the addresses they name are the ones in use
            needed = sorted({c[0] for c in crops})
            key = (id(plan), rows, cols, n_levels, proj, shape, tuple(src.pixels[i].data_ptr() for i in needed),
                   tuple(src.luts[i].data_ptr() for i in needed))
            prep = self._prepared.get(key)
            if prep is not None and "pools" in prep and prep["pools"] != tuple(
                    self._pools[k].data_ptr() if k in self._pools else 0 for k in ("pool2", "pool4")):
                prep = None                                    # (the coarse pools were re-allocated since)
            if prep is None:
                jobs, patches, keep = self._warp_jobs(src, crops, tables, origin=(left, top))
                prep = {"jobs": jobs, "patches": patches, "keep": keep, "plan": plan}
                self._prepared.pop(key, None)
                while len(self._prepared) >= self.prepared_max:
                    self._prepared.pop(next(iter(self._prepared)))
                self._prepared[key] = prep
            patches = prep["patches"]
            self._keep["warp"] = prep["keep"][:3] + (prep["jobs"],)
            seam = {"prepared": prep, "crops": crops, "src": src, "mosaic_h": plan.shape[0], "pixels": prep["keep"][3],
                    "want_covered": want_covered}
            used_known = plan._crops is not None and ("used", len(regions), proj, None, None, kind, n_levels) in plan._crops
            if prep.get("reads_known") != used_known:               # (per prepared window; redone once the plan's rectangles exist)
                prep["reads"] = self._images_read(regions, plan, kind, n_levels, proj, crops, top, left, shape)
                prep["reads_known"] = used_known
            seam["reads"] = prep["reads"]
            if after_warp is not None:
                def fire(buffer, multi, row0):
                    self.used_after_warp = True
                    after_warp(buffer, multi, dict(top=top, left=left, row0=row0, rows=local,
                                                   cols=local_cols if local_cols is not None else (0, shape[1])))
                seam["after_warp"] = fire
            strip = self._blend_into(holder, self.blend_multiband, patches, shape, n_levels, out_host=out_host,
                                     seam=seam, **window)
            return result(strip), patches
        state = self.new_owner_state(shape) if kind == "multiband" else None
        patches = self.warp_crops(src, crops, tables, origin=(left, top), owner_state=state)
        if kind == "multiband":
            strip = self._blend_into(holder, self.blend_multiband, patches, shape, n_levels,
                                     owner_state=state, out_host=out_host, **window)
        else:
            strip = self._blend_into(holder, self.blend_none if kind == "none" else self.blend_linear,
                                     patches, shape, out_host=out_host, **window)
            if want_covered:
                self.last_covered = self.covered_mask(patches, shape)
        return result(strip), patches

    def streamed_windows(self, plan, kind, n_levels, windows=8, used=None):
        """Plan of ``composite_streamed``: the order in which to upload the images (left edge
        first; images that straddle the +-pi seam last) and column
        windows [xa, xb) of the mosaic, on 64-column tile edges, with the number of uploads each
        one has to wait for — a window needs nothing beyond the first ``count`` images — sorted by
        that number: the order in which they can be composited."""
        n = len(plan.boxes)
        width = plan.shape[1]
        margin = self.col_margin(kind, n_levels)
        reach = self.blur_reach(kind, n_levels)
        runs = {i: geo.active_column_runs(i, box, plan, dilate=2 * reach)
                for i, box in enumerate(plan.boxes) if box[2] > box[0] and box[3] > box[1]}
        if used is not None:        # (``used_boxes``) where the seam plan reads each image: tighter than its box
            runs = {i: sorted((b[0], b[2]) for b in boxes) for i, boxes in used.items() if boxes}
        # left edge first — but an image with two column runs (it straddles the +-pi seam: both ends of
        # the mosaic wait for it) goes last: the first window then needs one column of images, not two
        late = os.environ.get("P360_STRADDLERS_LAST", "1") == "1"
        order = sorted(range(n), key=lambda i: (late and len(runs.get(i, ())) > 1, runs[i][0][0] if i in runs else 0,
                                                plan.boxes[i][1], i))
        rank = {i: r for r, i in enumerate(order)}
        last = np.zeros(width, dtype=np.int64)            # per mosaic column: rank of the last upload it needs
        for i, parts in runs.items():
            for a, b in parts:
                a, b = max(0, a - margin), min(width, b + margin)
                last[a:b] = np.maximum(last[a:b], rank[i])
        # per 64-column tile column: how many uploads (in that order) it waits for
        tiles = -(-width // 64)
        need = np.array([int(last[64 * t:64 * (t + 1)].max()) + 1 for t in range(tiles)])
        # consecutive tile columns at the same stage (k-th `windows`-th of the uploads) form a window;
        # windows run in the order their images arrive — not necessarily left to right: the right
        # end of a 360-degree mosaic belongs to the images that straddle the seam, which come first
        stage = -(-need * windows // max(n, 1))
        wins, a = [], 0
        for t in range(1, tiles + 1):
            if t == tiles or stage[t] != stage[a]:
                wins.append((64 * a, min(64 * t, width), int(need[a:t].max())))
                a = t
        # a window pays a halo on both sides: slivers join the neighbour that delays them least
        while len(wins) > 1:
            k = min(range(len(wins)), key=lambda i: wins[i][1] - wins[i][0])
            if wins[k][1] - wins[k][0] >= 512:
                break
            near = [j for j in (k - 1, k + 1) if 0 <= j < len(wins)]
            j = min(near, key=lambda i: max(wins[i][2], wins[k][2]))
            lo, hi = min(j, k), max(j, k)
            wins[lo:hi + 1] = [(wins[lo][0], wins[hi][1], max(wins[lo][2], wins[hi][2]))]
        wins.sort(key=lambda w: (w[2], w[0]))
        return order, wins

    def composite_streamed(self, regions, plan, kind, n_levels, proj, out_host, windows=8, bands=2, exact=False,
                           copy_to=None):
        """Upload + composite + download with both PCIe directions busy: the images are uploaded
        left edge first — of each only the rectangle the seam plan can sample (``source_rects``) —
        and as soon as the images a column window of the mosaic depends on have arrived that
        window is composited (exactly the bytes of the whole composite, see ``composite``) and
        downloaded, while the uploads for the windows to its right continue on their own stream.
        ``out_host``: pinned uint8 H x W x 3; ``copy_to``: the pageable array the pixels finally go
        to (a thread copies every band on as soon as it has landed; the call then returns when the
        last one is there).  Call ``finish_download`` afterwards."""
        rects = None if exact else self.source_rects(regions, plan, kind, n_levels, proj)
        order, wins = self.streamed_windows(plan, kind, n_levels, windows,
                                            used=None if rects is None else self.used_boxes(regions, plan, kind, n_levels, proj))
        self.prepared_max = max(self.prepared_max, len(wins) + 2)
        # the copies are issued window by window: staging a pageable image (a threaded memcpy into a
        # pinned slot) then overlaps the kernels and transfers of the windows before it
        src = self.upload(regions, overlap=True, order=order, reuse=True, rects_of=rects,
                          need=None if rects is None else set(rects), lazy=True)
        strips = []
        if copy_to is not None:
            self._start_copy_out(copy_to, out_host)
        try:
            for xa, xb, count in wins:
                src.issue(count)
                strip, _ = self.composite(regions, src, plan, kind, n_levels, proj, cols=(xa, xb), out_host=out_host,
                                          bands=bands, exact=exact)
                strips.append(strip)          # the download stream still reads it: keep it allocated
        finally:
            if copy_to is not None:
                self._join_copy_out()
        src.issue()
        self._keep["streamed"] = (strips, src)
        return src

    def _blend_into(self, holder, blender, patches, shape, *args, **kwargs):
        """Run a blender whose band callback needs to see the output buffer:
        the buffer is allocated here and handed to the blender."""
        holder["mosaic"] = torch.empty(tuple(shape) + (3,), dtype=torch.uint8, device=self.device)
        return blender(patches, shape, *args, mosaic=holder["mosaic"], **kwargs)
