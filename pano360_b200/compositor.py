"""Device-side driver of the compositing path: owns HBM buffers (torch is used
for allocation, streams and copies only) and sequences the sm_100a kernels of
``libpano360_b200.so`` through the C ABI.

Data layout in HBM
------------------
* source image      u8  [h][w][3]            as delivered by cv2.imread (BGR)
* sample LUT        f32 [256] per image      u8 -> float value (gain folded in)
* hat tables        f64 [h], [w]             shared by images of equal size
* inverse-map tabs  f64 [pw][3], [ph][3]     per patch (column part, row part)
* patch             f32 [ph][pw][4] RGBA + u8 [ph][pw] invalid mask
* coarse levels     f32 [h/f][w/f][4]        blurred f=2 (level 0) / f=4 (levels >= 1) images per patch
* owner / best      i32 [H][W], f32 [H][W]   running arg-max of alpha
* covered           u8  [H][W]               union of valid pixels
* mosaic            u8  [H][W][3]

A *row window* ``(ya, yb)`` restricts all work to mosaic rows [ya, yb) plus a
halo of the largest blur radius (strip sharding); patches are
cropped in rows but keep their true columns, so reflections happen at true
patch edges wherever they influence rows inside the window.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _lib, geometry as geo


def _require_cuda(device):
    if not torch.cuda.is_available():
        raise RuntimeError("pano360_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError(f"pano360_b200 runs on CUDA devices only, got {dev}")
    return dev


@dataclass
class DevicePatch:
    """One warped image resident in HBM.  Unpacks like the reference's patch
    triple ``(warped, mask, irange)`` (stitcher.py:318-319)."""

    rgba: torch.Tensor        # [ph, pw, 4] float32
    invalid: torch.Tensor     # [ph, pw] uint8 (1 = masked)
    box: tuple                # (x0, y0, x1, y1) in mosaic pixels
    index: int = 0

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

    pixels: list                       # u8 [h, w, c] tensors
    luts: list                         # f32 [256] tensors
    hats: dict = field(default_factory=dict)   # (h, w) -> (hat_y, hat_x) f64 tensors
    shapes: list = field(default_factory=list)

    @property
    def nbytes(self):
        return sum(p.numel() for p in self.pixels)


class Compositor:
    """Runs warp / gain / blend stages on one GPU."""

    def __init__(self, device=None):
        self.device = _require_cuda(device)
        _lib.load()
        self._pinned = {}
        self.trace = None      # list of (kernel, algorithmic_bytes, start_event, end_event) when enabled

    # -- plumbing -----------------------------------------------------------
    @property
    def stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _to_device(self, array, pinned_key=None):
        """Host ndarray -> device tensor through a (reused) pinned staging buffer."""
        array = np.ascontiguousarray(array)
        host = torch.from_numpy(array)
        if pinned_key is not None:
            stage, busy = self._pinned.get(pinned_key, (None, None))
            if busy is not None:
                busy.synchronize()            # previous async copy out of this buffer is done
            if stage is None or stage.shape != host.shape or stage.dtype != host.dtype:
                stage = torch.empty(host.shape, dtype=host.dtype, pin_memory=True)
            stage.copy_(host)
            dev = stage.to(self.device, non_blocking=True)
            busy = torch.cuda.Event()
            busy.record(torch.cuda.current_stream(self.device))
            self._pinned[pinned_key] = (stage, busy)
            return dev
        return host.to(self.device)

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

    def upload(self, regions, gains=None, pinned=None):
        """H2D copy of the u8 images (+ LUT / hat tables).  ``pinned`` may be a
        list of pinned uint8 tensors already holding the pixels."""
        src = DeviceSources([], [])
        for i, reg in enumerate(regions):
            img = reg.img if pinned is None else None
            if pinned is not None:
                dev_img = pinned[i].to(self.device, non_blocking=True)
            else:
                if img.dtype != np.uint8 or img.ndim != 3 or img.shape[2] not in (3, 4):
                    raise TypeError("region images must be uint8 HxWx3 (what the reference accepts)")
                host = torch.from_numpy(np.ascontiguousarray(img))
                dev_img = host.to(self.device, non_blocking=host.is_pinned())
            h, w = dev_img.shape[:2]
            src.pixels.append(self.pack_pixels(dev_img))
            src.shapes.append((h, w))
            if (h, w) not in src.hats:
                src.hats[(h, w)] = (self._to_device(geo.hat(h)), self._to_device(geo.hat(w)))
            gain = None if gains is None else gains[i]
            src.luts.append(self._to_device(geo.sample_lut(gain)))
        return src

    def pack_pixels(self, dev_img):
        """u8 x 3 -> u8 x 4 on the device: one aligned 32-bit word per pixel,
        so each bilinear tap of the warp is a single load."""
        if dev_img.shape[2] == 4:
            return dev_img
        h, w = dev_img.shape[:2]
        packed = torch.empty((h, w, 4), dtype=torch.uint8, device=self.device)
        _lib.call("p360_pack_rgbx", _lib.ptr(dev_img), _lib.ptr(packed), h * w, self.stream)
        return packed

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
        lut0 = self._to_device(geo.sample_lut(None))
        hat_y, hat_x = src.hats[(h, w)]
        nblocks = _lib.call("p360_pair_stats_blocks", h, w)
        partial = torch.empty(3 * nblocks, dtype=torch.float64, device=self.device)
        out = torch.zeros((len(todo), 3), dtype=torch.float64, device=self.device)
        for k, (i, j, inv) in enumerate(todo):
            if src.shapes[i] != (h, w) or src.shapes[j] != (h, w):
                raise ValueError("exposure equalisation needs equally sized images (as the reference)")
            inv_c = (C.c_double * 9)(*inv.ravel())
            _lib.call("p360_pair_overlap_stats", _lib.ptr(src.pixels[i]), _lib.ptr(src.pixels[j]),
                      h, w, src.pixels[i].shape[2], _lib.ptr(lut0), _lib.ptr(hat_y), _lib.ptr(hat_x),
                      inv_c, _lib.ptr(partial), out[k].data_ptr(), self.stream)
        sums = out.cpu().numpy()
        for (i, j, _), (cnt, s_i, s_j) in zip(todo, sums):
            sizes[i, j] = sizes[j, i] = cnt
            if cnt > 0:
                overlaps[i, j] = s_i / (3.0 * cnt)
                overlaps[j, i] = s_j / (3.0 * cnt)
        return overlaps, sizes, todo

    # -- K1: warp -------------------------------------------------------------
    def plan_crops(self, regions, plan, proj=geo.SphProj, rows=None, row_align=1, split_dilate=None):
        """Host side of the warp: which (row-cropped, column-split) boxes get
        warped, and their inverse-map tables.  ``rows=(ya, yb)`` crops boxes to
        those mosaic rows; a cropped top edge is moved up to a multiple of
        ``row_align`` rows below the box's true top so that coarse grids
        anchored at the crop coincide with those anchored at the true box.
        With ``split_dilate`` (columns) the all-invalid middle of seam-
        straddling boxes is dropped (``geometry.active_column_runs``): such an
        image yields two crops with the same image index.
        Returns (crops, tables) with crops = [(image, x0, y0, x1, y1, col_off, row_off)]."""
        crops, tabs, total = [], [], 0
        for i, (reg, box) in enumerate(zip(regions, plan.boxes)):
            x0, y0, x1, y1 = box
            ya, yb = (y0, y1) if rows is None else (max(y0, rows[0]), min(y1, rows[1]))
            if ya >= yb or x0 >= x1:
                continue
            ya = y0 + (ya - y0) // row_align * row_align
            runs = [(x0, x1)] if split_dilate is None else \
                geo.active_column_runs(reg, box, plan, proj, dilate=split_dilate)
            for cx0, cx1 in runs:
                col_tab, row_tab = geo.inverse_map_tables(reg, (cx0, ya, cx1, yb), plan, proj)
                crops.append((i, cx0, ya, cx1, yb, total, total + col_tab.size))
                tabs += [col_tab.ravel(), row_tab.ravel()]
                total += col_tab.size + row_tab.size
        return crops, (np.concatenate(tabs) if tabs else np.zeros(0))

    def warp_crops(self, src, crops, tables, origin=(0, 0), owner_state=None):
        """K1 over every crop.  Boxes of the returned patches are relative to
        ``origin`` (x, y).  With ``owner_state = (best, owner, covered)``
        (mosaic-sized, already initialised) the owner-map update of K2 is fused
        into the warp; patch k of the returned list is known as k there."""
        if not crops:
            return []
        dev_tabs = self._to_device(tables, pinned_key="tabs")
        ox, oy = origin
        patches = []
        for k, (i, x0, ya, x1, yb, off_c, off_r) in enumerate(crops):
            pw, ph = x1 - x0, yb - ya
            rgba = torch.empty((ph, pw, 4), dtype=torch.float32, device=self.device)
            invalid = torch.empty((ph, pw), dtype=torch.uint8, device=self.device)
            h, w = src.shapes[i]
            hat_y, hat_x = src.hats[(h, w)]
            if owner_state is None:
                fused = (0, 0, 0, None, None, None, 0)
                nbytes = 17 * pw * ph
            else:
                best, owner, covered = owner_state
                fused = (x0 - ox, ya - oy, k, _lib.ptr(best), _lib.ptr(owner), _lib.ptr(covered), owner.shape[1])
                nbytes = 30 * pw * ph
            self._traced("K1_warp", nbytes, "p360_warp_patch", _lib.ptr(src.pixels[i]), h, w,
                         src.pixels[i].shape[2], _lib.ptr(src.luts[i]), _lib.ptr(hat_y), _lib.ptr(hat_x),
                         dev_tabs.data_ptr() + 8 * off_c, dev_tabs.data_ptr() + 8 * off_r,
                         pw, ph, _lib.ptr(rgba), _lib.ptr(invalid), *fused, self.stream)
            patches.append(DevicePatch(rgba, invalid, (x0 - ox, ya - oy, x1 - ox, yb - oy), i))
        self._keepalive = dev_tabs
        return patches

    def warp(self, regions, src, plan, proj=geo.SphProj, rows=None, row_align=1, split_dilate=None):
        """plan_crops + warp_crops in absolute mosaic coordinates."""
        crops, tables = self.plan_crops(regions, plan, proj, rows, row_align, split_dilate)
        return self.warp_crops(src, crops, tables)

    def new_owner_state(self, shape):
        """(best, owner, covered) for a mosaic (or strip) of ``shape``."""
        h, w = shape
        return (torch.zeros((h, w), dtype=torch.float32, device=self.device),
                torch.full((h, w), -1, dtype=torch.int32, device=self.device),
                torch.zeros((h, w), dtype=torch.uint8, device=self.device))

    # -- blenders (device-resident patches in, device u8 mosaic out) ---------
    def _args(self, p):
        x0, y0, x1, y1 = p.box
        return x1 - x0, y1 - y0, x0, y0

    def blend_none(self, patches, shape):
        """stitcher.py:160-168."""
        h, w = shape
        mosaic = torch.zeros((h, w, 3), dtype=torch.uint8, device=self.device)
        for p in patches:
            pw, ph, x0, y0 = self._args(p)
            _lib.call("p360_paste", _lib.ptr(p.rgba), _lib.ptr(p.invalid), pw, ph, x0, y0,
                      _lib.ptr(mosaic), w, self.stream)
        return mosaic

    def blend_linear(self, patches, shape):
        """stitcher.py:171-183."""
        h, w = shape
        acc = torch.zeros((h, w, 4), dtype=torch.float32, device=self.device)
        for p in patches:
            pw, ph, x0, y0 = self._args(p)
            _lib.call("p360_linear_accumulate", _lib.ptr(p.rgba), _lib.ptr(p.invalid), pw, ph, x0, y0,
                      _lib.ptr(acc), w, self.stream)
        mosaic = torch.empty((h, w, 3), dtype=torch.uint8, device=self.device)
        _lib.call("p360_linear_finalize", _lib.ptr(acc), _lib.ptr(mosaic), h * w, self.stream)
        return mosaic

    def owner_map(self, patches, shape):
        """stitcher.py:196-204 without the H x W x N tensor: (owner, covered)."""
        h, w = shape
        best, owner, covered = self.new_owner_state(shape)
        for k, p in enumerate(patches):
            pw, ph, x0, y0 = self._args(p)
            self._traced("K2_owner_update", 30 * pw * ph, "p360_owner_update", _lib.ptr(p.rgba),
                         _lib.ptr(p.invalid), pw, ph, x0, y0, k, _lib.ptr(best), _lib.ptr(owner),
                         _lib.ptr(covered), w, self.stream)
        return owner, covered

    def blur(self, rgba, sigma, out=None, tmp=None):
        """cv2.GaussianBlur(rgba, (0, 0), sigma) on a device patch."""
        taps = geo.gaussian_taps(sigma)
        out = torch.empty_like(rgba) if out is None else out
        tmp = torch.empty_like(rgba) if tmp is None else tmp
        ph, pw = rgba.shape[:2]
        self._traced("K3_gauss_blur", 32 * pw * ph, "p360_gauss_blur", _lib.ptr(rgba), _lib.ptr(out),
                     _lib.ptr(tmp), pw, ph, taps.ctypes.data_as(C.POINTER(C.c_float)), len(taps),
                     self.stream)
        return out

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

    def coarse_levels(self, patch, k, owner, mosaic_w, n_levels, scratch):
        """Blurred coarse images of levels 0 .. L-2 of one patch
        (stitcher.py:218-229 evaluated at reduced resolution)."""
        pad, plan = geo.coarse_band_plan(n_levels)
        if not plan:
            return pad, [], []
        pw, ph, x0, y0 = self._args(patch)
        dims = (C.c_int32 * 4)()
        _lib.call("p360_pyramid_dims", pw, ph, pad, dims)
        w2, h2, w4, h4 = dims
        d2 = scratch["d2"][:h2 * w2 * 4].view(h2, w2, 4)
        d4 = scratch["d4"][:h4 * w4 * 4].view(h4, w4, 4)
        self._traced("K3a_pyramid_reduce", 25 * pw * ph, "p360_pyramid_reduce", _lib.ptr(patch.rgba), pw, ph,
                     x0, y0, k, _lib.ptr(owner), mosaic_w, pad, _lib.ptr(d2), _lib.ptr(d4), self.stream)
        lows, widths = [], []
        for shift, taps in plan:
            coarse = d2 if shift == 1 else d4
            out = torch.empty_like(coarse)
            tmp = scratch["tmp"][:coarse.numel()].view(coarse.shape)
            self.blur_taps(coarse, taps, out=out, tmp=tmp)
            lows.append(out)
            widths.append(coarse.shape[1])
        return pad, lows, widths

    def blend_multiband(self, patches, shape, n_levels=5, stages=None, owner_state=None):
        """stitcher.py:186-241.  The wide blurs are evaluated on coarse grids
        and every mosaic pixel gathers its bands from the patches covering it,
        in list order, so no mosaic-sized accumulator ever touches HBM."""
        h, w = shape
        if not 1 <= n_levels <= _lib.MAX_LEVELS:
            raise ValueError(f"n_levels must be in 1..{_lib.MAX_LEVELS}")
        if owner_state is None:
            owner, covered = self.owner_map(patches, shape)
        else:
            _, owner, covered = owner_state        # filled by the warp (fused K2)
        mosaic = torch.empty((h, w, 3), dtype=torch.uint8, device=self.device)
        if not patches:
            return mosaic.zero_()
        pad, plan = geo.coarse_band_plan(n_levels)
        scratch = {}
        if plan:
            biggest = max(((p.box[2] - p.box[0] + 2 * pad + 7) * (p.box[3] - p.box[1] + 2 * pad + 7)
                           for p in patches))
            scratch = {"d2": torch.empty(biggest + 64, dtype=torch.float32, device=self.device),
                       "d4": torch.empty(biggest // 4 + 64, dtype=torch.float32, device=self.device),
                       "tmp": torch.empty(biggest + 64, dtype=torch.float32, device=self.device)}
        descs = (_lib.BandPatch * len(patches))()
        keep = []
        for k, p in enumerate(patches):
            pw, ph, x0, y0 = self._args(p)
            _, lows, widths = self.coarse_levels(p, k, owner, w, n_levels, scratch)
            keep.append(lows)
            d = descs[k]
            d.rgba = p.rgba.data_ptr()
            for lvl, (low, lw) in enumerate(zip(lows, widths)):
                d.low[lvl], d.lw[lvl], d.shift[lvl] = low.data_ptr(), lw, plan[lvl][0]
            d.x0, d.y0, d.pw, d.ph, d.pad, d.index = x0, y0, pw, ph, pad, k
        table = np.frombuffer(bytes(descs), dtype=np.uint8)
        dev_table = self._to_device(table, pinned_key="bands")
        gathered = sum((p.box[2] - p.box[0]) * (p.box[3] - p.box[1]) for p in patches)
        self._traced("K4_multiband_collapse", 16 * gathered + 8 * h * w, "p360_multiband_collapse",
                     _lib.ptr(dev_table), len(patches), n_levels, _lib.ptr(owner), _lib.ptr(covered),
                     _lib.ptr(mosaic), h, w, self.stream)
        if stages is not None:
            stages.update(owner=owner, covered=covered, lows=keep)
        self._keepalive2 = (keep, dev_table, scratch)
        return mosaic

    def covered_mask(self, patches, shape):
        """Area of validity for the crop stage (stitcher.py:266-271)."""
        h, w = shape
        covered = torch.zeros((h, w), dtype=torch.uint8, device=self.device)
        for p in patches:
            pw, ph, x0, y0 = self._args(p)
            _lib.call("p360_cover_update", _lib.ptr(p.invalid), pw, ph, x0, y0, _lib.ptr(covered), w,
                      self.stream)
        return covered

    def blend(self, kind, patches, shape, n_levels=5):
        if kind == "none":
            return self.blend_none(patches, shape)
        if kind == "linear":
            return self.blend_linear(patches, shape)
        if kind == "multiband":
            return self.blend_multiband(patches, shape, n_levels)
        raise ValueError(f"unknown blender {kind!r}")

    # -- whole path, device resident ------------------------------------------
    def window_halo(self, kind, n_levels):
        """Rows of context a row window needs on each side: the reach of the
        widest coarse blur (reduce + blur + expand), 0 for pointwise blenders."""
        if kind != "multiband" or n_levels < 2:
            return 0
        return geo.coarse_band_plan(n_levels)[0] + 4

    def composite(self, regions, src, plan, kind, n_levels=5, proj=geo.SphProj, rows=None):
        """warp + blend for the whole mosaic or for a row window [ya, yb)
        (the returned strip has exactly yb - ya rows and is bit-identical to
        those rows of the full composite)."""
        halo = self.window_halo(kind, n_levels)
        if rows is None:
            ya, yb, wa, wb = 0, plan.shape[0], 0, plan.shape[0]
            crops, tables = self.plan_crops(regions, plan, proj, split_dilate=2 * halo)
        else:
            ya, yb = rows
            wa, wb = max(0, ya - halo), min(plan.shape[0], yb + halo)
            crops, tables = self.plan_crops(regions, plan, proj, rows=(wa, wb),
                                            row_align=4 if halo else 1, split_dilate=2 * halo)
        top = min([c[2] for c in crops] + [wa])                # aligned crops may start above wa
        shape = (wb - top, plan.shape[1])
        state = self.new_owner_state(shape) if kind == "multiband" else None
        patches = self.warp_crops(src, crops, tables, origin=(0, top), owner_state=state)
        if kind == "multiband":
            strip = self.blend_multiband(patches, shape, n_levels, owner_state=state)
        else:
            strip = self.blend(kind, patches, shape, n_levels)
        return strip[ya - top:ya - top + (yb - ya)], patches
