"""Host-side O(N) geometry of the compositing path, in float64 exactly as the
reference does it: projections (stitcher.py:73-104), per-image angular range
(:107-122), mosaic resolution and extent (:125-157), patch bounding boxes
(:283-297) and the separable inverse-map tables the warp kernel consumes
(:300-306).  Nothing here touches pixels.
"""
from __future__ import annotations

from dataclasses import dataclass
from functools import lru_cache

import numpy as np

PATCH_PAD = 10          # stitcher.py:296-297 (multiband only)
BORDER_SAMPLES = 100    # stitcher.py:109


class SphProj:
    """Forward / backward spherical projection (stitcher.py:73-87)."""

    KIND = "spherical"

    @staticmethod
    def hom2proj(pts):
        horiz = np.sqrt(pts[:, 0] ** 2 + pts[:, 2] ** 2)
        return np.stack([np.arctan2(pts[:, 0], pts[:, 2]), np.arctan2(pts[:, 1], horiz)], axis=-1)

    @staticmethod
    def proj2hom(pts):
        return np.stack([np.sin(pts[:, 0]), np.tan(pts[:, 1]), np.cos(pts[:, 0])], axis=-1)


class CylProj:
    """Forward / backward cylindrical projection (stitcher.py:90-104)."""

    KIND = "cylindrical"

    @staticmethod
    def hom2proj(pts):
        horiz = np.sqrt(pts[:, 0] ** 2 + pts[:, 2] ** 2)
        return np.stack([np.arctan2(pts[:, 0], pts[:, 2]), pts[:, 1] / horiz], axis=-1)

    @staticmethod
    def proj2hom(pts):
        return np.stack([np.sin(pts[:, 0]), pts[:, 1], np.cos(pts[:, 0])], axis=-1)


def hat(size):
    """Triangular 0 - 0.5 - 0 profile (stitcher.py:251-254), float64."""
    return 0.5 - np.abs((np.arange(size) - size / 2) / size)


def image_range_border(shape, hom, proj=SphProj):
    """(min, max) projected angles over 4 x 100 border samples; no wrap-around
    handling, like the reference (stitcher.py:107-122)."""
    h, w = shape
    along_x = np.linspace(0, w, BORDER_SAMPLES)
    along_y = np.linspace(0, h, BORDER_SAMPLES)
    n = BORDER_SAMPLES
    ring = np.empty((4 * n, 3))
    ring[:, 2] = 1.0
    ring[0 * n:1 * n, 0], ring[0 * n:1 * n, 1] = 0.0, along_y
    ring[1 * n:2 * n, 0], ring[1 * n:2 * n, 1] = w, along_y
    ring[2 * n:3 * n, 0], ring[2 * n:3 * n, 1] = along_x, 0.0
    ring[3 * n:4 * n, 0], ring[3 * n:4 * n, 1] = along_x, h
    ring -= np.array([w / 2, h / 2, 0])
    ang = proj.hom2proj(hom.dot(ring.T).T)
    return np.min(ang, axis=0), np.max(ang, axis=0), np.sort(ang[:, 0]), ang[:, 0].reshape(4, n).copy()


def image_range_corners(shape, hom, proj=SphProj):
    """Extent from the four corners, pushed across the +-pi seam if needed
    (stitcher.py:125-139)."""
    h, w = shape
    corners = np.array([[-w / 2, -h / 2, 1], [w / 2, -h / 2, 1],
                        [-w / 2, h / 2, 1], [w / 2, h / 2, 1]])
    ang = proj.hom2proj(hom.dot(corners.T).T)
    x_lo, x_hi = min(ang[0, 0], ang[2, 0]), max(ang[1, 0], ang[3, 0])
    y_lo, y_hi = min(ang[0, 1], ang[1, 1]), max(ang[2, 1], ang[3, 1])
    if x_lo > x_hi:
        x_hi += 2 * np.pi
    if y_lo > y_hi:
        y_hi += np.pi
    return np.array([x_lo, y_lo]), np.array([x_hi, y_hi])


@dataclass
class MosaicPlan:
    """Where every image lands (all integers are mosaic pixels)."""

    shape: tuple              # (H, W)
    resolution: np.ndarray    # rad/px for (theta, phi)
    origin: np.ndarray        # (theta_min, phi_min)
    boxes: list               # per image (x0, y0, x1, y1)
    ranges: list              # per image (min, max) angles
    border_theta: list = None # per image sorted longitudes of the border samples
    _rays: tuple = None       # cached (ray_x[W], ray_z[W], ray_y[H]) of proj2hom
    _runs: dict = None        # cached active_column_runs results
    _crops: dict = None       # cached Compositor.plan_crops results
    border_sides: list = None # per image 4 x BORDER_SAMPLES longitudes in order along each image side

    def rays(self, proj=SphProj):
        """``proj2hom`` evaluated once per mosaic column / row (it is separable:
        stitcher.py:84-87, :101-104): x and z components of the ray depend on
        the column only, the y component on the row only."""
        if self._rays is None or self._rays[0] is not proj:
            height, width = self.shape
            theta = np.arange(width + 1) * self.resolution[0] + self.origin[0]
            phi = np.arange(height + 1) * self.resolution[1] + self.origin[1]
            by_col = proj.proj2hom(np.stack([theta, np.zeros_like(theta)], axis=-1))
            by_row = proj.proj2hom(np.stack([np.zeros_like(phi), phi], axis=-1))
            self._rays = (proj, np.ascontiguousarray(by_col[:, 0]), np.ascontiguousarray(by_col[:, 2]),
                          np.ascontiguousarray(by_row[:, 1]))
        return self._rays[1:]


def plan_mosaic(regions, pad, max_resolution, proj=SphProj):
    """Mosaic shape and per-image boxes (stitcher.py:276-277, :283-297).
    ``pad`` is True for the multiband blender (10-px pad, clamped)."""
    samples = [image_range_border(r.img.shape[:2], r.hom(), proj) for r in regions]
    ranges = [(s[0], s[1]) for s in samples]
    lo = np.min([r[0] for r in ranges], axis=0)
    hi = np.max([r[1] for r in ranges], axis=0)
    mid = regions[len(regions) // 2]
    mid_lo, mid_hi = image_range_corners(mid.img.shape[:2], mid.hom(), proj)
    resolution = (mid_hi - mid_lo) / np.array(mid.img.shape[:2][::-1])
    longest = np.max((hi - lo) / resolution)
    if longest > max_resolution:
        resolution = resolution * (longest / max_resolution)
    target = (hi - lo) / resolution
    shape = tuple(int(t) for t in np.round(target))[::-1]
    limit = target.astype(np.int32)               # truncating cast, stitcher.py:297
    boxes = []
    for r_lo, r_hi in ranges:
        bottom = np.round((r_lo - lo) / resolution).astype(np.int32)
        top = np.round((r_hi - lo) / resolution).astype(np.int32)
        if pad:
            bottom = np.maximum(bottom - PATCH_PAD, np.int32([0, 0]))
            top = np.minimum(top + PATCH_PAD, limit)
        boxes.append((int(bottom[0]), int(bottom[1]), int(top[0]), int(top[1])))
    return MosaicPlan(shape, resolution, lo, boxes, ranges, [s[2] for s in samples], None, {}, {},
                      [s[3] for s in samples])


_plan_cache = {}


def plan_mosaic_cached(regions, pad, max_resolution, proj=SphProj):
    """``plan_mosaic`` remembered by rig geometry (image sizes, rotations, intrinsics): stitching the
    same rig again — the frames of a panoramic video, the steps of a benchmark — costs a hash of
    the camera matrices instead of 400 projected border samples per image, and the plan's own
    caches (column runs, crops, ray tables) stay warm."""
    key = (pad, float(max_resolution), proj,
           tuple((r.img.shape[:2], np.asarray(r.rot, np.float64).tobytes(), np.asarray(r.intr, np.float64).tobytes())
                 for r in regions))
    plan = _plan_cache.get(key)
    if plan is None:
        if len(_plan_cache) >= 16:
            _plan_cache.pop(next(iter(_plan_cache)))
        plan = _plan_cache[key] = plan_mosaic(regions, pad, max_resolution, proj)
    return plan


def active_column_runs(index, box, plan, dilate=0, margin=4, align=4):
    """Column ranges of the box of image ``index`` that can contain valid pixels.

    The reference gives an image that straddles theta = +-pi a full-mosaic-
    width box (no wrap handling in stitcher.py:107-122) of which
    all but the two ends is invalid.  The image footprint is a convex spherical
    quadrilateral, so the columns it touches are exactly the theta-extent of
    its border: if the sorted border samples leave one wide interior gap, every
    column inside the gap is invalid and the box is split in two.  Each part is
    grown by ``dilate`` columns (callers pass twice the reach of the widest
    blur, so that neither the dropped columns nor the reflection at the
    artificial edge can influence a pixel with non-zero weight) plus a small
    ``margin`` for the sampling of the border; the second part starts a
    multiple of ``align`` columns from the box origin.  Returns
    [(x0, x1), ...] inside the box, in ascending order; a single run equal to
    the box if no split."""
    cached = plan._runs.get((index, box, dilate, margin, align)) if plan._runs is not None else None
    if cached is not None:
        return cached
    x0, y0, x1, y1 = box
    theta = plan.border_theta[index]
    gaps = np.diff(theta)
    k = int(np.argmax(gaps))
    left_end = int(np.ceil((theta[k] - plan.origin[0]) / plan.resolution[0])) + margin + dilate
    right_start = int(np.floor((theta[k + 1] - plan.origin[0]) / plan.resolution[0])) - margin - dilate
    left_end, right_start = min(left_end, x1), max(right_start, x0)
    # keep coarse grids anchored at the second run in phase with those of the whole box
    right_start = x0 + (right_start - x0) // align * align
    # The gap between two sorted samples is empty only if the border itself does not cross it
    # between samples: a footprint that contains (or passes close to) a pole of the projection
    # covers a wide range of longitudes with a few border samples — neighbours along a side then
    # lie on both sides of the gap.  (Across the +-pi seam they differ by more than pi: the short
    # arc between them runs through the seam, not through the gap.)
    crossed = False
    if plan.border_sides is not None:
        side = plan.border_sides[index]
        a, b = side[:, :-1], side[:, 1:]
        lo_s, hi_s = np.minimum(a, b), np.maximum(a, b)
        crossed = bool(np.any((hi_s - lo_s <= np.pi) & (lo_s <= theta[k]) & (hi_s >= theta[k + 1])))
    # only worth (and only safe) when the gap dwarfs both the sampling step and the dilation
    if crossed or right_start - left_end < max(256, 4 * dilate) or gaps[k] < 20 * np.median(gaps):
        runs = [(x0, x1)]
    else:
        runs = [(a, b) for a, b in [(x0, left_end), (right_start, x1)] if b > a]
    if plan._runs is not None:
        plan._runs[(index, box, dilate, margin, align)] = runs
    return runs


def source_rect_needed(region, crop, plan, proj=SphProj, tile=(64, 32), tile_mask=None, tile_origin=(0, 0)):
    """Rows [r0, r1) and columns [c0, c1) of ``region.img`` that the warp can touch when it
    produces the mosaic box ``crop = (x0, y0, x1, y1)`` — so that only that rectangle needs to be
    uploaded and packed (a strip of a multi-GPU composite reads a fraction of every image it
    meets; with the seam plan an image is only read where it owns pixels or takes part in a seam).
    Returns (r0, r1, c0, c1); all zero if nothing is read.

    Conservative by construction: interval arithmetic per 64 x 32 tile of the box (ray tables ->
    K R ray -> source position, the arithmetic of the seam plan), widened by the bilinear taps and
    the 1/32-px rounding; positions beyond the image fold back by BORDER_REFLECT (cv2.remap at
    stitcher.py:315-316).  Anything uncertain — a tile not wholly in front of the camera, more
    than one reflection period — means the whole image.  ``tile_mask`` (bool [tiles_y, tiles_x] of a
    grid of ``tile``-sized cells whose cell (0, 0) starts at mosaic position ``tile_origin`` =
    (x, y)) restricts the box to the cells that are set."""
    x0, y0, x1, y1 = crop
    h, w = region.img.shape[:2]
    if x1 <= x0 or y1 <= y0:
        return 0, 0, 0, 0
    ray_x, ray_z, ray_y = plan.rays(proj)
    kr = np.asarray(region.proj(), dtype=np.float64)
    ox, oy = tile_origin
    # cell boundaries of the tile grid inside the box
    xs = np.unique(np.concatenate([[x0], np.arange(ox + (-(-(x0 - ox) // tile[0])) * tile[0], x1, tile[0]), [x1]]))
    ys = np.unique(np.concatenate([[y0], np.arange(oy + (-(-(y0 - oy) // tile[1])) * tile[1], y1, tile[1]), [y1]]))
    xs, ys = xs[(xs >= x0) & (xs <= x1)], ys[(ys >= y0) & (ys <= y1)]

    def per_cell(values, edges):
        starts = edges[:-1]
        return np.minimum.reduceat(values[edges[0]:edges[-1]], starts - edges[0]), \
            np.maximum.reduceat(values[edges[0]:edges[-1]], starts - edges[0])

    def scaled(k, lo, hi):
        return np.minimum(k * lo, k * hi), np.maximum(k * lo, k * hi)

    bx, bz, by = per_cell(ray_x, xs), per_cell(ray_z, xs), per_cell(ray_y, ys)

    def component(row):
        ax, az, ay = scaled(kr[row, 0], *bx), scaled(kr[row, 2], *bz), scaled(kr[row, 1], *by)
        return (ax[0] + az[0])[None, :] + ay[0][:, None], (ax[1] + az[1])[None, :] + ay[1][:, None]

    keep = np.ones((len(ys) - 1, len(xs) - 1), bool)
    if tile_mask is not None:
        ty = np.clip((ys[:-1] - oy) // tile[1], 0, tile_mask.shape[0] - 1)
        tx = np.clip((xs[:-1] - ox) // tile[0], 0, tile_mask.shape[1] - 1)
        keep = tile_mask[np.ix_(ty, tx)]
        if not keep.any():
            return 0, 0, 0, 0
    (px_lo, px_hi), (py_lo, py_hi), (pz_lo, pz_hi) = component(0), component(1), component(2)
    if not np.all(pz_lo[keep] > 1e-9):
        return 0, h, 0, w

    def extent(lo, hi, size):
        q = np.stack([lo[keep] / pz_lo[keep], lo[keep] / pz_hi[keep], hi[keep] / pz_lo[keep], hi[keep] / pz_hi[keep]])
        v_min, v_max = float(q.min()) + size / 2.0 - 2.0, float(q.max()) + size / 2.0 + 2.0
        if not (np.isfinite(v_min) and np.isfinite(v_max)) or v_min < -(size - 1) or v_max > 2 * (size - 1):
            return 0, size
        a, b = int(np.floor(v_min)), int(np.floor(v_max)) + 2
        if v_min < 0:                                   # positions before the image fold back onto 0 .. -v
            a, b = 0, max(b, int(np.ceil(-v_min)) + 2)
        if v_max > size - 1:                            # positions behind it fold back onto 2 size - 1 - v .. size - 1
            a, b = min(a, int(np.floor(2 * size - 1 - v_max)) - 2), size
        return max(a, 0), min(b, size)

    r0, r1 = extent(py_lo, py_hi, h)
    c0, c1 = extent(px_lo, px_hi, w)
    return r0, r1, c0, c1


def source_rows_needed(region, crop, plan, proj=SphProj, tile=(64, 32)):
    """Rows [r0, r1) of ``source_rect_needed``."""
    return source_rect_needed(region, crop, plan, proj, tile)[:2]


def inverse_map_tables(region, box, plan, proj=SphProj):
    """Separable float64 tables for one patch: ``p = K R proj2hom(theta, phi)``
    splits into a per-column part and a per-row part (stitcher.py:300-306).
    The kernels evaluate the same sum from the per-mosaic ray tables
    (``MosaicPlan.rays``); this form is kept for inspection and tests.

    Returns (col_tab [pw,3], row_tab [ph,3]) with p = col_tab[c] + row_tab[r].
    """
    x0, y0, x1, y1 = box
    ray_x, ray_z, ray_y = plan.rays(proj)
    k_r = region.proj()
    col_tab = ray_x[x0:x1, None] * k_r[:, 0][None, :] + ray_z[x0:x1, None] * k_r[:, 2][None, :]
    row_tab = ray_y[y0:y1, None] * k_r[:, 1][None, :]
    return np.ascontiguousarray(col_tab), np.ascontiguousarray(row_tab)


def gaussian_taps(sigma):
    """float32 taps of ``cv2.GaussianBlur(img, (0, 0), sigma)`` on float data:
    ksize = round(8 sigma + 1) | 1, exp(-x^2 / 2 sigma^2) / sum in float64
    (the call at stitcher.py:226)."""
    ksize = int(np.rint(sigma * 8 + 1)) | 1
    x = np.arange(ksize, dtype=np.float64) - (ksize - 1) / 2.0
    k = np.exp(-(x * x) / (2.0 * sigma * sigma))
    return (k / k.sum()).astype(np.float32)


def band_sigma(level):
    """sigma of pyramid level ``level`` (stitcher.py:218)."""
    return float(np.sqrt(2 * level + 1.0) * 4)


@lru_cache(maxsize=None)
def coarse_band_plan(n_levels):
    """How the blurs of levels 0 .. L-2 are evaluated on coarse grids
    (csrc/p360_pyramid.cu): level 0 on the f = 2 grid, higher levels on the
    f = 4 grid, each with sigma' = sqrt(sigma^2 - (f^2-1)/12 - f^2/6) / f —
    the area reduction (box of width f) and the bilinear expansion (triangle of
    half-width f) already contribute that much variance.  Returns
    ``(pad, [(shift, taps), ...])`` where ``pad`` (a multiple of 4) is how far
    the reflected extension of a patch must reach so that no expanded value
    depends on the coarse images' own borders."""
    levels, pad = [], 0
    for lvl in range(max(n_levels - 1, 0)):
        shift = 1 if lvl == 0 else 2
        f = 1 << shift
        sigma = band_sigma(lvl)
        taps = gaussian_taps(np.sqrt(sigma * sigma - (f * f - 1) / 12.0 - f * f / 6.0) / f)
        levels.append((shift, taps))
        pad = max(pad, f * ((len(taps) - 1) // 2 + 2))
    for _, taps in levels:
        taps.setflags(write=False)          # cached: shared by every caller
    return (pad + 3) // 4 * 4, levels


def sample_lut(gain=None):
    """float32 value of each u8 sample as the reference's float image holds it:
    ``u8.astype(f32) / 255`` (stitcher.py:259), then — with ``-e`` —
    ``clip(gain * v, 0, 1)`` evaluated in float64 and stored back into the
    float32 image (stitcher.py:66)."""
    lut = np.arange(256, dtype=np.uint8).astype(np.float32) / 255
    if gain is not None:
        lut = np.clip(np.float64(gain) * lut, 0, 1).astype(np.float32)
    return lut


def pair_homography(reg_i, reg_j, shape):
    """Un-centred pixel homography taking image j into image i's frame, and
    whether the reference skips the pair (a corner behind the camera)
    (stitcher.py:42-55)."""
    h, w = shape
    shift = np.array([[1, 0, w / 2], [0, 1, h / 2], [0, 0, 1]])
    unshift = np.array([[1, 0, -w / 2], [0, 1, -h / 2], [0, 0, 1]])
    k_r_i = reg_i.intr.dot(reg_i.rot)
    back_j = reg_j.rot.T.dot(np.linalg.inv(reg_j.intr))
    hom = shift.dot(k_r_i.dot(back_j)).dot(unshift)
    corners = np.array([[0, 0, 1], [w, 0, 1], [w, h, 1], [0, h, 1]])
    behind = bool(np.any(hom.dot(corners.T).T[:, 2] < 0))
    return hom, behind


def invert3x3(m):
    """Closed-form double-precision 3x3 inverse (what cv::invert does inside
    cv2.warpPerspective)."""
    m = np.asarray(m, dtype=np.float64)
    c00 = m[1, 1] * m[2, 2] - m[1, 2] * m[2, 1]
    c01 = m[1, 0] * m[2, 2] - m[1, 2] * m[2, 0]
    c02 = m[1, 0] * m[2, 1] - m[1, 1] * m[2, 0]
    d = 1.0 / (m[0, 0] * c00 - m[0, 1] * c01 + m[0, 2] * c02)
    return np.array([
        [c00 * d, (m[0, 2] * m[2, 1] - m[0, 1] * m[2, 2]) * d, (m[0, 1] * m[1, 2] - m[0, 2] * m[1, 1]) * d],
        [-c01 * d, (m[0, 0] * m[2, 2] - m[0, 2] * m[2, 0]) * d, (m[0, 2] * m[1, 0] - m[0, 0] * m[1, 2]) * d],
        [c02 * d, (m[0, 1] * m[2, 0] - m[0, 0] * m[2, 1]) * d, (m[0, 0] * m[1, 1] - m[0, 1] * m[1, 0]) * d]])


# ---- ingest: cv2.resize(img, None, fx=1/S, fy=1/S) tables (stitcher.py:418-421) ---------------
def resize_dsize(h, w, f):
    """Size cv2.resize gives for fx = fy = f: saturate_cast<int>(size * f), i.e. round-half-even."""
    return int(np.rint(h * f)), int(np.rint(w * f))


def resize_tables(n_src, n_dst, f, clamp_fraction):
    """First source index and the two 11-bit fixed-point weights per destination index along one
    axis, as OpenCV's resize.cpp builds xofs / ialpha (``clamp_fraction``: an index clamped at an
    image edge gets the fraction 0) and yofs / ibeta (fraction kept, the row loop clips the
    indices).  -> (int32 [n_dst], int16 [n_dst, 2])."""
    d = np.arange(n_dst, dtype=np.float64)
    pos = ((d + 0.5) * (1.0 / f) - 0.5).astype(np.float32)
    first = np.floor(pos).astype(np.int32)
    frac = (pos - first.astype(np.float32)).astype(np.float32)
    if clamp_fraction:
        low, high = first < 0, first >= n_src - 1
        frac[low | high] = 0.0
        first[low], first[high] = 0, n_src - 1
    weights = np.stack([np.float32(1.0) - frac, frac], axis=1).astype(np.float32) * np.float32(2048.0)
    return first, np.clip(np.rint(weights), -32768, 32767).astype(np.int16)
