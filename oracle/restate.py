"""TEST INFRASTRUCTURE — CPU restatement of pano360's compositing hot path.

This is the *oracle*: a NumPy (+ the same OpenCV calls the reference makes)
restatement of ``stitcher.stitch()`` and its three blenders, written to be
bit-identical on the uint8 mosaic to the unmodified reference while never
materialising the reference's ``H x W x N`` weights tensor (stitcher.py:196),
so that it also runs at sizes where the reference itself does not fit in RAM
(SURVEY.md F10).  Each function cites the reference lines it follows.

Pinning: ``tests/test_oracle_vs_reference.py`` runs this file against the
live reference (``oracle/ref_harness.py``) for every blender / projection /
gain combination, and ``tests/golden/*.npz`` (made by ``oracle/make_golden.py``
from the live reference) pins it where the reference checkout is absent.
The reference's own tests hold no vectors for this path, so
those differential fixtures are the pin.

Only ``tests/``, ``__graft_entry__.smoke()`` and the CPU legs of ``bench.py``
may import this module.  The product path (``pano360_b200``) never does.

``backend="cv2"`` calls ``cv2.remap`` / ``cv2.GaussianBlur`` /
``cv2.warpPerspective`` exactly where the reference does; ``backend="numpy"``
swaps in the pure-NumPy restatements of ``oracle/cv_semantics.py``.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import cv_semantics as cvs

PAD = 10                     # stitcher.py:296-297
N_BORDER = 100               # stitcher.py:109


# --------------------------------------------------------------------------
# projections (stitcher.py:73-104)
# --------------------------------------------------------------------------
def _angles_sph(v):
    return np.stack([np.arctan2(v[:, 0], v[:, 2]),
                     np.arctan2(v[:, 1], np.sqrt(v[:, 0] ** 2 + v[:, 2] ** 2))], axis=-1)


def _rays_sph(a):
    return np.stack([np.sin(a[:, 0]), np.tan(a[:, 1]), np.cos(a[:, 0])], axis=-1)


def _angles_cyl(v):
    return np.stack([np.arctan2(v[:, 0], v[:, 2]),
                     v[:, 1] / np.sqrt(v[:, 0] ** 2 + v[:, 2] ** 2)], axis=-1)


def _rays_cyl(a):
    return np.stack([np.sin(a[:, 0]), a[:, 1], np.cos(a[:, 0])], axis=-1)


PROJ = {"spherical": (_angles_sph, _rays_sph), "cylindrical": (_angles_cyl, _rays_cyl)}


# --------------------------------------------------------------------------
# weights (stitcher.py:251-263)
# --------------------------------------------------------------------------
def hat(n):
    """0 .. 0.5 .. 1/n triangular profile (stitcher.py:251-254)."""
    return 0.5 - np.abs((np.arange(n) - n / 2) / n)


def rgba_with_weights(img_u8):
    """u8 HxWx3 -> f32 HxWx4 with alpha = hat(h) (x) hat(w) (stitcher.py:257-263)."""
    h, w = img_u8.shape[:2]
    out = np.empty((h, w, 4), np.float32)
    out[..., :3] = img_u8.astype(np.float32) / 255
    out[..., 3] = hat(h)[:, None] * hat(w)[None, :]
    return out


# --------------------------------------------------------------------------
# geometry (stitcher.py:107-157, :283-297)
# --------------------------------------------------------------------------
def border_range(shape, hom, proj="spherical"):
    """Angular extent from 4 x 100 border samples (stitcher.py:107-122)."""
    h, w = shape
    sx, sy = np.linspace(0, w, N_BORDER), np.linspace(0, h, N_BORDER)
    one, zero = np.ones(N_BORDER), np.zeros(N_BORDER)
    pts = np.concatenate([np.stack([zero, sy, one], 1), np.stack([zero + w, sy, one], 1),
                          np.stack([sx, zero, one], 1), np.stack([sx, zero + h, one], 1)])
    pts = pts - np.array([w / 2, h / 2, 0])
    ang = PROJ[proj][0](hom.dot(pts.T).T)
    return ang.min(axis=0), ang.max(axis=0)


def corner_range(shape, hom, proj="spherical"):
    """Extent from the 4 corners with wrap-around push (stitcher.py:125-139)."""
    h, w = shape
    pts = np.array([[-w / 2, -h / 2, 1], [w / 2, -h / 2, 1], [-w / 2, h / 2, 1], [w / 2, h / 2, 1]])
    ang = PROJ[proj][0](hom.dot(pts.T).T)
    lo_x, hi_x = min(ang[0, 0], ang[2, 0]), max(ang[1, 0], ang[3, 0])
    lo_y, hi_y = min(ang[0, 1], ang[1, 1]), max(ang[2, 1], ang[3, 1])
    if lo_x > hi_x:
        hi_x += 2 * np.pi
    if lo_y > hi_y:
        hi_y += np.pi
    return np.array([lo_x, lo_y]), np.array([hi_x, hi_y])


@dataclass
class Plan:
    shape: tuple          # (H, W)
    resolution: np.ndarray
    origin: np.ndarray    # im_range[0]
    boxes: list           # per region (x0, y0, x1, y1)


def _hom(reg):
    return reg.rot.T.dot(np.linalg.inv(reg.intr))


def plan(regions, blend, max_resolution=1400, proj="spherical"):
    """Mosaic shape + per-image bounding boxes (stitcher.py:276-277, :142-157,
    :283-297)."""
    ranges = [border_range(r.img.shape[:2], _hom(r), proj) for r in regions]
    lo = np.min([r[0] for r in ranges], axis=0)
    hi = np.max([r[1] for r in ranges], axis=0)
    mid = regions[len(regions) // 2]
    mid_lo, mid_hi = corner_range(mid.img.shape[:2], _hom(mid), proj)
    res = (mid_hi - mid_lo) / np.array(mid.img.shape[:2][::-1])
    longest = np.max((hi - lo) / res)
    if longest > max_resolution:
        res = res * (longest / max_resolution)
    target = (hi - lo) / res
    shape = tuple(int(t) for t in np.round(target))[::-1]
    boxes = []
    for r_lo, r_hi in ranges:
        bot = np.round((r_lo - lo) / res).astype(np.int32)
        top = np.round((r_hi - lo) / res).astype(np.int32)
        if blend == "multiband":
            bot = np.maximum(bot - PAD, np.int32([0, 0]))
            top = np.minimum(top + PAD, target.astype(np.int32))
        boxes.append((int(bot[0]), int(bot[1]), int(top[0]), int(top[1])))
    return Plan(shape, res, lo, boxes)


# --------------------------------------------------------------------------
# warp (stitcher.py:299-319)
# --------------------------------------------------------------------------
def inverse_map(reg, box, pl, proj="spherical", rows=None):
    """Source coordinates + invalid mask for the rows ``rows=(r0, r1)`` of a
    patch box (stitcher.py:300-312)."""
    x0, y0, x1, y1 = box
    r0, r1 = rows if rows is not None else (0, y1 - y0)
    h, w = reg.img.shape[:2]
    yi, xi = np.indices((r1 - r0, x1 - x0))
    ang_x = (xi + x0) * pl.resolution[0] + pl.origin[0]
    ang_y = (yi + (y0 + r0)) * pl.resolution[1] + pl.origin[1]
    rays = PROJ[proj][1](np.stack([ang_x, ang_y], axis=-1).reshape(-1, 2))
    pix = (reg.intr.dot(reg.rot)).dot(rays.T).T.astype(np.float32)
    pix = pix.reshape(r1 - r0, x1 - x0, 3)
    invalid = pix[..., 2] < 0
    with np.errstate(divide="ignore", invalid="ignore"):
        xy = pix[..., :2] / pix[..., 2:3] + np.float32([w / 2, h / 2])
    invalid |= (xy[..., 0] < 0) | (xy[..., 0] > w - 1) | (xy[..., 1] < 0) | (xy[..., 1] > h - 1)
    return xy[..., 0], xy[..., 1], invalid


def warp_region(rgba, reg, box, pl, proj="spherical", backend="cv2", rows=None,
                chunk=512):
    """One patch ``(warped f32 ph x pw x 4, invalid bool)`` (stitcher.py:299-317).
    Row-chunked so seam-straddling 100-Mpix patches stay in memory."""
    x0, y0, x1, y1 = box
    r0, r1 = rows if rows is not None else (0, y1 - y0)
    warped = np.empty((r1 - r0, x1 - x0, 4), np.float32)
    invalid = np.empty((r1 - r0, x1 - x0), bool)
    for a in range(r0, r1, chunk):
        b = min(a + chunk, r1)
        mx, my, bad = inverse_map(reg, box, pl, proj, (a, b))
        if backend == "cv2":
            import cv2
            part = cv2.remap(rgba, mx, my, cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT)
        else:
            part = cvs.remap_linear_reflect(rgba, mx, my)
        part[..., 3] = part[..., 3] * (~bad)
        warped[a - r0:b - r0], invalid[a - r0:b - r0] = part, bad
    return warped, invalid


# --------------------------------------------------------------------------
# blenders (stitcher.py:160-241).  patches = [(warped, invalid, (sy, sx))]
# --------------------------------------------------------------------------
def paste(patches, shape):
    """Last valid writer wins, truncating cast (stitcher.py:160-168)."""
    out = np.zeros(shape + (3,), np.uint8)
    for warped, invalid, where in patches:
        px = (255 * warped[..., :3]).astype(np.uint8)
        out[where] = np.where(invalid[..., None], out[where], px)
    return out


def feather(patches, shape):
    """Weighted average with the hat weights (stitcher.py:171-183)."""
    acc = np.zeros(shape + (3,), np.float32)
    norm = np.zeros(shape, np.float32)
    for warped, invalid, where in patches:
        acc[where] += np.where(invalid[..., None], 0.0, warped[..., :3]) * warped[..., 3:4]
        norm[where] += warped[..., 3]
    norm[norm == 0] = 1
    acc /= norm[..., None]
    return (255 * acc).astype(np.uint8)


def owner_map(patches, shape, mode="auto"):
    """Index of the image with the largest weight per mosaic pixel, first
    maximum wins, -1 where no weight is positive (stitcher.py:196-204).
    ``stack`` builds the reference's H x W x N tensor; ``stream`` keeps a
    running maximum (strict > keeps the first maximum) — identical results."""
    n = len(patches)
    if mode == "auto":
        mode = "stack" if shape[0] * shape[1] * n * 4 < (6 << 30) else "stream"
    if mode == "stack":
        wts = np.zeros(shape + (n,), np.float32)
        for i, (warped, _, (sy, sx)) in enumerate(patches):
            wts[sy, sx, i] = warped[..., 3]
        any_pos = np.sum(wts, axis=-1) > 0
        own = wts.argmax(axis=-1)
        own[~any_pos] = -1
        return own
    best = np.zeros(shape, np.float32)
    own = np.full(shape, -1, np.int64)
    for i, (warped, _, where) in enumerate(patches):
        better = warped[..., 3] > best[where]
        own[where] = np.where(better, i, own[where])
        best[where] = np.where(better, warped[..., 3], best[where])
    return own


def band_sigma(level):
    return np.sqrt(2 * level + 1.0) * 4      # stitcher.py:218


def multiband(patches, shape, n_levels=5, backend="cv2", owner_mode="auto", stages=None):
    """Brown-Lowe multi-band blend on full-resolution DoG bands
    (stitcher.py:186-241).  Mutates the alpha channel of ``patches`` like the
    reference does.  ``stages`` (dict) receives intermediate arrays."""
    own = owner_map(patches, shape, owner_mode)
    for i, (warped, _, where) in enumerate(patches):
        warped[..., 3] = own[where] == i
    covered = np.zeros(shape, bool)
    out = np.zeros(shape + (3,), np.float32)
    prev = [None] * len(patches)
    if stages is not None:
        stages["owner"] = own
        stages["levels"] = []
    for lvl in range(n_levels):
        sigma = band_sigma(lvl)
        band_sum = np.zeros(shape + (3,), np.float32)
        wt_sum = np.zeros(shape, np.float32)
        last = lvl == n_levels - 1
        for i, (warped, invalid, where) in enumerate(patches):
            tile = prev[i] if prev[i] is not None else warped.copy()
            if not last:
                if backend == "cv2":
                    import cv2
                    blur = cv2.GaussianBlur(warped, (0, 0), sigma)
                else:
                    blur = cvs.gaussian_blur(warped, sigma)
                tile[..., :3] -= blur[..., :3]
                tile[..., 3] = blur[..., 3]
                prev[i] = blur
            band_sum[where] += tile[..., :3] * tile[..., 3:4]
            wt_sum[where] += tile[..., 3]
            if lvl == 0:
                covered[where] |= ~invalid
        band_sum[~covered, :] = 0
        wt_sum[wt_sum == 0] = 1
        if stages is not None:
            stages["levels"].append((band_sum.copy(), wt_sum.copy()))
        out += band_sum / wt_sum[..., None]
    if stages is not None:
        stages["covered"] = covered
    return (255 * np.clip(out, 0.0, 1.0)).astype(np.uint8)


BLENDERS = {"none": paste, "linear": feather, "multiband": multiband}


# --------------------------------------------------------------------------
# exposure gains (stitcher.py:24-66)
# --------------------------------------------------------------------------
def solve_gains(overlaps, sizes, stdn=0.1, stdg=2):
    """Brown-Lowe eq. (29) normal equations (stitcher.py:24-33)."""
    n1 = (sizes + sizes.T) / (stdn * stdn)
    n2 = sizes / (stdg * stdg)
    lhs = np.diag(np.sum(n1 * overlaps * overlaps + n2, axis=1)) - n1 * overlaps * overlaps.T
    return np.linalg.solve(lhs, np.sum(n2, axis=1))


def pair_statistics(regions, rgba, backend="cv2"):
    """Overlap pixel counts and mean intensities for every image pair
    (stitcher.py:38-63), destination zero-initialised."""
    n = len(regions)
    overlaps, sizes = np.zeros((n, n)), np.zeros((n, n))
    h, w = rgba[0].shape[:2]
    shift = np.array([[1, 0, w / 2], [0, 1, h / 2], [0, 0, 1]])
    unshift = np.array([[1, 0, -w / 2], [0, 1, -h / 2], [0, 0, 1]])
    corners = np.array([[0, 0, 1], [w, 0, 1], [w, h, 1], [0, h, 1]])
    for i in range(n):
        for j in range(i + 1, n):
            k_r_i = regions[i].intr.dot(regions[i].rot)
            back_j = regions[j].rot.T.dot(np.linalg.inv(regions[j].intr))
            hom = shift.dot(k_r_i.dot(back_j)).dot(unshift)
            if np.any(hom.dot(corners.T).T[:, 2] < 0):
                continue
            if backend == "cv2":
                import cv2
                seen = cv2.warpPerspective(rgba[j], hom, (w, h), dst=np.zeros((h, w, 4), np.float32),
                                           borderMode=cv2.BORDER_TRANSPARENT)
            else:
                seen, _ = cvs.warp_perspective_transparent(rgba[j], hom, w, h)
            both = seen[..., 3] != 0
            sizes[i, j] = sizes[j, i] = np.sum(both)
            if sizes[i, j] == 0:
                continue
            overlaps[i, j] = np.mean(rgba[i][both, :3])
            overlaps[j, i] = np.mean(seen[both, :3])
    return overlaps, sizes


def equalize(regions, rgba, backend="cv2"):
    """Apply the solved gains in place to the float sources (stitcher.py:65-66)."""
    overlaps, sizes = pair_statistics(regions, rgba, backend)
    gains = solve_gains(overlaps, sizes)
    for img, g in zip(rgba, gains):
        img[..., :3] = np.clip(g * img[..., :3], 0, 1)
    return gains, overlaps, sizes


# --------------------------------------------------------------------------
# crop (stitcher.py:266-271, :340-369)
# --------------------------------------------------------------------------
def valid_mask(patches, shape):
    """Area of validity (stitcher.py:266-271)."""
    valid = np.zeros(shape, dtype=bool)
    for _, invalid, irange in patches:
        valid[irange] |= ~invalid
    return valid


def _crop_rect(valid):
    """The scan of stitcher.py:346-367, statement by statement (jitted with Numba when it is
    installed, like the reference's try_jit at :330-337): rows [y0, y1) x columns [x0, x1)."""
    height, width = valid.shape
    heights = np.zeros(width, dtype=np.int32)
    lefts = np.zeros(width, dtype=np.int32)
    rights = np.zeros(width, dtype=np.int32)
    area = 0
    ll = rr = hh = last = 0
    for i in range(height):
        for j in range(width):
            heights[j] = (heights[j] + 1) if valid[i, j] else 0
        for j in range(width):
            lefts[j] = j
            while lefts[j] > 0 and heights[j] <= heights[lefts[j] - 1]:
                lefts[j] = lefts[lefts[j] - 1]
        for j in range(width - 1, 0, -1):          # (rights[0] is never written: it stays 0)
            rights[j] = j
            while rights[j] < width - 1 and heights[j] <= heights[rights[j] + 1]:
                rights[j] = rights[rights[j] + 1]
        for j in range(width):
            new_area = (rights[j] - lefts[j] + 1) * heights[j]
            if new_area > area:
                area = new_area
                ll, rr, hh, last = lefts[j], rights[j], heights[j], i
    if area == 0:
        return 0, 0, 0, 0
    return last - hh + 1, last + 1, ll, rr + 1


try:
    import numba as _numba
    _crop_rect = _numba.njit(cache=False)(_crop_rect)
except ImportError:                                  # pure Python then: small masks only
    pass


def crop_rect(valid):
    return tuple(int(v) for v in _crop_rect(np.ascontiguousarray(valid, dtype=np.bool_)))


# --------------------------------------------------------------------------
# whole path (stitcher.py:274-327)
# --------------------------------------------------------------------------
def build_patches(regions, blend="none", equalize_gains=False, max_resolution=1400,
                  proj="spherical", backend="cv2", window=None, halo=0):
    """(patches, plan).  With ``window=(y0, y1, x0, x1)`` every patch is cropped
    to the window grown by ``halo`` rows/cols (clipped to its true box, so
    reflections still happen at true box edges) and slices are returned in
    window coordinates."""
    pl = plan(regions, blend, max_resolution, proj)
    lazy = window is not None and not equalize_gains      # only convert the images the window touches
    rgba = [None if lazy else rgba_with_weights(r.img) for r in regions]
    if equalize_gains:
        equalize(regions, rgba, backend)
    patches = []
    for k, (reg, box) in enumerate(zip(regions, pl.boxes)):
        x0, y0, x1, y1 = box
        if window is None:
            warped, invalid = warp_region(rgba[k], reg, box, pl, proj, backend)
            patches.append((warped, invalid, np.s_[y0:y1, x0:x1]))
            continue
        wy0, wy1, wx0, wx1 = window
        cy0, cy1 = max(y0, wy0 - halo), min(y1, wy1 + halo)
        cx0, cx1 = max(x0, wx0 - halo), min(x1, wx1 + halo)
        if cy0 >= cy1 or cx0 >= cx1:
            continue
        src = rgba[k] if rgba[k] is not None else rgba_with_weights(reg.img)
        sub = (cx0, y0, cx1, y1)
        warped, invalid = warp_region(src, reg, sub, pl, proj, backend, rows=(cy0 - y0, cy1 - y0))
        oy, ox = wy0 - halo, wx0 - halo
        patches.append((warped, invalid, np.s_[cy0 - oy:cy1 - oy, cx0 - ox:cx1 - ox]))
    return patches, pl


def stitch(regions, blend="none", equalize_gains=False, n_levels=5, max_resolution=1400,
           proj="spherical", backend="cv2", owner_mode="auto", stages=None, crop=False):
    """uint8 H x W x 3 mosaic; inputs are not modified."""
    patches, pl = build_patches(regions, blend, equalize_gains, max_resolution, proj, backend)
    if stages is not None:
        stages["plan"] = pl
        stages["patches"] = [(w.copy(), m.copy(), s) for w, m, s in patches]
    valid = valid_mask(patches, pl.shape) if crop else None      # (multiband overwrites nothing of the masks)
    if blend == "multiband":
        mosaic = multiband(patches, pl.shape, n_levels, backend, owner_mode, stages)
    else:
        mosaic = BLENDERS[blend](patches, pl.shape)
    if crop:                                                     # stitcher.py:322-325
        y0, y1, x0, x1 = crop_rect(valid)
        mosaic = mosaic[y0:y1, x0:x1]
    return mosaic


def stitch_window(regions, window, blend="multiband", equalize_gains=False, n_levels=5,
                  max_resolution=1e9, proj="spherical", backend="cv2"):
    """Exact mosaic pixels inside ``window=(y0, y1, x0, x1)`` without building
    the rest of the mosaic (for parity at sizes the reference cannot hold).
    The halo is the largest blur radius, so every value inside the window sees
    the same neighbourhood (and the same true-edge reflections) as in the full
    mosaic."""
    wy0, wy1, wx0, wx1 = window
    halo = 0
    if blend == "multiband" and n_levels > 1:
        halo = (cvs.gaussian_ksize(band_sigma(n_levels - 2)) - 1) // 2
    patches, _ = build_patches(regions, blend, equalize_gains, max_resolution, proj, backend,
                               window=window, halo=halo)
    shape = (wy1 - wy0 + 2 * halo, wx1 - wx0 + 2 * halo)
    if blend == "multiband":
        full = multiband(patches, shape, n_levels, backend, "stream")
    else:
        full = BLENDERS[blend](patches, shape)
    return full[halo:halo + wy1 - wy0, halo:halo + wx1 - wx0]
