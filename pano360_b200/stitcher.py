"""Drop-in replacement for the compositing half of the reference's
``stitcher.py``: same names, signatures, module globals and CLI, with the
projection + blending stages running as sm_100a kernels (no CPU fallback).

Reference interface mirrored here (stitcher.py):
  :17       MAX_RESOLUTION            :24-33    find_gains
  :36-66    equalize_gains            :73-104   SphProj / CylProj
  :160-241  no_blend / linear_blend / multiband_blend      :244-248 BLENDERS
  :251-263  _hat / _add_weights       :266-271  _valid
  :274-327  stitch                    :340-369  crop_mosaic     :390-451 main

Registration (features.py, bundle_adj.py) is untouched host code: ``stitch``
takes the ``bundle_adj.Image`` list that ``traverse()`` returns or that the
reference caches in ``ba_<name>.pkl``.

Differences a caller can observe (all friendlier):
  * inputs are not mutated (the reference overwrites ``reg.img``/``reg.range``);
  * blenders also accept device-resident patches (``DevicePatch``) besides the
    reference's ``(warped, mask, irange)`` NumPy triples;
  * ``-e`` never reads uninitialised memory: the overlap warp
    starts from a zero destination, which is what the reference intends.
"""
from __future__ import annotations

import argparse
import logging
import os
import pickle
import time

import numpy as np

from . import geometry as geo
from .camera import hom_to_from as _hom_to_from, load_regions  # noqa: F401
from .compositor import Compositor, DevicePatch
from .geometry import CylProj, SphProj  # noqa: F401  (module globals, looked up at call time)

MAX_RESOLUTION = 1400

# Column windows of the streamed end-to-end pipeline (Compositor.composite_streamed); 0 = off.
# B200, cfg4 (36 x 4000x3000 -> 8819 x 31654): 46.9 ms without; 40.2 ms with 3 ROW windows (the tail
# was the third of the mosaic that needs the bottom row of images, profiles/r02_e2e_probe.log);
# column windows depend on a few images each, so the tail is one narrow window.
STREAM_WINDOWS = int(os.environ.get("P360_STREAM_WINDOWS", "12"))
STREAM_MIN_PIXELS = 1 << 24

_compositors = {}


def _compositor(device=None):
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("pano360_b200.stitcher needs a CUDA device; there is no CPU fallback")
    key = torch.cuda.current_device() if device is None else device
    if key not in _compositors:
        _compositors[key] = Compositor(device)
    return _compositors[key]


# ---------------------------------------------------------------------------
# exposure gains (stitcher.py:24-66)
# ---------------------------------------------------------------------------
def find_gains(overlaps, sizes, stdn=0.1, stdg=2):
    """Gains minimising intensity discrepancies on overlaps — Brown & Lowe
    eq. (29) (stitcher.py:24-33).  N x N host solve."""
    weight_n = (sizes + sizes.T) / (stdn * stdn)
    weight_g = sizes / (stdg * stdg)
    coupling = weight_n * overlaps
    lhs = np.diag(np.sum(coupling * overlaps + weight_g, axis=1)) - coupling * overlaps.T
    return np.linalg.solve(lhs, np.sum(weight_g, axis=1))


def equalize_gains(regions, _src=None):
    """Per-image exposure gains from pairwise overlap statistics computed on
    the GPU (stitcher.py:36-66).  Returns the gain vector; ``stitch`` folds it
    into each image's sample LUT (``clip(g * rgb, 0, 1)`` before sampling)."""
    comp = _compositor()
    src = comp.upload(regions) if _src is None else _src
    logging.debug("Equalizing gain...")
    overlaps, sizes, _ = comp.pair_statistics(regions, src)
    return find_gains(overlaps, sizes)


# ---------------------------------------------------------------------------
# blenders (stitcher.py:160-248)
# ---------------------------------------------------------------------------
def _device_patches(comp, patches):
    import torch
    out = []
    for k, patch in enumerate(patches):
        if isinstance(patch, DevicePatch):
            out.append(patch)
            continue
        warped, mask, (rows, cols) = patch
        if warped.dtype != np.float32 or warped.ndim != 3 or warped.shape[2] != 4:
            raise TypeError("patches must hold float32 HxWx4 images (stitcher.py:259)")
        rgba = torch.from_numpy(np.ascontiguousarray(warped)).to(comp.device)
        invalid = torch.from_numpy(np.ascontiguousarray(mask, dtype=np.uint8)).to(comp.device)
        out.append(DevicePatch(rgba, invalid, (cols.start, rows.start, cols.stop, rows.stop), k))
    return out


def no_blend(patches, shape):
    """Paste the patches to the mosaic without blending (stitcher.py:160-168)."""
    comp = _compositor()
    return comp.blend_none(_device_patches(comp, patches), tuple(shape)).cpu().numpy()


def linear_blend(patches, shape):
    """Linearly blend patches (stitcher.py:171-183)."""
    comp = _compositor()
    return comp.blend_linear(_device_patches(comp, patches), tuple(shape)).cpu().numpy()


def multiband_blend(patches, shape, n_levels=5):
    """Multi-band blending, Brown & Lowe 2007 (stitcher.py:186-241).  Like the
    reference, the alpha channel of the patches is overwritten with the
    owner mask."""
    comp = _compositor()
    dev = _device_patches(comp, patches)
    from .compositor import EXACT_BELOW
    if n_levels > 1 and dev and min(min(p.shape) for p in dev) < EXACT_BELOW:     # (see Compositor.needs_exact)
        return comp.blend_multiband_exact(dev, tuple(shape), n_levels).cpu().numpy()
    return comp.blend_multiband(dev, tuple(shape), n_levels).cpu().numpy()


BLENDERS = {
    "none": no_blend,
    "linear": linear_blend,
    "multiband": multiband_blend,
}


def _blend_kind(blender):
    """Map a blender callable (ours, or a same-named one such as the
    reference's) to a kernel path; None for foreign callables."""
    for kind, ours in BLENDERS.items():
        if blender is ours:
            return kind
    name = getattr(blender, "__name__", "")
    return {"no_blend": "none", "linear_blend": "linear", "multiband_blend": "multiband"}.get(name)


def _hat(size):
    """Triangular function 0-0.5-0 of a given size (stitcher.py:251-254)."""
    return geo.hat(size)


def _add_weights(img):
    """float32 RGBA view of a u8 image with the hat-product weight in alpha
    (stitcher.py:257-263).  Host helper for API parity; the kernels never
    materialise this image."""
    height, width = img.shape[:2]
    out = np.empty((height, width, 4), np.float32)
    out[..., :3] = img.astype(np.float32) / 255
    out[..., 3] = _hat(height)[:, None] * _hat(width)[None, :]
    return out


def _valid(patches, shape):
    """Area of validity, for crop (stitcher.py:266-271)."""
    comp = _compositor()
    return comp.covered_mask(_device_patches(comp, patches), tuple(shape)).cpu().numpy().astype(bool)


# ---------------------------------------------------------------------------
# stitch (stitcher.py:274-327)
# ---------------------------------------------------------------------------
def _download(mosaic_dev, out=None):
    """Device uint8 mosaic -> host ndarray (into ``out`` when given; a pinned
    ``out`` makes the copy a plain DMA)."""
    import torch
    if out is None:
        return mosaic_dev.cpu().numpy()
    if out.shape != tuple(mosaic_dev.shape) or out.dtype != np.uint8 or not out.flags.c_contiguous:
        raise ValueError(f"out must be a C-contiguous uint8 array of shape {tuple(mosaic_dev.shape)}")
    host = torch.from_numpy(out)
    host.copy_(mosaic_dev, non_blocking=host.is_pinned())
    torch.cuda.current_stream(mosaic_dev.device).synchronize()
    return out


def _is_pinned_out(out, shape):
    import torch
    return (out is not None and out.dtype == np.uint8 and out.flags.c_contiguous
            and out.shape == tuple(shape) + (3,) and torch.from_numpy(out).is_pinned())


def stitch(regions, blender=no_blend, equalize=False, crop=False, n_levels=None, out=None):
    """Stitch the images together; returns the uint8 H x W x 3 mosaic.

    Extra optional arguments (defaults keep the reference behaviour):
    ``n_levels`` overrides the band count of the multiband blender (by default
    the blender's own default applies, stitcher.py:321); ``out`` is a
    preallocated uint8 H x W x 3 array (e.g. pinned memory) to receive the
    mosaic."""
    comp = _compositor()
    kind = _blend_kind(blender)
    proj = globals()["SphProj"]            # honours `stitcher.SphProj = stitcher.CylProj`
    levels = 5
    if kind == "multiband":
        levels = n_levels if n_levels is not None else (blender.__defaults__ or (5,))[0]
    plan = geo.plan_mosaic_cached(regions, pad=(kind == "multiband"),
                                  max_resolution=globals()["MAX_RESOLUTION"], proj=proj)
    # where the mosaic lands: the caller's array if it is pinned, else a pinned staging buffer
    # from which the rows are copied (by a few threads) into the caller's / a fresh array as
    # they arrive
    if out is not None and (out.shape != plan.shape + (3,) or out.dtype != np.uint8 or not out.flags.c_contiguous):
        raise ValueError(f"out must be a C-contiguous uint8 array of shape {plan.shape + (3,)}")
    direct_out = _is_pinned_out(out, plan.shape)
    big = plan.shape[0] * plan.shape[1] >= STREAM_MIN_PIXELS
    # views a few blur radii small: owner masks can be slivers the coarse grids cannot resolve
    exact = kind == "multiband" and levels > 1 and comp.needs_exact(regions)

    def landing():
        if direct_out:
            return out, None
        result = out if out is not None else np.empty(plan.shape + (3,), np.uint8)
        return comp.host_stage(plan.shape + (3,)), result

    if STREAM_WINDOWS and kind is not None and not equalize and not crop and big:
        # both PCIe directions at once: row windows are composited and downloaded while the
        # images of the windows below are still being uploaded
        stage, result = landing()
        comp.composite_streamed(regions, plan, kind, levels, proj, stage, windows=STREAM_WINDOWS, exact=exact,
                                copy_to=result)
        comp.finish_download(result, stage)
        comp.release()
        return out if direct_out else result
    src = comp.upload(regions, overlap=not equalize, reuse=True)       # warp starts while late images still upload
    if equalize:
        comp.set_gains(src, equalize_gains(regions, src))
    if kind is None:                       # foreign blender: the reference's one-box-per-image NumPy triples
        patches = comp.warp(regions, src, plan, proj)
        mosaic = blender([p.to_numpy() for p in patches], plan.shape)
    elif crop:
        # the union of valid pixels stays in HBM, the rectangle is found there (K9) and only the
        # cropped mosaic crosses PCIe
        logging.debug("Cropping...")
        mosaic_dev, patches = comp.composite(regions, src, plan, kind, levels, proj, want_covered=True, exact=exact)
        y0, y1, x0, x1 = comp.crop_rect(comp.last_covered[:plan.shape[0]])
        mosaic = mosaic_dev[y0:y1, x0:x1].contiguous().cpu().numpy()
    elif direct_out or big:
        stage, result = landing()
        mosaic_dev, patches = comp.composite(regions, src, plan, kind, levels, proj, out_host=stage, exact=exact)
        comp.finish_download(result, stage)
        mosaic = out if direct_out else result
    else:
        mosaic_dev, patches = comp.composite(regions, src, plan, kind, levels, proj, exact=exact)
        mosaic = _download(mosaic_dev, out)
    if crop and kind is None:
        logging.debug("Cropping...")
        mosaic = crop_mosaic(mosaic, _valid(patches, plan.shape))
    del patches
    comp.release()          # everything has been waited for (the mosaic is on the host)
    return mosaic


# ---------------------------------------------------------------------------
# crop (stitcher.py:340-369) — K9 on the device
# ---------------------------------------------------------------------------
def crop_mosaic(mosaic, valid):
    """Remove the black borders: the largest all-valid axis-aligned rectangle (histogram method,
    stitcher.py:340-369), found on the device (p360_crop_rect) with the reference's scan order
    and strict '>' update, so ties resolve identically — including its quirk that column 0
    never extends to the right (its loop at :359 stops at j = 1).  Returns a view of ``mosaic``."""
    import torch
    comp = _compositor()
    covered = torch.from_numpy(np.ascontiguousarray(valid, dtype=np.uint8)).to(comp.device)
    y0, y1, x0, x1 = comp.crop_rect(covered)
    return mosaic[y0:y1, x0:x1, :]


# ---------------------------------------------------------------------------
# CLI (stitcher.py:372-457)
# ---------------------------------------------------------------------------
def idx_to_keypoints(matches, kpts):
    """Replace keypoint indices with homogeneous coordinates (stitcher.py:372-387)."""
    kpts = [np.concatenate([kp, np.ones((kp.shape[0], 1))], axis=1) for kp in kpts]
    matches = matches.item()
    return {i: {j: (np.concatenate([kpts[i][m[:, 0]], kpts[j][m[:, 1]]], axis=1), h, len(m))
                for j, (m, h) in col.items()} for i, col in matches.items()}


def main(argv=None):
    """Script entry point: same flags and cache files as the reference
    (stitcher.py:390-451).  Registration is delegated to the reference's own
    ``features`` / ``bundle_adj`` modules, which must be importable when the
    caches are cold."""
    import cv2
    parser = argparse.ArgumentParser(description="Stitch images.")
    parser.add_argument('path', type=str, help="directory with the images to process.")
    parser.add_argument("-s", "--shrink", type=float, default=2,
                        help="downsample the images by this amount.")
    parser.add_argument("--ba", default="incr", choices=["none", "incr", "last"],
                        help="bundle adjustment type.")
    parser.add_argument("--equalize", "-e", action="store_true",
                        help="equalize image gain before stitching.")
    parser.add_argument("--crop", "-c", action="store_true", help="remove the black borders.")
    parser.add_argument("--blend", "-b", default='multiband', choices=list(BLENDERS.keys()),
                        help="blending algorithm.")
    parser.add_argument("-o", "--out", type=str, help="save result to this file")
    parser.add_argument("--no-show", action="store_true",
                        help="(extra) do not open a window with the result.")
    args = parser.parse_args(argv)

    from .ingest import read_images
    name = f"{os.path.basename(os.path.normpath(args.path))}_s{args.shrink}"
    # threaded decode; the `-s` shrink (cv2.resize at stitcher.py:420-421) runs on the device, bit-exact
    imgs = read_images(args.path, args.shrink)

    try:
        regions = load_regions(f"ba_{name}.pkl")
    except IOError:
        from bundle_adj import traverse          # the reference's, untouched
        try:
            arr = np.load(f"matches_{name}.npz", allow_pickle=True)
            kpts, matches = arr['kpts'], arr['matches']
        except IOError:
            from features import matching        # the reference's, untouched
            kpts, matches = matching(imgs)
            np.savez(f"matches_{name}.npz", kpts=kpts, matches=matches)
        start = time.time()
        regions = traverse(imgs, idx_to_keypoints(matches, kpts), badjust=args.ba)
        logging.info(f"Image registration, time: {time.time() - start}")
        with open(f"ba_{name}.pkl", 'wb') as fid:
            pickle.dump(regions, fid, protocol=pickle.HIGHEST_PROTOCOL)

    start = time.time()
    mosaic = stitch(regions, blender=BLENDERS[args.blend], equalize=args.equalize, crop=args.crop)
    logging.info(f"Built mosaic, time: {time.time() - start}")

    if args.out:
        cv2.imwrite(args.out, mosaic)
    if not args.no_show:
        try:
            cv2.imshow("Mosaic", mosaic)
            if cv2.waitKey(0) & 0xff == 27:
                cv2.destroyAllWindows()
        except cv2.error:
            logging.info("no display available; skipping imshow")
    return mosaic


if __name__ == '__main__':
    logging.basicConfig(level=logging.DEBUG)
    main()
