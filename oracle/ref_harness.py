"""TEST INFRASTRUCTURE — loader for the *live, unmodified* reference.

Only usable where the reference checkout exists (``/root/reference`` in the
build container, or ``baseline/_ref`` if a driver dropped a copy there); it
never travels to the GPU box.  Used by ``oracle/make_golden.py`` to generate
the committed fixtures under ``tests/golden/`` and by the ``not gpu`` tests
that pin ``oracle/restate.py`` against the real thing.

Nothing in the reference tree is edited; all adaptations are applied from the
outside:

* F6  ``cv2.xfeatures2d`` shim so ``import stitcher`` works on OpenCV >= 4.4.
* F7  ``cv2.warpPerspective(..., BORDER_TRANSPARENT)`` gets a zeroed ``dst``
      (the reference reads uninitialised memory in ``equalize_gains``,
      stitcher.py:56-58).
* F13 ``NUMBA_CACHE_DIR`` redirect + ``dont_write_bytecode`` so nothing is
      written into the reference tree.
* F4/F5/F11 ``MAX_RESOLUTION`` / ``n_levels`` / projection are module globals
      looked up at call time, so they are overridden per call and restored.
"""
from __future__ import annotations

import contextlib
import copy
import functools
import os
import sys
import types

import numpy as np

_CANDIDATES = ("/root/reference",
               os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                            "baseline", "_ref"))
_state = {}


def reference_dir():
    for cand in _CANDIDATES:
        if os.path.isfile(os.path.join(cand, "stitcher.py")):
            return cand
    return None


def available():
    return reference_dir() is not None


def load():
    """Import the reference's ``stitcher`` and ``bundle_adj`` (once)."""
    if "stitcher" in _state:
        return _state["stitcher"], _state["bundle_adj"]
    ref = reference_dir()
    if ref is None:
        raise RuntimeError("reference checkout not present (looked in %s)" % (_CANDIDATES,))
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/p360_numba_cache")
    sys.dont_write_bytecode = True
    import cv2
    if not hasattr(cv2, "xfeatures2d"):
        cv2.xfeatures2d = types.SimpleNamespace(SIFT_create=cv2.SIFT_create)
    if not getattr(cv2.warpPerspective, "_p360_zero_dst", False):
        orig_wp = cv2.warpPerspective

        def warp_zero_dst(src, mat, dsize, dst=None, flags=cv2.INTER_LINEAR,
                          borderMode=cv2.BORDER_CONSTANT, borderValue=0):
            if dst is None and borderMode == cv2.BORDER_TRANSPARENT:
                dst = np.zeros((dsize[1], dsize[0]) + src.shape[2:], src.dtype)
            return orig_wp(src, mat, dsize, dst=dst, flags=flags,
                           borderMode=borderMode, borderValue=borderValue)
        warp_zero_dst._p360_zero_dst = True
        cv2.warpPerspective = warp_zero_dst
    sys.path.insert(0, ref)
    try:
        import bundle_adj
        import stitcher
    finally:
        sys.path.remove(ref)
    _state["stitcher"], _state["bundle_adj"] = stitcher, bundle_adj
    return stitcher, bundle_adj


def to_ref_regions(regions):
    """Fresh ``bundle_adj.Image`` objects (the reference mutates its inputs,
    SURVEY.md F12)."""
    _, ba = load()
    return [ba.Image(np.array(r.img, copy=True), np.array(r.rot, dtype=np.float64),
                     np.array(r.intr, dtype=np.float64)) for r in regions]


@contextlib.contextmanager
def _overrides(max_resolution, n_levels, proj):
    st, _ = load()
    saved = (st.MAX_RESOLUTION, st.multiband_blend.__defaults__, st.SphProj)
    st.MAX_RESOLUTION = max_resolution
    st.multiband_blend.__defaults__ = (n_levels,)
    if proj == "cylindrical":
        st.SphProj = st.CylProj
    elif proj != "spherical":
        raise ValueError(proj)
    try:
        yield st
    finally:
        st.MAX_RESOLUTION, st.multiband_blend.__defaults__, st.SphProj = saved


def ref_stitch(regions, blend="none", equalize=False, n_levels=5,
               proj="spherical", max_resolution=1400, crop=False, capture=None):
    """Run the reference's ``stitch()`` (stitcher.py:274) on a copy of
    ``regions``.  If ``capture`` is a dict it receives the ``patches`` list
    and ``shape`` exactly as handed to the blender (pre-blend copies)."""
    with _overrides(max_resolution, n_levels, proj) as st:
        blender = st.BLENDERS[blend]
        if capture is not None:
            inner = blender

            @functools.wraps(inner)
            def spy(patches, shape, *args):
                capture["patches"] = [(w.copy(), m.copy(), r) for w, m, r in patches]
                capture["shape"] = shape
                return inner(patches, shape, *args)
            # keep the identity test at stitcher.py:295 true for multiband
            if blend == "multiband":
                st.multiband_blend = spy
            blender = spy
        try:
            return st.stitch(to_ref_regions(regions), blender=blender,
                             equalize=equalize, crop=crop)
        finally:
            if capture is not None and blend == "multiband":
                st.multiband_blend = inner


def ref_blend(patches, shape, blend="multiband", n_levels=5):
    """Run one of the reference blenders on (copies of) captured patches."""
    st, _ = load()
    patches = [(w.copy(), m.copy(), r) for w, m, r in patches]
    if blend == "multiband":
        return st.multiband_blend(patches, shape, n_levels)
    return st.BLENDERS[blend](patches, shape)


def ref_gains(regions):
    """Gains the reference's ``equalize_gains`` (stitcher.py:36-66) applies:
    returned as the per-image factor recovered from ``find_gains``."""
    st, _ = load()
    regs = to_ref_regions(regions)
    for reg in regs:
        reg.img = st._add_weights(reg.img)
    cap = {}
    orig = st.find_gains

    def spy(overlaps, sizes, *a, **k):
        cap["overlaps"], cap["sizes"] = overlaps.copy(), sizes.copy()
        cap["gains"] = orig(overlaps, sizes, *a, **k)
        return cap["gains"]
    st.find_gains = spy
    try:
        st.equalize_gains(regs)
    finally:
        st.find_gains = orig
    return cap


def deepcopy_regions(regions):
    return copy.deepcopy(regions)
