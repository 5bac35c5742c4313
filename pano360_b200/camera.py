"""Host-side camera record consumed by the compositing path.

Mirror of the parts of the reference's ``bundle_adj`` module that the hot path
*reads* (bundle_adj.py:18-38 ``Image``/``hom``/``proj``/``_hom_to_from``,
:82-87 ``intrinsics``, :96-101 ``rotation_to_mat``).  Bundle adjustment itself
stays in the reference; ``stitch()`` only duck-types on ``img``/``rot``/``intr``
so objects unpickled from the reference's ``ba_<name>.pkl`` work unchanged.
"""
from __future__ import annotations

import io
import pickle
from dataclasses import dataclass, field

import numpy as np


def _zero_range():
    return (np.zeros(2), np.zeros(2))


@dataclass
class Image:
    """One registered view: pixels + world->camera rotation + intrinsics.

    ``intr`` is expressed in image-centred pixel coordinates (principal point
    ~0); the compositor adds (w/2, h/2) itself (stitcher.py:310).
    """

    img: np.ndarray
    rot: np.ndarray
    intr: np.ndarray
    range: tuple = field(default_factory=_zero_range)

    def hom(self):
        """Centred pixel -> world ray: R^T K^-1 (bundle_adj.py:27-29)."""
        return self.rot.T.dot(np.linalg.inv(self.intr))

    def proj(self):
        """World ray -> centred pixel: K R (bundle_adj.py:31-33)."""
        return self.intr.dot(self.rot)


def hom_to_from(cam_to, cam_from):
    """Homography taking centred pixels of ``cam_from`` into ``cam_to``
    (bundle_adj.py:36-38)."""
    k_r = np.asarray(cam_to.intr).dot(cam_to.rot)
    return k_r.dot(np.asarray(cam_from.rot).T.dot(np.linalg.inv(cam_from.intr)))


def intrinsics(focal, center=(0.0, 0.0)):
    """K = [[f,0,cx],[0,f,cy],[0,0,1]] (bundle_adj.py:82-87; the reference
    uses focal[0] for both axes, so do we)."""
    if isinstance(focal, (list, tuple)):
        focal = focal[0]
    return np.array([[focal, 0.0, center[0]],
                     [0.0, focal, center[1]],
                     [0.0, 0.0, 1.0]])


def rotation_to_mat(rad):
    """Rodrigues formula, exponential map -> 3x3 (bundle_adj.py:96-101)."""
    rad = np.asarray(rad, dtype=np.float64)
    ang = np.linalg.norm(rad)
    axis = rad / ang if ang else rad
    skew = np.array([[0.0, -axis[2], axis[1]],
                     [axis[2], 0.0, -axis[0]],
                     [-axis[1], axis[0], 0.0]])
    return np.eye(3) + skew * np.sin(ang) + (1.0 - np.cos(ang)) * skew.dot(skew)


class _RegionUnpickler(pickle.Unpickler):
    """Loads the reference's ``ba_<name>.pkl`` without its ``bundle_adj``
    module on ``sys.path``: the pickle names ``bundle_adj.Image``; map it
    onto :class:`Image` when the real module cannot be imported."""

    def find_class(self, module, name):
        if module == "bundle_adj" and name == "Image":
            try:
                return super().find_class(module, name)
            except (ImportError, AttributeError):
                return Image
        return super().find_class(module, name)


def load_regions(path_or_bytes):
    """Read a list of regions from a reference-format PKL (stitcher.py:430-432)."""
    if isinstance(path_or_bytes, (bytes, bytearray)):
        return _RegionUnpickler(io.BytesIO(path_or_bytes)).load()
    with open(path_or_bytes, "rb") as fid:
        return _RegionUnpickler(fid).load()
