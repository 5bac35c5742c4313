"""Synthetic panorama inputs.

N pinhole views are rendered from a random-texture equirectangular
environment map with known focal length and rotations, then perturbed
photometrically (gain / offset / noise) and geometrically (small rotation
jitter) so that blending and interpolation are actually exercised.

The output is a list of :class:`pano360_b200.camera.Image` — the same record
``bundle_adj.traverse()`` produces / the reference caches in ``ba_<name>.pkl``
(bundle_adj.py:18-33).  Pure host code (NumPy + cv2); used by the tests, by
``bench.py`` and by the golden-fixture generator.
"""
from __future__ import annotations

from dataclasses import dataclass, replace

import cv2
import numpy as np

from .camera import Image, intrinsics, rotation_to_mat


@dataclass(frozen=True)
class Workload:
    """One BASELINE.json configuration (SURVEY.md §8 table / §8(d) inputs)."""

    name: str
    width: int
    height: int
    focal: float
    yaws: tuple            # radians, one per view (view order = list order)
    pitches: tuple         # radians, one per view
    blend: str             # none | linear | multiband
    n_levels: int = 5
    equalize: bool = False
    env_shape: tuple = (1024, 2048)
    max_resolution: float = 1e9
    env_seed: int = 0
    view_seed: int = 7

    @property
    def n_views(self):
        return len(self.yaws)


def _ring(n, step, centre=0.0):
    return tuple(centre + step * (i - (n - 1) / 2.0) for i in range(n))


def workload(name: str, scale: float = 1.0, **over) -> Workload:
    """The five BASELINE.json configs; ``scale`` shrinks w, h and f together
    (same field of view, fewer pixels) for fast parity cases."""
    if name == "cfg1":    # 4-view 640x480 spherical, multiband 5 bands
        wl = Workload("cfg1", 640, 480, 600.0, _ring(4, 0.35), (0.0,) * 4,
                      "multiband", 5, False, (1024, 2048), 1400)
    elif name == "cfg2":  # 8-view 1080p, linear + exposure gain
        wl = Workload("cfg2", 1920, 1080, 1800.0, _ring(8, 0.35), (0.0,) * 8,
                      "linear", 5, True, (2048, 4096))
    elif name == "cfg3":  # 12-view 4000x3000, 2 pitch rows x 6 yaw, 6 bands
        yaws = _ring(6, 0.5) * 2
        pitches = (-0.26,) * 6 + (0.26,) * 6
        wl = Workload("cfg3", 4000, 3000, 4775.0, yaws, pitches,
                      "multiband", 6, False, (4096, 8192))
    elif name == "cfg4":  # 36-view full ring, 3 pitch rows x 12 yaw (30 deg, half-step offset)
        step = np.pi / 6
        yaws = _ring(12, step) * 3
        pitches = (-0.5236,) * 12 + (0.0,) * 12 + (0.5236,) * 12
        wl = Workload("cfg4", 4000, 3000, 4775.0, yaws, pitches,
                      "multiband", 5, False, (4096, 8192))
    elif name == "cfg5":  # one of the 64 six-view 1080p panoramas
        wl = Workload("cfg5", 1920, 1080, 1800.0, _ring(6, 0.35), (0.0,) * 6,
                      "multiband", 5, False, (2048, 4096))
    else:
        raise ValueError(f"unknown workload {name!r}")
    if scale != 1.0:
        wl = replace(wl, width=int(round(wl.width / scale)),
                     height=int(round(wl.height / scale)),
                     focal=wl.focal / scale,
                     env_shape=(max(256, int(wl.env_shape[0] / min(scale, 4))),
                                max(512, int(wl.env_shape[1] / min(scale, 4)))))
    if over:
        wl = replace(wl, **over)
    return wl


def make_env_map(shape=(1024, 2048), seed=0):
    """Equirectangular u8 texture: smooth low-frequency colour + white noise."""
    he, we = shape
    rng = np.random.default_rng(seed)
    low = rng.random((max(he // 16, 2), max(we // 16, 2), 3), dtype=np.float32)
    env = cv2.resize(low, (we, he), interpolation=cv2.INTER_CUBIC)
    env += 0.15 * rng.random((he, we, 3), dtype=np.float32)
    env -= env.min()
    env *= 255.0 / env.max()
    return env.astype(np.uint8)


def camera_rotation(pitch, yaw):
    """World->camera rotation of a view looking at (yaw, pitch)."""
    return rotation_to_mat([pitch, 0.0, 0.0]) @ rotation_to_mat([0.0, yaw, 0.0])


def render_view(env, width, height, focal, rot):
    """Sample the env map along the rays of a pinhole camera (bilinear, wrap
    in longitude).  Ray convention = ``Image.hom()`` (bundle_adj.py:27-29);
    angles = ``SphProj.hom2proj`` (stitcher.py:77-81)."""
    he, we = env.shape[:2]
    xs = np.arange(width, dtype=np.float64) - width / 2.0
    ys = np.arange(height, dtype=np.float64) - height / 2.0
    hom = rot.T @ np.linalg.inv(intrinsics(focal))
    # ray = hom @ (x, y, 1): separable in x and y
    rx = hom[:, 0][:, None] * xs[None, :]            # 3 x W
    ry = hom[:, 1][:, None] * ys[None, :] + hom[:, 2][:, None]   # 3 x H
    ray = rx[:, None, :] + ry[:, :, None]            # 3 x H x W
    theta = np.arctan2(ray[0], ray[2])
    phi = np.arctan2(ray[1], np.hypot(ray[0], ray[2]))
    u = ((theta + np.pi) / (2 * np.pi) * we).astype(np.float32)
    v = ((phi + np.pi / 2) / np.pi * he).astype(np.float32)
    # one wrapped column each side so bilinear taps across the seam are right
    padded = np.concatenate([env[:, -1:], env, env[:, :1]], axis=1)
    return cv2.remap(padded, u + 1.0, v, cv2.INTER_LINEAR,
                     borderMode=cv2.BORDER_REPLICATE)


def make_views(wl: Workload, noise=0.0, jitter=2e-3, photometric=True,
               env=None, only=None):
    """Render all views of a workload and return ``list[Image]``.

    Perturbations use ``default_rng(wl.view_seed)``: gain U[0.8,1.2],
    per-channel offset N(0,4), optional white noise +-``noise`` grey levels,
    and a registration error of ``jitter`` rad applied to the *reported*
    rotation (the pixels are rendered with the true one).  ``only`` (a set of
    view indices) skips rendering the other views — they get a zero stub of
    the right shape — while keeping cameras and random draws identical.
    """
    if env is None and (only is None or len(only)):
        env = make_env_map(wl.env_shape, wl.env_seed)
    rng = np.random.default_rng(wl.view_seed)
    k_mat = intrinsics(wl.focal)
    regions = []
    for idx, (yaw, pitch) in enumerate(zip(wl.yaws, wl.pitches)):
        rot = camera_rotation(pitch, yaw)
        gain = rng.uniform(0.8, 1.2)
        offs = rng.normal(0.0, 4.0, size=3)
        jit = rng.normal(0.0, jitter, size=3) if jitter else np.zeros(3)
        rot_reported = rot @ rotation_to_mat(jit) if jitter else rot
        if only is not None and idx not in only:
            if noise:
                rng.uniform(-noise, noise, size=(wl.height, wl.width, 3))
            stub = np.broadcast_to(np.zeros((), np.uint8), (wl.height, wl.width, 3))
            regions.append(Image(stub, rot_reported, k_mat.copy()))
            continue
        img = render_view(env, wl.width, wl.height, wl.focal, rot)
        if photometric or noise:
            pix = img.astype(np.float32)
            if photometric:
                pix = pix * np.float32(gain) + offs.astype(np.float32)
            if noise:
                pix += rng.uniform(-noise, noise, size=pix.shape).astype(np.float32)
            img = np.clip(pix, 0, 255).astype(np.uint8)
        regions.append(Image(np.ascontiguousarray(img), rot_reported, k_mat.copy()))
    return regions


def camera_only(wl: Workload):
    """Views without pixels (geometry studies): img is a zero-size-cost stub
    with the right shape via broadcasting."""
    k_mat = intrinsics(wl.focal)
    stub = np.broadcast_to(np.zeros((), np.uint8), (wl.height, wl.width, 3))
    return [Image(stub, camera_rotation(p, y), k_mat.copy())
            for y, p in zip(wl.yaws, wl.pitches)]
