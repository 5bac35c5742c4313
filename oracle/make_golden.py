"""TEST INFRASTRUCTURE — regenerates ``tests/golden/*.npz`` from the LIVE,
unmodified reference (``/root/reference`` via ``oracle/ref_harness.py``).

Run from the repo root in the build container (the reference does not exist
on the GPU box):

    python -B -m oracle.make_golden

Fixtures
--------
tiny4.npz     4 views 160x120 (cfg1 at 1/4 scale, +-20 grey-level noise):
              the input pixels and cameras, the reference mosaic for every
              blender x {gain off,on} x {spherical,cylindrical}, a 6-band
              uncapped multiband mosaic, the patches the reference hands to
              its blender (invalid mask + bbox of every patch, warped RGBA of
              patch 1) for the linear and multiband cases, and the gain-solver inputs/outputs.
cfg1.npz      BASELINE config 1 at full size (4 x 640x480, multiband, 5
              bands): reference mosaic + sha256 of the regenerated inputs.
ring12.npz    12 views 200x150 on a closed 360 degree ring, 2 rows
              (seam-straddling full-width patches), multiband
              5 bands + linear.
"""
from __future__ import annotations

import hashlib
import os
from dataclasses import replace

import numpy as np

from pano360_b200 import synth
from . import ref_harness as rh

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def inputs_digest(regions):
    h = hashlib.sha256()
    for r in regions:
        h.update(np.ascontiguousarray(r.img).tobytes())
        h.update(np.ascontiguousarray(r.rot, dtype=np.float64).tobytes())
        h.update(np.ascontiguousarray(r.intr, dtype=np.float64).tobytes())
    return h.hexdigest()


def pack_inputs(regions):
    return {"imgs": np.stack([r.img for r in regions]),
            "rots": np.stack([r.rot for r in regions]),
            "intrs": np.stack([r.intr for r in regions])}


def tiny4_inputs():
    return synth.make_views(synth.workload("cfg1", scale=4.0), noise=20.0)


def cfg1_inputs():
    return synth.make_views(synth.workload("cfg1"))


def ring12_inputs():
    wl = synth.workload("cfg4", scale=20.0)
    step = np.pi / 3
    yaws = tuple(step * (i - 2.5) for i in range(6)) * 2
    wl = replace(wl, yaws=yaws, pitches=(-0.3,) * 6 + (0.3,) * 6, focal=150.0)
    return synth.make_views(wl, noise=10.0)


def make_tiny4():
    regs = tiny4_inputs()
    out = pack_inputs(regs)
    for blend in ("none", "linear", "multiband"):
        for eq in (False, True):
            for proj in ("spherical", "cylindrical"):
                key = f"mosaic_{blend}_{'eq' if eq else 'raw'}_{proj[:3]}"
                out[key] = rh.ref_stitch(regs, blend, equalize=eq, n_levels=5, proj=proj,
                                         max_resolution=1400)
    out["mosaic_multiband_L6_uncapped"] = rh.ref_stitch(regs, "multiband", n_levels=6,
                                                        max_resolution=1e9)
    out["mosaic_multiband_L1"] = rh.ref_stitch(regs, "multiband", n_levels=1)
    out["mosaic_multiband_L2"] = rh.ref_stitch(regs, "multiband", n_levels=2)
    for blend in ("linear", "multiband"):
        cap = {}
        rh.ref_stitch(regs, blend, capture=cap)
        out[f"patch_shape_{blend}"] = np.array(cap["shape"])
        for i, (warped, mask, (sy, sx)) in enumerate(cap["patches"]):
            if i == 1:      # one full-precision patch per blender keeps the file small
                out[f"patch_{blend}_{i}_warped"] = warped
            out[f"patch_{blend}_{i}_mask"] = mask
            out[f"patch_{blend}_{i}_box"] = np.array([sx.start, sy.start, sx.stop, sy.stop])
    gains = rh.ref_gains(regs)
    out["gain_overlaps"], out["gain_sizes"], out["gains"] = (gains["overlaps"], gains["sizes"],
                                                              gains["gains"])
    np.savez_compressed(os.path.join(OUT, "tiny4.npz"), **out)


def make_cfg1():
    regs = cfg1_inputs()
    wl = synth.workload("cfg1")
    np.savez_compressed(
        os.path.join(OUT, "cfg1.npz"),
        digest=np.array(inputs_digest(regs)),
        mosaic_multiband=rh.ref_stitch(regs, "multiband", n_levels=wl.n_levels,
                                       max_resolution=wl.max_resolution))


def make_ring12():
    regs = ring12_inputs()
    out = pack_inputs(regs)
    out["mosaic_multiband"] = rh.ref_stitch(regs, "multiband", n_levels=5, max_resolution=1e9)
    out["mosaic_linear"] = rh.ref_stitch(regs, "linear", max_resolution=1e9)
    out["mosaic_none"] = rh.ref_stitch(regs, "none", max_resolution=1e9)
    np.savez_compressed(os.path.join(OUT, "ring12.npz"), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    make_tiny4()
    make_cfg1()
    make_ring12()
    for name in sorted(os.listdir(OUT)):
        print(name, os.path.getsize(os.path.join(OUT, name)) // 1024, "KiB")


if __name__ == "__main__":
    main()
