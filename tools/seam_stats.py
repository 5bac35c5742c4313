"""How much of a mosaic lies within blur reach of an owner seam?  (CPU only, host geometry.)

Sizes the "seam-band sparsity" step of DESIGN.md §8: the coarse Gaussian levels are only
consumed where two owner regions meet within the widest level's support, so the share of
64x32 collapse tiles touching that band bounds what reduce/blur need to compute.

    python tools/seam_stats.py [cfg4] [--step 8]
"""
import argparse
import os
import sys

import numpy as np
from scipy.ndimage import binary_dilation

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from pano360_b200 import geometry as geo, synth  # noqa: E402


def owner_on_grid(regs, plan, step):
    """Owner index on a `step`-pixel grid: argmax of the hat-product weight, first max wins."""
    H, W = plan.shape
    ys, xs = np.arange(0, H, step), np.arange(0, W, step)
    rx, rz, ry = plan.rays()
    best = np.zeros((len(ys), len(xs)), np.float32)
    own = np.full(best.shape, -1, np.int32)
    for i, (reg, box) in enumerate(zip(regs, plan.boxes)):
        x0, y0, x1, y1 = box
        sx, sy = xs[(xs >= x0) & (xs < x1)], ys[(ys >= y0) & (ys < y1)]
        kr = reg.proj()
        h, w = reg.img.shape[:2] if reg.img is not None else reg.shape[:2]
        p = (kr[:, 0][None, None, :] * rx[sx][None, :, None] + kr[:, 1][None, None, :] * ry[sy][:, None, None]
             + kr[:, 2][None, None, :] * rz[sx][None, :, None])
        z = p[..., 2]
        with np.errstate(divide="ignore", invalid="ignore"):
            u, v = p[..., 0] / z + w / 2, p[..., 1] / z + h / 2
        ok = (z > 0) & (u >= 0) & (u <= w - 1) & (v >= 0) & (v <= h - 1)
        a = np.where(ok, (0.5 - np.abs(u / w - 0.5)) * (0.5 - np.abs(v / h - 0.5)), 0).astype(np.float32)
        iy, ix = (sy // step)[:, None], (sx // step)[None, :]
        cur = best[iy, ix]
        better = a > cur
        best[iy, ix] = np.where(better, a, cur)
        own[iy, ix] = np.where(better, i, own[iy, ix])
    return own


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload", nargs="?", default="cfg4")
    ap.add_argument("--step", type=int, default=8)
    args = ap.parse_args()
    wl = synth.workload(args.workload)
    regs = synth.make_views(wl, only=set())
    plan = geo.plan_mosaic(regs, True, 1e9)
    own = owner_on_grid(regs, plan, args.step)
    seam = np.zeros_like(own, bool)
    seam[:, 1:] |= own[:, 1:] != own[:, :-1]
    seam[1:, :] |= own[1:, :] != own[:-1, :]
    print(f"{args.workload}: mosaic {plan.shape[0]}x{plan.shape[1]}, owned {100 * (own >= 0).mean():.1f}%")
    th, tw = max(32 // args.step, 1), max(64 // args.step, 1)
    for reach in (56, 64, 112, 120):
        r = int(np.ceil(reach / args.step))
        band = binary_dilation(seam, structure=np.ones((2 * r + 1, 2 * r + 1), bool))
        hh, ww = band.shape[0] // th * th, band.shape[1] // tw * tw
        tiles = band[:hh, :ww].reshape(hh // th, th, ww // tw, tw).any(axis=(1, 3))
        print(f"  reach {reach:3d} px: pixels within reach {100 * band.mean():5.1f}%, 64x32 tiles touching {100 * tiles.mean():5.1f}%")


if __name__ == "__main__":
    main()
