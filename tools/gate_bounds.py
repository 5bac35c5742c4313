"""Study for gating the warp: which patches can own a pixel of a 64 x 32
mosaic tile, decided from the geometry alone, BEFORE anything is sampled?

Ownership is arg-max of alpha = hat_y(v) * hat_x(u) over the patches, (u, v) the source position
of the mosaic pixel.  Interval arithmetic over a tile — ray tables -> K R ray -> u, v -> alpha —
gives alpha_min / alpha_max per (patch, tile); a patch whose alpha_max lies below another patch's
alpha_min anywhere on the tile can never win there.  CAND(tile) = everything else.  The warp would
only have to run where a patch is in CAND within the blur chain's reach (at cfg4 60 % of its
blocks are ever read, tools/seam_map_stats.py).

This tool checks on the HOST build of the kernels that CAND is conservative (every true owner of
every tile is in it) on random rigs and reports how tight it is.  Development tooling only.

    python tools/gate_bounds.py --cases 100
"""
import argparse
import os
import sys

import numpy as np
import pytest

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from pano360_b200 import _lib, geometry as geo, synth  # noqa: E402
from tests.emul import harness  # noqa: E402

TX, TY = 64, 32


def _tile_interval(values, size):
    """Per tile of `size` entries: (min, max) of a 1-D table."""
    n = -(-len(values) // size)
    padded = np.concatenate([values, np.full(n * size - len(values), values[-1])]).reshape(n, size)
    return padded.min(axis=1), padded.max(axis=1)


def _imul(k, lo, hi):
    a, b = k * lo, k * hi
    return np.minimum(a, b), np.maximum(a, b)


def _hat_bounds(lo, hi, size):
    """Bounds of the interpolated hat table over source coordinates [lo, hi] clipped to the
    valid range [0, size - 1]; (min, max, any_valid, all_valid)."""
    any_valid = (hi >= 0) & (lo <= size - 1)
    all_valid = (lo >= 0) & (hi <= size - 1)
    a, b = np.clip(lo, 0, size - 1), np.clip(hi, 0, size - 1)
    hat = lambda t: 0.5 - np.abs(t - size / 2.0) / size              # >= the interpolated table
    lower = lambda t: np.minimum(0.5 - np.abs(np.floor(t) - size / 2.0) / size,
                                 0.5 - np.abs(np.ceil(t) - size / 2.0) / size)   # <= the interpolant (concave)
    peak_inside = (a <= size / 2.0) & (b >= size / 2.0)
    upper = np.where(peak_inside, 0.5, np.maximum(hat(a), hat(b)))
    return np.minimum(lower(a), lower(b)), upper, any_valid, all_valid


def alpha_bounds(reg, box, plan, proj=geo.SphProj):
    """(alpha_min, alpha_max) of one patch per mosaic tile (0, 0 where it has no valid pixel)."""
    height, width = plan.shape
    ray_x, ray_z, ray_y = plan.rays(proj)
    ray_x, ray_z, ray_y = ray_x[:width], ray_z[:width], ray_y[:height]
    (x_lo, x_hi), (z_lo, z_hi) = _tile_interval(ray_x, TX), _tile_interval(ray_z, TX)
    y_lo, y_hi = _tile_interval(ray_y, TY)
    kr = np.asarray(reg.proj(), dtype=np.float64)
    h, w = reg.img.shape[:2]

    def comp(row):
        ax = _imul(kr[row, 0], x_lo, x_hi)
        az = _imul(kr[row, 2], z_lo, z_hi)
        ay = _imul(kr[row, 1], y_lo, y_hi)
        return (ax[0] + az[0])[None, :] + ay[0][:, None], (ax[1] + az[1])[None, :] + ay[1][:, None]

    (px_lo, px_hi), (py_lo, py_hi), (pz_lo, pz_hi) = comp(0), comp(1), comp(2)
    front = pz_lo > 1e-9                                  # the whole tile in front of the camera
    safe_lo = np.where(front, pz_lo, 1.0)
    safe_hi = np.where(front, pz_hi, 1.0)

    def quotient(lo, hi):
        c = np.stack([lo / safe_lo, lo / safe_hi, hi / safe_lo, hi / safe_hi])
        return c.min(axis=0), c.max(axis=0)

    slack = 1.0 / 32 + 1e-3                               # 1/32-px fixed point + float32 rounding of the quotient
    u_lo, u_hi = quotient(px_lo, px_hi)
    v_lo, v_hi = quotient(py_lo, py_hi)
    hx = _hat_bounds(u_lo + w / 2.0 - slack, u_hi + w / 2.0 + slack, w)
    hy = _hat_bounds(v_lo + h / 2.0 - slack, v_hi + h / 2.0 + slack, h)
    any_valid = np.where(front, hx[2] & hy[2], pz_hi > 0)                 # partly behind the camera: anything goes
    all_valid = front & hx[3] & hy[3]
    a_max = np.where(front, hx[1] * hy[1], 0.25) * (1 + 1e-5)
    a_min = np.where(all_valid, hx[0] * hy[0], 0.0) * (1 - 1e-5)
    x0, y0, x1, y1 = box                                                   # nothing outside the patch box
    ty, tx = a_max.shape
    inside = np.zeros((ty, tx), bool)
    inside[max(y0, 0) // TY:-(-min(y1, height) // TY), max(x0, 0) // TX:-(-min(x1, width) // TX)] = True
    any_valid &= inside
    # a patch only dominates a tile it covers completely: beyond its box it has no pixels
    txa, tya = np.arange(tx) * TX, np.arange(ty) * TY
    cols = (x0 <= txa) & (x1 >= np.minimum(txa + TX, width))
    rows = (y0 <= tya) & (y1 >= np.minimum(tya + TY, height))
    full = rows[:, None] & cols[None, :]
    return np.where(any_valid & all_valid & full, a_min, 0.0), np.where(any_valid, a_max, 0.0), any_valid


def candidates(regs, plan, proj=geo.SphProj, crops=None):
    """bool [n, tiles_y, tiles_x]: patch may own a pixel of the tile.  ``crops`` = [(image, x0, y0,
    x1, y1), ...] replaces the boxes of the plan (seam-split pieces, row windows)."""
    if crops is None:
        crops = [(i,) + tuple(b) for i, b in enumerate(plan.boxes)]
    bounds = [alpha_bounds(regs[c[0]], tuple(c[1:5]), plan, proj) for c in crops]
    best_min = np.max([b[0] for b in bounds], axis=0)
    return np.stack([b[2] & (b[1] >= best_min) for b in bounds])


def exact_owners(comp, regs, plan, proj):
    """bool [n, tiles_y, tiles_x] from the owner keys the kernels produce (host build)."""
    src = comp.upload(regs)
    state = comp.new_owner_state(plan.shape)
    crops, tables = comp.plan_crops(regs, plan, proj)
    comp.warp_crops(src, crops, tables, owner_state=state)
    owner = comp.owner_map(None, plan.shape, owner_state=state)[0].numpy()
    n = len(regs)
    ty, tx = -(-plan.shape[0] // TY), -(-plan.shape[1] // TX)
    present = np.zeros((n, ty, tx), bool)
    yy, xx = np.nonzero(owner >= 0)
    image_of = np.array([c[0] for c in crops])
    present[image_of[owner[yy, xx]], yy // TY, xx // TX] = True
    return present


def main():
    import fuzz_host
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=60)
    ap.add_argument("--seed", type=int, default=3)
    args = ap.parse_args()
    patcher = pytest.MonkeyPatch()
    comp = harness.install(patcher)
    misses = tight_c = tight_p = 0
    try:
        cases = [("cfg4/8", synth.make_views(synth.workload("cfg4", scale=8.0), noise=5.0), False),
                 ("cfg3/8", synth.make_views(synth.workload("cfg3", scale=8.0), noise=5.0), False)]
        for k in range(args.cases):
            case = fuzz_host.random_case(np.random.default_rng(args.seed * 7919 + k))
            cases.append((f"random {k} ({case['layout']}, {len(case['regs'])} views)", case["regs"], case["cylindrical"]))
        for name, regs, cyl in cases:
            proj = geo.CylProj if cyl else geo.SphProj
            plan = geo.plan_mosaic(regs, True, 1e9, proj)
            cand = candidates(regs, plan, proj)
            present = exact_owners(comp, regs, plan, proj)
            missed = present & ~cand
            if missed.any():
                misses += 1
                print("NOT CONSERVATIVE:", name, "patch/tile", np.argwhere(missed)[:5].tolist())
            tight_c += int(cand.sum())
            tight_p += int(present.sum())
            if not name.startswith("random"):
                print(f"{name}: {int(present.sum())} (patch, tile) pairs own a pixel, {int(cand.sum())} are candidates, "
                      f"{len(regs)} patches x {cand.shape[1] * cand.shape[2]} tiles")
    finally:
        patcher.undo()
    print(f"{len(cases)} rigs, {misses} not conservative; candidates / true owners = {tight_c / max(tight_p, 1):.2f}")
    return 1 if misses else 0


if __name__ == "__main__":
    sys.exit(main())
