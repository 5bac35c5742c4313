"""How much of reduce / blur do the seam-band maps skip, compared with the owned boxes?

Runs warp + p360_tile_maps_build of a workload at reduced scale on the HOST build of the
kernels (tests/emul — development tooling, no GPU needed) and evaluates both skip predicates
for every 32 x 32 reduce block.  Also reports the share of collapse tiles that stay on the
single-owner shortcut.

    python tools/seam_map_stats.py cfg4 --scale 4
"""
import argparse
import os
import sys

import numpy as np
import pytest

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from pano360_b200 import _lib, geometry as geo, synth  # noqa: E402
from tests.emul import harness  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload", nargs="?", default="cfg4")
    ap.add_argument("--scale", type=float, default=4.0)
    args = ap.parse_args()
    patcher = pytest.MonkeyPatch()
    comp = harness.install(patcher)
    try:
        wl = synth.workload(args.workload, scale=args.scale)
        regs = synth.make_views(wl, noise=5.0)
        plan = geo.plan_mosaic(regs, True, 1e9)
        src = comp.upload(regs)
        comp.seam_maps = True
        comp.composite(regs, src, plan, "multiband", wl.n_levels)
        table = np.frombuffer(comp._keep["collapse"][0].numpy().tobytes(), dtype=_lib.BAND_PATCH)
        maps, (bits, multi) = comp._keep["bands"][3], comp._keep["bands"][4]
        tiles_x, tiles_y, words = int(maps["tiles_x"][0]), int(maps["tiles_y"][0]), int(maps["words"][0])
        row0 = int(maps["row0"][0])
        cells = tiles_x * tiles_y
        cap = int(maps["work_cap"][0])
        flat = bits.numpy().view(np.uint32)
        print(f"last work list (vertical blur): {int(flat[0])} blocks of a {cap}-block scratch")
        planes = flat[2 + 2 * cap:].reshape(3, tiles_y, tiles_x, words)
        cand, need = planes[1], planes[2]
        multi = multi.numpy().reshape(tiles_y, tiles_x)
        n_cand = sum(np.bitwise_count(cand[..., w]).astype(np.int64) for w in range(words))
        print(f"{args.workload} / {args.scale:g}: mosaic {plan.shape[0]}x{plan.shape[1]}, {len(table)} patches, "
              f"{cells} tiles: {100 * np.mean(n_cand == 0):.1f}% empty, {100 * np.mean((n_cand >= 1) & (multi == 0)):.1f}% "
              f"single owner (shortcut), {100 * np.mean(multi != 0):.1f}% blended")
        def blocks_hit(has, xs, ys, bw, bh):
            """grid of blocks at window px (xs, ys), bw x bh px each: which touch a tile with the bit set?"""
            cs = np.zeros((tiles_y + 1, tiles_x + 1), np.int64)
            cs[1:, 1:] = has.astype(np.int64).cumsum(0).cumsum(1)
            xa, xb = np.clip(xs >> 6, 0, tiles_x - 1), np.clip((xs + bw - 1) >> 6, 0, tiles_x - 1) + 1
            ya, yb = np.clip((ys - row0) >> 5, 0, tiles_y - 1), np.clip((ys + bh - 1 - row0) >> 5, 0, tiles_y - 1) + 1
            return (cs[yb][:, xb] - cs[ya][:, xb] - cs[yb][:, xa] + cs[ya][:, xa]) > 0

        # the kernels' block shapes (coarse cells) first, then alternatives worth measuring
        shapes = [("H", 256, 4), ("V", 32, 64), ("H", 64, 4), ("H", 64, 16), ("H", 128, 8), ("V", 32, 32), ("V", 16, 64)]
        blur = {}
        for k, rec in enumerate(table):
            pad, w4, h4 = int(rec["pad"]), int(rec["w4"]), int(rec["h4"])
            has = (need[..., k >> 5] >> np.uint32(k & 31)) & 1
            for f, cw, ch in ((2, 2 * w4, 2 * h4), (4, w4, h4)):
                for name, bw, bh in shapes:
                    xs = np.arange(-(-cw // bw)) * bw * f - pad + int(rec["x0"])
                    ys = np.arange(-(-ch // bh)) * bh * f - pad + int(rec["y0"])
                    hit = blocks_hit(has, xs, ys, bw * f, bh * f)
                    acc = blur.setdefault((name, bw, bh, f), [0, 0])
                    acc[0] += int(hit.sum()) * bw * bh
                    acc[1] += cw * ch
        for (name, bw, bh, f), (run, tot) in blur.items():
            print(f"blur {name} blocks of {bw}x{bh} cells, f={f}: {run / 1e6:.1f} of {tot / 1e6:.1f} Mcells "
                  f"({100 * run / max(tot, 1):.1f}%)")
        # upper bound for gating the warp itself: blocks (64 x 16 px) of a patch whose pixels anybody
        # reads = tiles where it owns a pixel or is reduced / blended, grown by one tile for the
        # overhang of reduce blocks and reflections
        from scipy.ndimage import maximum_filter
        present = planes[0]
        k1_run = k1_tot = 0
        for k, rec in enumerate(table):
            bit = np.uint32(k & 31)
            used = (((present[..., k >> 5] | need[..., k >> 5]) >> bit) & 1).astype(bool)
            used = maximum_filter(used, size=(3, 3), mode="constant")
            xs = np.arange(-(-int(rec["pw"]) // 64)) * 64 + int(rec["x0"])
            ys = np.arange(-(-int(rec["ph"]) // 16)) * 16 + int(rec["y0"])
            hit = blocks_hit(used, xs, ys, 64, 16)
            k1_run += int(hit.sum())
            k1_tot += hit.size
        print(f"warp blocks (64x16 px) whose output is ever read: {k1_run} of {k1_tot} ({100 * k1_run / k1_tot:.1f}%)")
        run_own = run_maps = total = 0
        for k, rec in enumerate(table):
            pad, w4, h4 = int(rec["pad"]), int(rec["w4"]), int(rec["h4"])
            bx = np.arange(-(-4 * w4 // 32)) * 32 - pad
            by = np.arange(-(-h4 // 8)) * 32 - pad
            ox0, oy0, ox1, oy1 = (int(v) for v in rec["own"])
            grow = 2 * pad + 4
            if ox1 > ox0 and oy1 > oy0:
                near_x = (bx < ox1 + grow) & (bx + 32 > ox0 - grow)
                near_y = (by < oy1 + grow) & (by + 32 > oy0 - grow)
                run_own += int(near_x.sum()) * int(near_y.sum())
            has = (need[..., k >> 5] >> np.uint32(k & 31)) & 1
            tx0 = np.clip((bx + int(rec["x0"])) >> 6, 0, tiles_x - 1)
            tx1 = np.clip((bx + int(rec["x0"]) + 31) >> 6, 0, tiles_x - 1)
            ty0 = np.clip((by + int(rec["y0"]) - row0) >> 5, 0, tiles_y - 1)
            ty1 = np.clip((by + int(rec["y0"]) + 31 - row0) >> 5, 0, tiles_y - 1)
            hit = has[ty0][:, tx0] | has[ty0][:, tx1] | has[ty1][:, tx0] | has[ty1][:, tx1]
            run_maps += int(hit.sum())
            total += len(bx) * len(by)
        print(f"reduce blocks (32x32 px): {total} in the grids; owned boxes run {run_own} ({100 * run_own / total:.1f}%), "
              f"seam-band maps run {run_maps} ({100 * run_maps / total:.1f}%)  ->  {run_own / max(run_maps, 1):.2f}x fewer")
    finally:
        patcher.undo()


if __name__ == "__main__":
    main()
