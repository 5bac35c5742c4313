"""Randomised parity campaign on the HOST build of the kernels (tests/emul): random rigs
(view count, image size incl. odd and tiny, focal length, yaw / pitch incl. rings that straddle
the +-pi seam and steep pitches, roll), blenders, band counts, -e, projection, resolution cap,
forced seam-band maps, seam plan / direct tiles, random row windows — each compared with the CPU oracle
(none / linear bit-exact, multiband and -e within max|d| <= 2 and PSNR >= 45 dB) and, for windows,
with the whole mosaic byte for byte.

Development tooling: needs no GPU, runs until --cases or --seconds are used up, prints every
failing case with the seed that reproduces it.

    python tools/fuzz_host.py --cases 200 --seed 1
"""
import argparse
import os
import sys
import time
import traceback
from dataclasses import replace

import numpy as np
import pytest

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from oracle import restate as rs  # noqa: E402
from pano360_b200 import geometry as geo, synth  # noqa: E402
from pano360_b200.camera import Image, rotation_to_mat  # noqa: E402
from tests.conftest import psnr  # noqa: E402
from tests.emul import harness  # noqa: E402


def random_case(rng):
    n = int(rng.choice([1, 2, 3, 4, 5, 6, 8, 12, 12, 40, 70]))           # 40 / 70: multi-word tile bitmaps
    # (with views this small ownership can degenerate into slivers a pixel or two wide, which the
    # coarse evaluation of the blurs resolves poorly: DESIGN.md §2, known limitation — run_case
    # accepts deviations > 2 only on such slivers)
    width, height = int(rng.integers(24, 260)), int(rng.integers(24, 200))
    if n >= 40:
        width, height = int(rng.integers(24, 90)), int(rng.integers(24, 70))
    focal = float(rng.uniform(0.6, 2.5) * max(width, height))
    layout = rng.choice(["ring", "arc", "grid", "scatter"])
    if layout == "ring":                      # full circle: boxes straddle the +-pi seam
        yaws = np.linspace(-np.pi, np.pi, n, endpoint=False) + rng.uniform(-0.3, 0.3)
        pitches = rng.uniform(-0.2, 0.2, n)
    elif layout == "arc":
        step = rng.uniform(0.3, 0.9) * width / focal
        yaws = step * (np.arange(n) - (n - 1) / 2) + rng.uniform(-3.0, 3.0)
        pitches = rng.uniform(-0.1, 0.1, n)
    elif layout == "grid":
        cols = max(1, n // 2)
        step = rng.uniform(0.4, 0.8) * width / focal
        yaws = np.array([step * (i % cols) for i in range(n)]) + rng.uniform(-3.0, 3.0)
        pitches = np.array([(i // cols - 0.5) * rng.uniform(0.3, 0.7) * height / focal for i in range(n)])
    else:
        yaws, pitches = rng.uniform(-3.1, 3.1, n), rng.uniform(-1.0, 1.0, n)
    wl = synth.Workload("fuzz", width, height, focal, tuple(float(v) for v in yaws), tuple(float(v) for v in pitches),
                        "multiband", 5, False, (256, 512), 1e9, int(rng.integers(1 << 30)), int(rng.integers(1 << 30)))
    regs = synth.make_views(wl, noise=float(rng.choice([0.0, 5.0, 40.0])))
    mixed = n > 1 and rng.random() < 0.25
    if mixed:                                 # every other view at another size (and focal length)
        w2, h2 = int(rng.integers(24, 200)), int(rng.integers(24, 160))
        other = synth.make_views(replace(wl, width=w2, height=h2, focal=wl.focal * w2 / width), noise=5.0)
        regs = [other[i] if i % 2 else regs[i] for i in range(n)]
    if rng.random() < 0.5:                    # roll + shuffled list order
        regs = [Image(r.img, rotation_to_mat([0.0, 0.0, float(rng.uniform(-0.5, 0.5))]) @ r.rot, r.intr) for r in regs]
        regs = [regs[i] for i in rng.permutation(n)]
    return dict(regs=regs, blend=str(rng.choice(["none", "linear", "multiband", "multiband"])),
                equalize=bool(rng.random() < 0.3) and not mixed, levels=int(rng.choice([1, 2, 3, 5, 5, 6, 8])),
                cylindrical=bool(rng.random() < 0.25), cap=float(rng.choice([1e9, 1e9, 1400, 300])),
                maps=[None, True, False][int(rng.integers(3))], direct=bool(rng.random() < 0.7), layout=str(layout))


def only_on_slivers(regs, case, levels, bad):
    """Do all offending pixels lie on owner regions at most 3 pixels wide (in x or in y)?"""
    stages = {}
    patches, pl = rs.build_patches(regs, "multiband", case["equalize"], case["cap"],
                                   "cylindrical" if case["cylindrical"] else "spherical")
    rs.multiband(patches, pl.shape, levels, stages=stages)
    own = stages["owner"]
    for y, x in np.argwhere(bad):
        row, col = own[y], own[:, x]
        width = 1 + sum(1 for d in (-1, 1) for k in range(1, 4) if 0 <= x + d * k < len(row) and np.all(row[min(x, x + d * k):max(x, x + d * k) + 1] == own[y, x]))
        height = 1 + sum(1 for d in (-1, 1) for k in range(1, 4) if 0 <= y + d * k < len(col) and np.all(col[min(y, y + d * k):max(y, y + d * k) + 1] == own[y, x]))
        if min(width, height) > 3:
            return False
    return True


def run_case(st, comp, case):
    regs, blend, levels = case["regs"], case["blend"], case["levels"]
    st.MAX_RESOLUTION = case["cap"]
    st.SphProj = geo.CylProj if case["cylindrical"] else geo.SphProj
    comp.seam_maps = case["maps"]
    comp.direct = case["direct"]
    proj = st.SphProj
    try:
        want = rs.stitch(regs, blend, case["equalize"], levels, case["cap"],
                         proj="cylindrical" if case["cylindrical"] else "spherical")
    except Exception as exc:       # e.g. singular gain system for views without overlap: same error expected
        try:
            st.stitch(regs, blender=st.BLENDERS[blend], equalize=case["equalize"], n_levels=levels)
        except type(exc):
            return None
        raise AssertionError(f"the oracle raised {exc!r}, the kernels' path did not")
    got = st.stitch(regs, blender=st.BLENDERS[blend], equalize=case["equalize"], n_levels=levels)
    assert got.shape == want.shape, (got.shape, want.shape)
    diff = np.abs(got.astype(np.int16) - want.astype(np.int16))
    if blend == "multiband" or case["equalize"]:
        # -e: the gains come from float64 sums in another order than NumPy's float32 pairwise means
        # (rtol ~1e-7), which can move a LUT entry by an ulp and a truncated uint8 by one level
        # (views below 128 px are blended at full resolution — Compositor.needs_exact — since owner
        # regions a pixel or two wide, which only such views produce, are beyond the coarse grids)
        assert diff.max() <= 2 and psnr(got, want) >= 45.0, (blend, int(diff.max()), psnr(got, want))
    else:
        assert diff.max() == 0, (blend, int(diff.max()), int((diff > 0).sum()))
    return got


def run_blender_api(st, case):
    """The drop-in blender functions on the oracle's own NumPy patches (external patches take the
    owner-update kernel instead of the fused competition of the warp)."""
    regs, blend = case["regs"], case["blend"]
    proj = "cylindrical" if case["cylindrical"] else "spherical"
    patches, pl = rs.build_patches(regs, blend, False, case["cap"], proj)
    want = rs.BLENDERS[blend]([(w.copy(), m.copy(), s) for w, m, s in patches], pl.shape)
    got = st.BLENDERS[blend](patches, pl.shape)
    diff = np.abs(got.astype(np.int16) - want.astype(np.int16))
    if blend == "multiband":
        assert diff.max() <= 2 and psnr(got, want) >= 45.0, ("blender api", int(diff.max()), psnr(got, want))
    else:
        assert diff.max() == 0, ("blender api", blend, int(diff.max()))


def run_windows(comp, case, whole, rng):
    """Random windows of the same composite against the whole mosaic (no gains: the window API
    takes the sources as uploaded): row windows, rows x tile-aligned columns, the same written in
    place into a whole-mosaic image, and — where the seam plan applies — the composite from sources
    scrambled outside the plan's source rectangles."""
    if case["equalize"]:
        return
    import torch
    regs, blend, levels = case["regs"], case["blend"], case["levels"]
    proj = geo.CylProj if case["cylindrical"] else geo.SphProj
    plan = geo.plan_mosaic(regs, blend == "multiband", case["cap"], proj)
    src = comp.upload(regs)
    h, w = plan.shape
    exact = comp.needs_exact(regs)
    tiles = -(-w // 64)
    for k in range(3):
        if h < 2:
            break
        ya, yb = sorted(int(v) for v in rng.choice(h + 1, 2, replace=False))
        cols = None
        if k and tiles > 1:
            ta, tb = sorted(int(v) for v in rng.choice(tiles + 1, 2, replace=False))
            cols = (64 * ta, min(64 * tb, w))
        rows = (ya, yb) if k < 2 else None
        y0, y1 = rows or (0, h)
        x0, x1 = cols or (0, w)
        strip = comp.composite(regs, src, plan, blend, levels, proj, rows=rows, cols=cols,
                               exact=exact)[0].numpy()           # (the path stitch() took)
        assert np.array_equal(strip, whole[y0:y1, x0:x1]), ("window", rows, cols)
        if not exact:
            target = torch.full((h, w, 3), 9, dtype=torch.uint8)
            comp.composite(regs, src, plan, blend, levels, proj, rows=rows, cols=cols, out_dev=(target.data_ptr(), w, target))
            got = target.numpy()
            assert np.array_equal(got[y0:y1, x0:x1], whole[y0:y1, x0:x1]), ("in place", rows, cols)
            got[y0:y1, x0:x1] = 9
            assert (got == 9).all(), ("in place: outside touched", rows, cols)
    rects = comp.source_rects(regs, plan, blend, levels, proj)
    if rects is not None and comp.direct:
        noisy = []
        for i, reg in enumerate(regs):
            img = rng.integers(0, 256, reg.img.shape, dtype=np.uint8)
            if i in rects:
                r0, r1, c0, c1 = rects[i]
                img[r0:r1, c0:c1] = reg.img[r0:r1, c0:c1]
            noisy.append(Image(img, reg.rot, reg.intr))
        again = comp.composite(noisy, comp.upload(noisy), plan, blend, levels, proj)[0].numpy()
        assert np.array_equal(again, whole), "source rectangles"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=100)
    ap.add_argument("--seconds", type=float, default=1e9)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    patcher = pytest.MonkeyPatch()
    comp = harness.install(patcher)
    from pano360_b200 import stitcher as st
    st._compositors[0] = comp
    failures, t0, done = [], time.time(), 0
    try:
        for k in range(args.cases):
            if time.time() - t0 > args.seconds:
                break
            seed = args.seed * 100003 + k
            rng = np.random.default_rng(seed)
            case = random_case(rng)
            tag = (f"seed {seed}: {len(case['regs'])} x {case['regs'][0].img.shape[1]}x{case['regs'][0].img.shape[0]} "
                   f"{case['layout']} {case['blend']} L{case['levels']} eq={case['equalize']} cyl={case['cylindrical']} "
                   f"cap={case['cap']:g} maps={case['maps']} direct={case['direct']}")
            try:
                whole = run_case(st, comp, case)
                if whole is not None:
                    run_windows(comp, case, whole, rng)
                if not case["equalize"] and rng.random() < 0.3:
                    run_blender_api(st, case)
            except Exception as exc:                              # keep going: collect every failure
                failures.append((tag, exc))
                print("FAIL", tag, "->", repr(exc)[:300], flush=True)
                if not isinstance(exc, AssertionError):
                    traceback.print_exc()
            done += 1
            if done % 25 == 0:
                print(f"... {done} cases, {len(failures)} failures, {time.time() - t0:.0f} s", flush=True)
    finally:
        patcher.undo()
    print(f"{done} cases, {len(failures)} failures, {time.time() - t0:.0f} s")
    return 1 if failures else 0


if __name__ == "__main__":
    sys.exit(main())
