"""Host-side product code that needs no GPU: geometry, gain solve, crop,
camera record / PKL loading, synthetic generator, C-ABI symbol table."""
import ctypes
import os
import pickle
import re

import numpy as np
import pytest

from oracle import restate as rs
from pano360_b200 import _lib, camera, geometry as geo, synth
from .conftest import ROOT, load_golden, regions_from_golden


def test_plan_matches_oracle_for_every_workload():
    for name, blend in [("cfg1", "multiband"), ("cfg2", "linear"), ("cfg3", "multiband"),
                        ("cfg4", "multiband"), ("cfg5", "none")]:
        wl = synth.workload(name)
        regs = synth.camera_only(wl)
        ours = geo.plan_mosaic(regs, blend == "multiband", wl.max_resolution)
        theirs = rs.plan(regs, blend, wl.max_resolution)
        assert ours.shape == theirs.shape and ours.boxes == theirs.boxes
        assert np.array_equal(ours.resolution, theirs.resolution)
    assert ours.shape == (1216, 5349)


def test_survey_mosaic_sizes():
    want = {"cfg2": (1216, 6720), "cfg3": (6036, 15318), "cfg4": (8807, 31676), "cfg5": (1216, 5349)}
    for name, shape in want.items():
        wl = synth.workload(name)
        assert geo.plan_mosaic(synth.camera_only(wl), False, 1e9).shape == shape


@pytest.mark.parametrize("proj,name", [(geo.SphProj, "spherical"), (geo.CylProj, "cylindrical")])
def test_inverse_map_tables_reproduce_the_reference_map(proj, name):
    data = load_golden("tiny4")
    regs = regions_from_golden(data)
    plan = geo.plan_mosaic(regs, True, 1400, proj)
    opl = rs.plan(regs, "multiband", 1400, name)
    for reg, box in zip(regs, plan.boxes):
        col, row = geo.inverse_map_tables(reg, box, plan, proj)
        p = (col[None, :, :] + row[:, None, :]).astype(np.float32)
        with np.errstate(divide="ignore", invalid="ignore"):
            x = p[..., 0] / p[..., 2] + np.float32(reg.img.shape[1] / 2)
            y = p[..., 1] / p[..., 2] + np.float32(reg.img.shape[0] / 2)
        mx, my, _ = rs.inverse_map(reg, box, opl, name)
        # same float64 products summed in a different association: identical
        # after the float32 cast except on rare rounding ties
        assert np.mean(x != mx) < 1e-3 and np.nanmax(np.abs(x - mx)) < 1e-3
        assert np.mean(y != my) < 1e-3


def test_projection_round_trips():
    rng = np.random.default_rng(0)
    pts = rng.normal(size=(10, 3))
    pts /= np.linalg.norm(pts, axis=1, keepdims=True)
    for proj in (geo.SphProj, geo.CylProj):
        back = proj.proj2hom(proj.hom2proj(pts))
        back /= np.linalg.norm(back, axis=1, keepdims=True)
        np.testing.assert_almost_equal(back, pts)


def test_gaussian_taps_and_lut():
    import cv2
    for lvl in range(5):
        sigma = geo.band_sigma(lvl)
        taps = geo.gaussian_taps(sigma)
        assert np.array_equal(taps, cv2.getGaussianKernel(len(taps), sigma, cv2.CV_32F).ravel())
    img = np.arange(256, dtype=np.uint8).reshape(16, 16, 1).repeat(3, 2)
    ref = rs.rgba_with_weights(img)
    assert np.array_equal(geo.sample_lut()[img[..., 0]], ref[..., 0])
    gained = np.clip(np.float64(1.137) * ref[..., :3], 0, 1).astype(np.float32)   # np.float64 gain: f64 product
    assert np.array_equal(geo.sample_lut(1.137)[img[..., 0]], gained[..., 0])


def test_find_gains_recovers_gains():
    from pano360_b200.geometry import invert3x3
    rng = np.random.default_rng(42)
    size = 10
    gains = 1 + 0.1 * rng.normal(size=size)
    overlaps = 100 + 10 * rng.normal(size=(size, size))
    for i in range(size):
        for j in range(i + 1, size):
            overlaps[i, j] = overlaps[j, i] * gains[j] / gains[i]
    sizes = rng.normal(size=(size, size)) + 10
    # import lazily: stitcher imports torch
    from pano360_b200 import stitcher
    ratio = stitcher.find_gains(overlaps, sizes) / gains
    np.testing.assert_almost_equal(ratio, np.full(size, ratio[0]))
    np.testing.assert_allclose(stitcher.find_gains(overlaps, sizes), rs.solve_gains(overlaps, sizes), rtol=1e-12)
    m = rng.normal(size=(3, 3)) + 3 * np.eye(3)
    np.testing.assert_allclose(invert3x3(m) @ m, np.eye(3), atol=1e-12)


def test_camera_record_and_pickle_roundtrip(tmp_path):
    reg = camera.Image(np.zeros((4, 6, 3), np.uint8), camera.rotation_to_mat([0.1, -0.2, 0.05]),
                       camera.intrinsics(500.0))
    np.testing.assert_almost_equal(reg.hom().dot(reg.proj()), np.eye(3))
    rot = reg.rot
    np.testing.assert_almost_equal(rot.T.dot(rot), np.eye(3))
    # a PKL written by the reference names `bundle_adj.Image`; emulate that
    import sys, types
    fake = types.ModuleType("bundle_adj")
    fake.Image = type("Image", (), {})
    fake.Image.__module__ = "bundle_adj"
    sys.modules["bundle_adj"] = fake
    try:
        obj = fake.Image(); obj.img, obj.rot, obj.intr, obj.range = reg.img, reg.rot, reg.intr, reg.range
        blob = pickle.dumps([obj], protocol=pickle.HIGHEST_PROTOCOL)
    finally:
        del sys.modules["bundle_adj"]
    path = tmp_path / "ba_x_s1.0.pkl"
    path.write_bytes(blob)
    loaded = camera.load_regions(str(path))
    assert isinstance(loaded[0], camera.Image)
    np.testing.assert_array_equal(loaded[0].rot, reg.rot)
    np.testing.assert_almost_equal(camera.hom_to_from(reg, reg), np.eye(3))


def test_synthetic_generator_is_deterministic_and_perturbed():
    wl = synth.workload("cfg1", scale=4.0)
    a, b = synth.make_views(wl), synth.make_views(wl)
    assert all(np.array_equal(x.img, y.img) for x, y in zip(a, b))
    clean = synth.make_views(wl, photometric=False, jitter=0)
    assert any(not np.array_equal(x.img, y.img) for x, y in zip(a, clean))
    assert a[0].img.shape == (120, 160, 3) and a[0].img.dtype == np.uint8


def test_c_abi_exports_every_declared_symbol():
    """The shared library loads and exports exactly what include/*.h declares
    (no compute calls here: there is no GPU)."""
    header = open(os.path.join(ROOT, "include", "pano360_b200.h")).read()
    declared = set(re.findall(r"^\s*(?:int|int64_t)\s+(p360_\w+)\s*\(", header, flags=re.M))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    if not os.path.exists(_lib.LIB_PATH):
        from pano360_b200 import build
        build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert _lib.call("p360_version") == 100
    assert _lib.call("p360_pair_stats_blocks", 480, 640) == 300
    # argument validation happens before any CUDA call and sets the error text
    rc = _lib.load().p360_gauss_blur(None, None, None, 4, 4, None, 3, None)
    assert rc == -22 and "p360_gauss_blur" in _lib.last_error()


def test_job_records_match_the_header(tmp_path):
    """The NumPy mirrors of the job records (filled on the host, shipped as raw
    bytes) have exactly the layout the C header declares: compile a probe
    against include/pano360_b200.h with gcc and compare sizes and offsets."""
    import subprocess
    records = {"p360_warp_job": _lib.WARP_JOB, "p360_blur_job": _lib.BLUR_JOB,
               "p360_band_patch": _lib.BAND_PATCH, "p360_pair_job": _lib.PAIR_JOB,
               "p360_tile_maps": _lib.TILE_MAPS}
    lines = ["#include <stdio.h>", "#include <stddef.h>", '#include "pano360_b200.h"', "int main(void) {"]
    for cname, dtype in records.items():
        lines.append(f'printf("{cname} size %zu\\n", sizeof({cname}));')
        for field in dtype.names:
            lines.append(f'printf("{cname} {field} %zu\\n", offsetof({cname}, {field}));')
    lines += ["return 0; }"]
    src = tmp_path / "probe.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    seen = {}
    for line in filter(None, out):
        cname, field, value = line.split()
        seen[(cname, field)] = int(value)
    for cname, dtype in records.items():
        assert seen[(cname, "size")] == dtype.itemsize, cname
        for field in dtype.names:
            assert seen[(cname, field)] == dtype.fields[field][1], (cname, field)


def test_band_edges_are_translation_invariant():
    from pano360_b200.compositor import band_edges
    for ya, yb, bands in [(0, 1103, 4), (57, 1160, 4), (10, 13, 8), (5, 5, 3)]:
        cuts = band_edges(ya, yb, bands)
        assert cuts[0][0] == ya and cuts[-1][1] == yb and all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
        shifted = band_edges(ya + 1000, yb + 1000, bands)
        assert [(a + 1000, b + 1000) for a, b in cuts] == shifted


def test_active_column_runs_split_only_seam_straddlers():
    wl = synth.workload("cfg4")
    regs = synth.camera_only(wl)
    plan = geo.plan_mosaic(regs, True, 1e9)
    total = 0
    for i, box in enumerate(plan.boxes):
        runs = geo.active_column_runs(i, box, plan, dilate=120)
        wide = box[2] - box[0] > 20000
        assert len(runs) == (2 if wide else 1)
        assert runs[0][0] == box[0] and runs[-1][1] == box[2]
        if wide:
            assert (runs[1][0] - box[0]) % 4 == 0            # coarse-grid phase preserved
        total += sum(b - a for a, b in runs) * (box[3] - box[1])
    assert total < 0.56 * sum((b[2] - b[0]) * (b[3] - b[1]) for b in plan.boxes)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "pano360_b200")
    for fname in os.listdir(pkg):
        if fname.endswith(".py"):
            text = open(os.path.join(pkg, fname)).read()
            assert "oracle" not in re.sub(r'""".*?"""', "", text, flags=re.S), fname


def test_missing_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pano360_b200 import stitcher
    regs = synth.make_views(synth.workload("cfg1", scale=8.0))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        stitcher.stitch(regs, stitcher.multiband_blend)


def test_no_seam_split_when_the_border_crosses_the_gap():
    """tools/fuzz_host.py seed 2900468: 70 small views, cylindrical; view 62 looks so steeply up
    that its footprint winds around the pole.  Its border samples leave a 384-column gap in
    longitude although the footprint covers those columns (330 765 valid pixels): neighbouring
    samples along one side lie on both sides of the gap, so the box must stay whole — while the
    seam-straddling views of the same rig are still split."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("fuzz_host", os.path.join(ROOT, "tools", "fuzz_host.py"))
    fuzz = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fuzz)
    case = fuzz.random_case(np.random.default_rng(2900468))
    regs = case["regs"]
    assert case["cylindrical"] and len(regs) == 70
    plan = geo.plan_mosaic(regs, False, 1e9, geo.CylProj)
    runs = [geo.active_column_runs(i, b, plan, dilate=0) for i, b in enumerate(plan.boxes)]
    assert runs[62] == [(plan.boxes[62][0], plan.boxes[62][2])]
    assert sum(len(r) == 2 for r in runs) >= 5


def test_every_launching_entry_point_is_counted():
    """bench.py's gpu_launches is summed from _lib._LAUNCHES: an entry point that launches kernels
    and is missing there silently undercounts (a trailing comment once swallowed one)."""
    from pano360_b200 import _lib
    no_kernel = {"p360_version", "p360_last_error", "p360_device_info", "p360_pyramid_dims",
                 "p360_pair_stats_blocks", "p360_blur_set_taps", "p360_crop_scratch_bytes",
                 "p360_copy_rect"}                    # (a DMA, not a kernel)
    missing = set(_lib.SIGNATURES) - no_kernel - set(_lib._LAUNCHES)
    assert not missing, missing
    assert all(v >= 1 for v in _lib._LAUNCHES.values())
