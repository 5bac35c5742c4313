"""Differential pin of the oracle against the LIVE reference (only where
/root/reference exists; skipped on the GPU box)."""
import numpy as np
import pytest

from oracle import ref_harness as rh, restate as rs
from pano360_b200 import synth

pytestmark = pytest.mark.reference


@pytest.fixture(scope="module")
def views():
    return synth.make_views(synth.workload("cfg1", scale=2.0), noise=20.0)


@pytest.mark.parametrize("blend", ["none", "linear", "multiband"])
@pytest.mark.parametrize("eq", [False, True])
def test_bit_identical(views, blend, eq):
    ref = rh.ref_stitch(views, blend, equalize=eq, n_levels=5, max_resolution=1400)
    assert np.array_equal(rs.stitch(views, blend, eq, 5, 1400), ref)


def test_cylindrical_six_bands_uncapped(views):
    ref = rh.ref_stitch(views, "multiband", n_levels=6, proj="cylindrical", max_resolution=1e9)
    assert np.array_equal(rs.stitch(views, "multiband", False, 6, 1e9, "cylindrical"), ref)


def test_two_row_layout():
    wl = synth.workload("cfg3", scale=16.0)
    regs = synth.make_views(wl, noise=10.0)
    ref = rh.ref_stitch(regs, "multiband", n_levels=6, max_resolution=1e9)
    assert np.array_equal(rs.stitch(regs, "multiband", False, 6, 1e9), ref)


@pytest.mark.parametrize("blend", ["none", "multiband"])
def test_crop_is_the_references(views, blend):
    """-c: the oracle's crop (valid mask + largest-rectangle scan) gives the reference's cropped mosaic."""
    ref = rh.ref_stitch(views, blend, n_levels=5, max_resolution=1400, crop=True)
    got = rs.stitch(views, blend, False, 5, 1400, crop=True)
    assert got.shape == ref.shape and np.array_equal(got, ref)


def test_crop_scan_against_the_references_on_random_masks():
    st, _ = rh.load()
    rng = np.random.default_rng(12)
    for trial in range(8):
        h, w = int(rng.integers(3, 70)), int(rng.integers(3, 1200))
        valid = rng.random((h, w)) > 0.02 * (trial + 1)
        valid[:, 0] |= trial % 2 == 0                     # exercise the column-0 quirk
        mosaic = rng.integers(0, 255, (h, w, 3), dtype=np.uint8)
        want = st.crop_mosaic(mosaic, valid)
        y0, y1, x0, x1 = rs.crop_rect(valid)
        assert np.array_equal(mosaic[y0:y1, x0:x1], want), trial


def test_reference_unit_tests_still_pass():
    """The reference's own 8 unit tests (pano_tests.py) through the harness:
    regression for the untouched host code we lean on."""
    import importlib.util
    import os
    import sys
    import unittest
    rh.load()
    ref = rh.reference_dir()
    sys.path.insert(0, ref)
    try:
        spec = importlib.util.spec_from_file_location("pano_tests", os.path.join(ref, "pano_tests.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        np.random.seed(42)
        result = unittest.TextTestRunner(verbosity=0).run(
            unittest.defaultTestLoader.loadTestsFromModule(mod))
    finally:
        sys.path.remove(ref)
    assert result.testsRun == 8 and result.wasSuccessful()
