import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/p360_numba_cache")
sys.dont_write_bytecode = True

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs the live reference checkout (/root/reference)")


def pytest_collection_modifyitems(config, items):
    from oracle import ref_harness
    if ref_harness.available():
        return
    skip = pytest.mark.skip(reason="reference checkout not present on this machine")
    for item in items:
        if "reference" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def regions_from_golden(data):
    from pano360_b200.camera import Image
    return [Image(np.ascontiguousarray(img), rot.copy(), intr.copy())
            for img, rot, intr in zip(data["imgs"], data["rots"], data["intrs"])]


def psnr(a, b):
    err = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return np.inf if err == 0 else 10 * np.log10(255.0 ** 2 / err)


def assert_mosaic_close(got, want, max_abs=2, min_psnr=45.0, what=""):
    """north_star tolerance: max |delta| <= 2 on uint8 and PSNR >= 45 dB."""
    assert got.shape == want.shape, f"{what}: shape {got.shape} != {want.shape}"
    assert got.dtype == np.uint8
    diff = np.abs(got.astype(np.int16) - want.astype(np.int16))
    assert diff.max() <= max_abs, f"{what}: max|d|={diff.max()} at {np.argwhere(diff == diff.max())[:4]}"
    assert psnr(got, want) >= min_psnr, f"{what}: psnr={psnr(got, want):.1f}"
    return int(diff.max()), int((diff > 0).sum())
