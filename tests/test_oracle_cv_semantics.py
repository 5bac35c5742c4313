"""Pins oracle/cv_semantics.py (the NumPy restatement of the OpenCV calls at
stitcher.py:315, :226, :56) against the installed cv2 (4.13.0 in this image)."""
import cv2
import numpy as np
import pytest

from oracle import cv_semantics as cs


def test_remap_bit_exact_including_special_coordinates():
    rng = np.random.default_rng(1)
    src = rng.random((97, 131, 4), dtype=np.float32)
    mx = (rng.random((120, 200), dtype=np.float32) * 200 - 30).astype(np.float32)
    my = (rng.random((120, 200), dtype=np.float32) * 150 - 25).astype(np.float32)
    special = [np.nan, np.inf, -np.inf, 1e9, -1e9, 40000, -40000, 2.0 ** 26, -0.0, 130.99]
    mx[0, :len(special)] = special
    my[1, :len(special)] = special
    want = cv2.remap(src, mx, my, cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT)
    assert np.array_equal(cs.remap_linear_reflect(src, mx, my), want)


@pytest.mark.parametrize("level", range(5))
def test_gaussian_kernel_and_ksize(level):
    sigma = np.sqrt(2 * level + 1.0) * 4
    ks = cs.gaussian_ksize(sigma)
    assert ks == [33, 57, 73, 87, 97][level]
    assert np.array_equal(cs.gaussian_kernel(sigma), cv2.getGaussianKernel(ks, sigma, cv2.CV_32F).ravel())


@pytest.mark.parametrize("shape", [(120, 150), (20, 30), (1, 40), (50, 1)])
def test_gaussian_blur_close(shape):
    rng = np.random.default_rng(2)
    img = rng.random(shape + (4,), dtype=np.float32)
    for sigma in (4.0, 12.0):
        want = cv2.GaussianBlur(img, (0, 0), sigma)
        assert np.abs(cs.gaussian_blur(img, sigma) - want).max() < 2e-6


def test_warp_perspective_transparent():
    rng = np.random.default_rng(3)
    src = rng.random((50, 60, 4), dtype=np.float32) + 0.5
    hom = np.array([[0.8, 0.1, 20.3], [-0.07, 0.75, 17.7], [1e-4, -2e-4, 1.0]])
    want = cv2.warpPerspective(src, hom, (120, 90), dst=np.zeros((90, 120, 4), np.float32),
                               borderMode=cv2.BORDER_TRANSPARENT)
    got, written = cs.warp_perspective_transparent(src, hom, 120, 90)
    assert np.array_equal(written, (want != 0).any(-1))
    assert np.abs(got - want).max() < 1e-6
    assert np.array_equal(cv2.invert(hom)[1], cs.invert3x3(hom))


def test_resize_restatement_is_cv2_resize():
    """cv_semantics.resize_u8 (the published resize.cpp arithmetic) against the installed cv2 — the
    pin of the ingest oracle (stitcher.py:418-421)."""
    import cv2
    from oracle import cv_semantics as cs
    rng = np.random.default_rng(1)
    for shape in [(480, 640, 3), (375, 500, 3), (97, 131, 4), (301, 403), (64, 64, 3), (31, 50, 3), (5, 3, 3)]:
        img = rng.integers(0, 256, shape, dtype=np.uint8)
        for shrink in (2, 1.5, 3, 4, 2.5, 1.7, 1.01, 2.0000001, 1.25, 5, 6.3):
            dh, dw = cs.resize_dsize(shape[0], shape[1], 1 / shrink)
            if dh < 1 or dw < 1:
                continue
            want = cv2.resize(img, None, fx=1 / shrink, fy=1 / shrink)
            got = cs.resize_u8(img, 1 / shrink)
            assert got.shape == want.shape and np.array_equal(got, want), (shape, shrink)
