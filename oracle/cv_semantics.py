"""TEST INFRASTRUCTURE — NumPy restatement of the OpenCV primitives the
reference's hot path delegates to.

The reference's per-pixel arithmetic lives in a third-party dependency that
is *not* vendored and *not* pinned (Readme.md:21-25 ``conda install opencv``);
the de-facto oracle version is the one installed in this image,
**opencv-python-headless 4.13.0** (SURVEY.md §8c).  The call sites are
stitcher.py:315-316 (``cv2.remap``), :226 (``cv2.GaussianBlur``), :56-57
(``cv2.warpPerspective``).  The semantics restated here follow OpenCV's
published algorithm (modules/imgproc/src/imgwarp.cpp ``remapBilinear`` /
``WarpPerspectiveInvoker``, filter.dispatch.cpp, smooth.dispatch.cpp) and are
pinned against the installed ``cv2`` by ``tests/test_oracle_cv_semantics.py``.
Nothing here is imported by the product path.
"""
from __future__ import annotations

import numpy as np

INTER_BITS = 5
INTER_TAB = 1 << INTER_BITS          # 32 sub-pixel positions per axis
INT_MIN = -(1 << 31)


def _round_to_fixed(v32):
    """cvRound(v * 32) as x86 ``cvtps2dq`` does it: round-half-even, with NaN,
    +-inf and anything outside int32 mapped to INT_MIN ("integer indefinite")."""
    scaled = v32.astype(np.float32) * np.float32(INTER_TAB)
    bad = ~np.isfinite(scaled) | (scaled >= np.float32(2.0 ** 31)) | (scaled < np.float32(-2.0 ** 31))
    out = np.rint(np.where(bad, np.float32(0), scaled)).astype(np.int64)
    out[bad] = INT_MIN
    return out


def reflect(p, n):
    """BORDER_REFLECT (fedcba|abcdefgh|hgfedcb), applied until in range."""
    if n == 1:
        return np.zeros_like(p)
    q = np.mod(p, 2 * n)
    return np.where(q < n, q, 2 * n - 1 - q)


def reflect101(p, n):
    """BORDER_REFLECT_101 (gfedcb|abcdefgh|gfedcba), applied until in range."""
    if n == 1:
        return np.zeros_like(p)
    q = np.mod(p, 2 * n - 2)
    return np.where(q < n, q, 2 * n - 2 - q)


def fixed_point_coords(map_x, map_y):
    """(ix, iy, fx, fy): int16-saturated integer parts and 1/32 fractions."""
    sx, sy = _round_to_fixed(map_x), _round_to_fixed(map_y)
    ix = np.clip(sx >> INTER_BITS, -32768, 32767)
    iy = np.clip(sy >> INTER_BITS, -32768, 32767)
    return ix, iy, (sx & (INTER_TAB - 1)), (sy & (INTER_TAB - 1))


def bilinear_taps(src, ix, iy, fx, fy, border=reflect):
    """out = ((s00*w00 + s01*w01) + s10*w10) + s11*w11 in float32, the weight
    table being float32 products of (1-f, f) — ``remapBilinear`` scalar path."""
    h, w = src.shape[:2]
    x0, x1 = border(ix, w), border(ix + 1, w)
    y0, y1 = border(iy, h), border(iy + 1, h)
    ax = (fx.astype(np.float32) / np.float32(INTER_TAB))
    ay = (fy.astype(np.float32) / np.float32(INTER_TAB))
    one = np.float32(1)
    w00 = ((one - ay) * (one - ax))[..., None]
    w01 = ((one - ay) * ax)[..., None]
    w10 = (ay * (one - ax))[..., None]
    w11 = (ay * ax)[..., None]
    src = src.astype(np.float32, copy=False)
    if src.ndim == 2:
        src = src[..., None]
    acc = src[y0, x0] * w00
    acc = acc + src[y0, x1] * w01
    acc = acc + src[y1, x0] * w10
    acc = acc + src[y1, x1] * w11
    return acc


def remap_linear_reflect(src, map_x, map_y):
    """``cv2.remap(src, map_x, map_y, INTER_LINEAR, borderMode=BORDER_REFLECT)``
    for float32 sources and two float32 maps (stitcher.py:315-316)."""
    ix, iy, fx, fy = fixed_point_coords(map_x, map_y)
    out = bilinear_taps(src, ix, iy, fx, fy, reflect)
    return out if np.ndim(src) == 3 else out[..., 0]


def gaussian_ksize(sigma):
    """ksize chosen by ``GaussianBlur(img, (0, 0), sigma)`` for float images:
    cvRound(8*sigma + 1) | 1."""
    return int(np.rint(sigma * 8 + 1)) | 1


def gaussian_kernel(sigma, ksize=None):
    """``cv2.getGaussianKernel(ksize, sigma, CV_32F)``: exp(-x^2/2s^2)/sum
    evaluated in double, cast to float32."""
    ksize = gaussian_ksize(sigma) if ksize is None else ksize
    x = np.arange(ksize, dtype=np.float64) - (ksize - 1) / 2.0
    k = np.exp(-(x * x) / (2.0 * sigma * sigma))
    return (k / k.sum()).astype(np.float32)


def gaussian_blur(img, sigma):
    """``cv2.GaussianBlur(img, (0, 0), sigma)`` on float32 HxWxC with the
    default BORDER_REFLECT_101: horizontal pass then vertical pass, float32
    accumulation (stitcher.py:226)."""
    k = gaussian_kernel(sigma)
    r = (len(k) - 1) // 2
    h, w = img.shape[:2]
    cols = reflect101(np.arange(-r, w + r), w)
    rows = reflect101(np.arange(-r, h + r), h)
    tmp = np.zeros_like(img, dtype=np.float32)
    for t in range(len(k)):
        tmp += k[t] * img[:, cols[t:t + w]]
    out = np.zeros_like(tmp)
    for t in range(len(k)):
        out += k[t] * tmp[rows[t:t + h]]
    return out


def invert3x3(mat):
    """Closed-form 3x3 inverse in double, as ``cv::invert`` does for 3x3."""
    m = np.asarray(mat, dtype=np.float64)
    det = (m[0, 0] * (m[1, 1] * m[2, 2] - m[1, 2] * m[2, 1])
           - m[0, 1] * (m[1, 0] * m[2, 2] - m[1, 2] * m[2, 0])
           + m[0, 2] * (m[1, 0] * m[2, 1] - m[1, 1] * m[2, 0]))
    d = 1.0 / det
    out = np.empty((3, 3))
    out[0, 0] = (m[1, 1] * m[2, 2] - m[1, 2] * m[2, 1]) * d
    out[0, 1] = (m[0, 2] * m[2, 1] - m[0, 1] * m[2, 2]) * d
    out[0, 2] = (m[0, 1] * m[1, 2] - m[0, 2] * m[1, 1]) * d
    out[1, 0] = (m[1, 2] * m[2, 0] - m[1, 0] * m[2, 2]) * d
    out[1, 1] = (m[0, 0] * m[2, 2] - m[0, 2] * m[2, 0]) * d
    out[1, 2] = (m[0, 2] * m[1, 0] - m[0, 0] * m[1, 2]) * d
    out[2, 0] = (m[1, 0] * m[2, 1] - m[1, 1] * m[2, 0]) * d
    out[2, 1] = (m[0, 1] * m[2, 0] - m[0, 0] * m[2, 1]) * d
    out[2, 2] = (m[0, 0] * m[1, 1] - m[0, 1] * m[1, 0]) * d
    return out


def perspective_fixed_coords(hom, width, height):
    """Fixed-point source coordinates ``cv2.warpPerspective`` (INTER_LINEAR,
    forward matrix ``hom``) assigns to each destination pixel: the matrix is
    inverted, X = saturate_int(rint(32 * x'/w')) in double."""
    inv = invert3x3(hom)
    xs = np.arange(width, dtype=np.float64)[None, :]
    ys = np.arange(height, dtype=np.float64)[:, None]
    x0 = inv[0, 0] * xs + inv[0, 1] * ys + inv[0, 2]
    y0 = inv[1, 0] * xs + inv[1, 1] * ys + inv[1, 2]
    w0 = inv[2, 0] * xs + inv[2, 1] * ys + inv[2, 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        scale = np.where(w0 != 0, INTER_TAB / w0, 0.0)
    fx = np.clip(x0 * scale, float(INT_MIN), float((1 << 31) - 1))
    fy = np.clip(y0 * scale, float(INT_MIN), float((1 << 31) - 1))
    big_x = np.rint(fx).astype(np.int64)
    big_y = np.rint(fy).astype(np.int64)
    ix = np.clip(big_x >> INTER_BITS, -32768, 32767)
    iy = np.clip(big_y >> INTER_BITS, -32768, 32767)
    return ix, iy, big_x & (INTER_TAB - 1), big_y & (INTER_TAB - 1)


def warp_perspective_transparent(src, hom, width, height):
    """``cv2.warpPerspective(src, hom, (w, h), borderMode=BORDER_TRANSPARENT)``
    into a ZERO destination for a 4-channel float32 source.
    Pinned empirically against opencv 4.13.0: a destination pixel is written
    iff its top-left tap lies inside the source (0 <= ix <= w-1 and
    0 <= iy <= h-1); taps past the right/bottom edge are clamped (replicated);
    every other destination pixel keeps its previous (zero) value.
    Returns (warped, written)."""
    sh, sw = src.shape[:2]
    ix, iy, fx, fy = perspective_fixed_coords(hom, width, height)
    written = (ix >= 0) & (ix <= sw - 1) & (iy >= 0) & (iy <= sh - 1)
    val = bilinear_taps(src, np.where(written, ix, 0), np.where(written, iy, 0),
                        fx, fy, lambda p, n: np.clip(p, 0, n - 1))
    return np.where(written[..., None], val, np.float32(0)), written


# ---- cv2.resize(img, None, fx=1/S, fy=1/S) on uint8 (stitcher.py:418-421, the `-s` flag) --------
# OpenCV's published algorithm (modules/imgproc/src/resize.cpp): the destination size is
# cvRound(size * f) (round-half-even); INTER_LINEAR on 8-bit images runs in 11-bit fixed point
# (HResizeLinear / VResizeLinear with FixedPtCast<22>), except that an exact 2x shrink is rerouted
# to INTER_AREA's integer 2x2 mean (resizeAreaFast_).
RESIZE_COEF_BITS = 11
RESIZE_COEF_SCALE = 1 << RESIZE_COEF_BITS


def resize_dsize(h, w, f):
    """Size of cv2.resize(..., dsize=None, fx=f, fy=f): saturate_cast<int>(size * f) = cvRound."""
    return int(np.rint(h * f)), int(np.rint(w * f))


def resize_linear_tables(n_src, n_dst, f, clamp_fraction=True):
    """Per destination index: the first source index and the two fixed-point weights
    (int16) of the linear resampling along one axis — resize.cpp's xofs / ialpha (columns: an
    index clamped at an image edge gets the fraction 0) and yofs / ibeta (rows: the fraction is
    kept, the two row indices are clipped into the image by the row loop)."""
    scale = 1.0 / f
    d = np.arange(n_dst, dtype=np.float64)
    fx = ((d + 0.5) * scale - 0.5).astype(np.float32)
    sx = np.floor(fx).astype(np.int64)
    fx = (fx - sx.astype(np.float32)).astype(np.float32)
    if clamp_fraction:
        low = sx < 0
        fx[low], sx[low] = 0.0, 0
        high = sx >= n_src - 1
        fx[high], sx[high] = 0.0, n_src - 1

    def to_short(v):
        return np.clip(np.rint(v.astype(np.float32) * np.float32(RESIZE_COEF_SCALE)), -32768, 32767).astype(np.int64)
    return sx, to_short(np.float32(1.0) - fx), to_short(fx)


def resize_u8(img, f):
    """cv2.resize(img, None, fx=f, fy=f) (default INTER_LINEAR) for uint8 HxWxC, f <= 1."""
    h, w = img.shape[:2]
    dh, dw = resize_dsize(h, w, f)
    if (dh, dw) == (h, w):                         # "source and destination are of same size: simple copy"
        return img.copy()
    scale = 1.0 / f
    s = img.astype(np.int64)
    if abs(scale - 2.0) < np.finfo(np.float64).eps:
        # INTER_AREA (resizeAreaFast_): 2 x 2 integer mean with rounding; a block that sticks out of
        # an odd-sized image averages the pixels it has (float division, round-half-even)
        fh, fw = min(dh, h // 2), min(dw, w // 2)
        out = np.zeros((dh, dw) + img.shape[2:], np.uint8)
        out[:fh, :fw] = ((s[0:2 * fh:2, 0:2 * fw:2] + s[0:2 * fh:2, 1:2 * fw:2] + s[1:2 * fh:2, 0:2 * fw:2]
                          + s[1:2 * fh:2, 1:2 * fw:2] + 2) >> 2).astype(np.uint8)
        for dy in range(dh):
            for dx in range(dw):
                if dy < fh and dx < fw:
                    continue
                block = s[2 * dy:min(2 * dy + 2, h), 2 * dx:min(2 * dx + 2, w)]
                count = block.shape[0] * block.shape[1]
                mean = block.reshape(count, -1).sum(0).astype(np.float32) / np.float32(count)
                out[dy, dx] = np.clip(np.rint(mean), 0, 255).astype(np.uint8).reshape(img.shape[2:])
        return out
    sx, a0, a1 = resize_linear_tables(w, dw, f)
    sy, b0, b1 = resize_linear_tables(h, dh, f, clamp_fraction=False)
    x1 = np.minimum(sx + 1, w - 1)
    y0, y1 = np.clip(sy, 0, h - 1), np.clip(sy + 1, 0, h - 1)
    shape = (1, dw) + (1,) * (img.ndim - 2)
    rows = s[:, sx] * a0.reshape(shape) + s[:, x1] * a1.reshape(shape)          # horizontal pass: int, 11 fractional bits
    top, bot = rows[y0], rows[y1]
    shape = (dh, 1) + (1,) * (img.ndim - 2)
    out = (((b0.reshape(shape) * (top >> 4)) >> 16) + ((b1.reshape(shape) * (bot >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)
