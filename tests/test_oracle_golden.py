"""The CPU oracle (oracle/restate.py) against the committed fixtures that the
LIVE reference produced (oracle/make_golden.py).  Bit-exact on uint8."""
import numpy as np
import pytest

from oracle import restate as rs
from .conftest import load_golden, regions_from_golden

CASES = [(b, e, p) for b in ("none", "linear", "multiband") for e in (False, True)
         for p in ("spherical", "cylindrical")]


@pytest.fixture(scope="module")
def tiny4():
    data = load_golden("tiny4")
    return data, regions_from_golden(data)


@pytest.mark.parametrize("blend,eq,proj", CASES)
def test_tiny4_all_modes(tiny4, blend, eq, proj):
    data, regs = tiny4
    want = data[f"mosaic_{blend}_{'eq' if eq else 'raw'}_{proj[:3]}"]
    got = rs.stitch(regs, blend, eq, 5, 1400, proj)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("key,levels,cap", [("mosaic_multiband_L6_uncapped", 6, 1e9),
                                            ("mosaic_multiband_L1", 1, 1400),
                                            ("mosaic_multiband_L2", 2, 1400)])
def test_tiny4_levels(tiny4, key, levels, cap):
    data, regs = tiny4
    assert np.array_equal(rs.stitch(regs, "multiband", False, levels, cap), data[key])
    assert np.array_equal(rs.stitch(regs, "multiband", False, levels, cap, owner_mode="stream"), data[key])


def test_tiny4_patches_and_gains(tiny4):
    data, regs = tiny4
    for blend in ("linear", "multiband"):
        patches, pl = rs.build_patches(regs, blend)
        assert tuple(data[f"patch_shape_{blend}"]) == pl.shape
        for i, (warped, invalid, (sy, sx)) in enumerate(patches):
            assert list(data[f"patch_{blend}_{i}_box"]) == [sx.start, sy.start, sx.stop, sy.stop]
            assert np.array_equal(data[f"patch_{blend}_{i}_mask"], invalid)
        assert np.array_equal(data[f"patch_{blend}_1_warped"], patches[1][0])
    rgba = [rs.rgba_with_weights(r.img) for r in regs]
    gains, overlaps, sizes = rs.equalize(regs, rgba)
    assert np.array_equal(sizes, data["gain_sizes"])
    np.testing.assert_array_equal(overlaps, data["gain_overlaps"])
    np.testing.assert_allclose(gains, data["gains"], rtol=1e-12)


def test_ring12_seam_straddlers():
    data = load_golden("ring12")
    regs = regions_from_golden(data)
    for blend in ("none", "linear", "multiband"):
        got = rs.stitch(regs, blend, False, 5, 1e9)
        assert np.array_equal(got, data[f"mosaic_{blend}"]), blend


def test_cfg1_full_size():
    from oracle.make_golden import cfg1_inputs, inputs_digest
    data = load_golden("cfg1")
    regs = cfg1_inputs()
    if inputs_digest(regs) != str(data["digest"]):
        pytest.skip("synthetic generator produces different pixels on this machine")
    assert np.array_equal(rs.stitch(regs, "multiband", False, 5, 1400), data["mosaic_multiband"])


def test_window_mode_is_exact(tiny4):
    data, regs = tiny4
    full = data["mosaic_multiband_L6_uncapped"]
    h, w = full.shape[:2]
    for win in [(0, 30, 0, w), (h - 25, h, 10, w - 10), (h // 3, h // 3 + 40, w // 4, w // 4 + 90)]:
        got = rs.stitch_window(regs, win, "multiband", False, 6, 1e9)
        assert np.array_equal(got, full[win[0]:win[1], win[2]:win[3]]), win
    lin = rs.stitch(regs, "linear", False, 5, 1e9)
    got = rs.stitch_window(regs, (5, 50, 7, 200), "linear", False, 5, 1e9)
    assert np.array_equal(got, lin[5:50, 7:200])


def test_numpy_backend_within_tolerance(tiny4):
    """The pure-NumPy restatement of the OpenCV primitives stays within the
    north_star tolerance of the reference (blur rounding only)."""
    data, regs = tiny4
    got = rs.stitch(regs, "multiband", True, 5, 1400, backend="numpy")
    want = data["mosaic_multiband_eq_sph"]
    assert np.abs(got.astype(int) - want.astype(int)).max() <= 1
    for blend in ("none", "linear"):
        assert np.array_equal(rs.stitch(regs, blend, False, 5, 1400, backend="numpy"),
                              data[f"mosaic_{blend}_raw_sph"])
