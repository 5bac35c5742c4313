"""The GPU parity tests, re-run in the CPU tier against the kernels' own source compiled for
the host (tests/emul: every CUDA thread a fiber, barriers and warp collectives emulated).

What this checks without a GPU: the job tables the host code builds, the launch sequence,
and the logic of every kernel, against the same golden fixtures and oracle comparisons as
``-m gpu``.  What it cannot check: anything about timing, and the few float expressions
nvcc contracts to FMA on its own (multiband only; within the same tolerance).

Only the small cases are re-run here — the whole CPU suite has to stay within minutes.
The sm_100a build remains the only thing the package loads (tests/emul/harness.py patches
the binding from the outside).
"""
import pytest

from . import test_gpu_parity as gpu
from .emul import harness


@pytest.fixture()
def comp(monkeypatch):
    return harness.install(monkeypatch)


@pytest.fixture()
def st(comp):
    from pano360_b200 import stitcher
    return stitcher


tiny4 = gpu.tiny4
restore_globals = gpu.restore_globals

# the same test bodies, collected here without the gpu mark and bound to the fixtures above
test_golden_tiny4 = gpu.test_golden_tiny4
test_golden_levels_and_resolution_cap = gpu.test_golden_levels_and_resolution_cap
test_inputs_are_not_mutated = gpu.test_inputs_are_not_mutated
test_warp_stage_matches_reference_patches = gpu.test_warp_stage_matches_reference_patches
test_gains_match_reference = gpu.test_gains_match_reference
test_blur_stage_matches_cv2 = gpu.test_blur_stage_matches_cv2
test_owner_map_matches_oracle = gpu.test_owner_map_matches_oracle
test_coarse_levels_track_the_reference_blurs = gpu.test_coarse_levels_track_the_reference_blurs
test_blenders_accept_reference_style_patches = gpu.test_blenders_accept_reference_style_patches
test_foreign_blender_gets_numpy_patches = gpu.test_foreign_blender_gets_numpy_patches
test_oracle_seeded_cfg1_half = gpu.test_oracle_seeded_cfg1_half
test_golden_cfg1_full_size = gpu.test_golden_cfg1_full_size
test_golden_ring12_seam_straddlers = gpu.test_golden_ring12_seam_straddlers
test_cli_with_reference_style_caches = gpu.test_cli_with_reference_style_caches
test_two_row_six_band_layout = gpu.test_two_row_six_band_layout
test_many_small_views = gpu.test_many_small_views
test_edge_cases = gpu.test_edge_cases
test_crop_rectangle_matches_the_reference_scan = gpu.test_crop_rectangle_matches_the_reference_scan
test_device_resize_is_cv2_resize = gpu.test_device_resize_is_cv2_resize
test_cropped_stitch_matches_oracle = gpu.test_cropped_stitch_matches_oracle
test_row_window_equals_full_mosaic = gpu.test_row_window_equals_full_mosaic
test_row_windows_cut_anywhere = gpu.test_row_windows_cut_anywhere
test_column_windows_equal_full_mosaic = gpu.test_column_windows_equal_full_mosaic
test_windows_written_in_place = gpu.test_windows_written_in_place
test_tiles_final_after_the_tile_warp = gpu.test_tiles_final_after_the_tile_warp
test_source_rectangles_cover_every_tap = gpu.test_source_rectangles_cover_every_tap
test_view_over_the_pole = gpu.test_view_over_the_pole
test_unpacked_source_layout_is_equivalent = gpu.test_unpacked_source_layout_is_equivalent
test_batched_packing_equals_per_image = gpu.test_batched_packing_equals_per_image
test_partial_row_uploads_are_sufficient = gpu.test_partial_row_uploads_are_sufficient
test_seam_split_is_exact = gpu.test_seam_split_is_exact
test_full_resolution_path_is_the_reference_to_rounding = gpu.test_full_resolution_path_is_the_reference_to_rounding
test_tiny_views_with_sliver_owners = gpu.test_tiny_views_with_sliver_owners
test_seam_band_maps_are_exact = gpu.test_seam_band_maps_are_exact
test_seam_plan_is_conservative_and_cut_independent = gpu.test_seam_plan_is_conservative_and_cut_independent
test_window_without_any_image = gpu.test_window_without_any_image
test_blur_kernel_generic_taps = gpu.test_blur_kernel_generic_taps
test_batched_blur_paths = gpu.test_batched_blur_paths


def test_streamed_windows_pipeline(st, monkeypatch, restore_globals):
    """The end-to-end pipeline that overlaps the two PCIe directions (uploads ordered left edge
    first, only the rectangle of each image the seam plan reads; column windows composited and
    downloaded as their images arrive) returns the bytes of the plain stitch, for every blender
    and window count.  (The gpu tier runs it at full size: test_full_size_cfg3_windows_and_properties /
    the bench's checksum.)"""
    import torch
    regs = gpu.synth.make_views(gpu.synth.workload("cfg3", scale=8.0), noise=10.0)
    st.MAX_RESOLUTION = 10 ** 9
    monkeypatch.setattr(st, "STREAM_MIN_PIXELS", 0)
    for kind in ("multiband", "linear", "none"):
        monkeypatch.setattr(st, "STREAM_WINDOWS", 0)
        want = st.stitch(regs, blender=st.BLENDERS[kind])
        for windows in (2, 4):
            monkeypatch.setattr(st, "STREAM_WINDOWS", windows)
            out = torch.empty(want.shape, dtype=torch.uint8, pin_memory=True).numpy()
            out[:] = 7
            got = st.stitch(regs, blender=st.BLENDERS[kind], out=out)
            assert got is out and gpu.np.array_equal(got, want), (kind, windows)
    comp = st._compositor()
    plan = gpu.geo.plan_mosaic(regs, True, 1e9)
    for used in (None, comp.used_boxes(regs, plan, "multiband", 5)):
        order, wins = comp.streamed_windows(plan, "multiband", 5, 3, used=used)
        assert sorted(order) == list(range(len(regs))) and wins[-1][2] == len(regs) and len(wins) >= 2
        assert all(a[2] <= b[2] for a, b in zip(wins, wins[1:]))          # in the order their images arrive
        cover = sorted(w[:2] for w in wins)                                # ... they tile the mosaic's columns
        assert cover[0][0] == 0 and cover[-1][1] == plan.shape[1] and all(a[1] == b[0] for a, b in zip(cover, cover[1:]))
        assert all(a % 64 == 0 for a, _ in cover)


def test_streamed_window_plan_of_the_benchmark_ring(comp):
    """The window plan of the streamed pipeline on the benchmark rig (cameras only, full size): the
    six images that straddle the +-pi seam are uploaded last, the windows tile the mosaic's columns
    on tile edges in the order their images arrive, the first one waits for a single column of
    images, and the seam plan's rectangles save a sixth of the upload."""
    wl = gpu.synth.workload("cfg4")
    regs = gpu.synth.make_views(wl, only=set())
    plan = gpu.geo.plan_mosaic(regs, True, 1e9)
    rects = comp.source_rects(regs, plan, "multiband", wl.n_levels)
    used = comp.used_boxes(regs, plan, "multiband", wl.n_levels)
    order, wins = comp.streamed_windows(plan, "multiband", wl.n_levels, 12, used=used)
    straddlers = {i for i, boxes in used.items() if len(boxes) > 1}
    assert len(straddlers) == 6 and set(order[-6:]) == straddlers
    assert wins[0][2] <= 3 and wins[-1][2] == len(regs) and all(a[2] <= b[2] for a, b in zip(wins, wins[1:]))
    cover = sorted(w[:2] for w in wins)
    assert cover[0][0] == 0 and cover[-1][1] == plan.shape[1] and all(a[1] == b[0] and a[0] % 64 == 0 for a, b in zip(cover, cover[1:]))
    share = sum((r1 - r0) * (c1 - c0) for r0, r1, c0, c1 in rects.values()) / (len(regs) * wl.width * wl.height)
    assert 0.7 < share < 0.86, share


def test_pageable_buffers_are_staged(st, comp, monkeypatch, restore_globals):
    """Pageable inputs go through the ring of pinned slots (threaded memcpy + asynchronous DMA),
    a pageable / absent ``out`` through the pinned staging buffer, band by band — streamed and
    not: the bytes of the plain stitch."""
    regs = gpu.synth.make_views(gpu.synth.workload("cfg3", scale=8.0), noise=10.0)
    st.MAX_RESOLUTION = 10 ** 9
    monkeypatch.setattr(st, "STREAM_WINDOWS", 0)
    want = st.stitch(regs, blender=st.multiband_blend)
    monkeypatch.setattr(st, "_is_pinned_out", lambda out, shape: False)
    monkeypatch.setattr(st, "STREAM_MIN_PIXELS", 0)
    comp = st._compositor()
    comp.stage_min_bytes = 0
    for windows in (0, 3):
        monkeypatch.setattr(st, "STREAM_WINDOWS", windows)
        got = st.stitch(regs, blender=st.multiband_blend)
        assert gpu.np.array_equal(got, want), windows
        mine = gpu.np.full(want.shape, 9, gpu.np.uint8)
        assert st.stitch(regs, blender=st.multiband_blend, out=mine) is mine and gpu.np.array_equal(mine, want)
    assert len(comp._ring) == 3 and comp._out_stage is not None
    with pytest.raises(ValueError):
        st.stitch(regs, blender=st.multiband_blend, out=gpu.np.zeros((3, 3, 3), gpu.np.uint8))


def test_smoke_entry_point(comp, capsys):
    """__graft_entry__.smoke() — what the driver runs first on the GPU box — end to end."""
    import __graft_entry__
    __graft_entry__.smoke()
    assert "smoke multiband" in capsys.readouterr().out


def test_c_abi_error_path(comp):
    from pano360_b200 import _lib
    with pytest.raises(RuntimeError, match="p360_gauss_blur"):
        _lib.call("p360_gauss_blur", None, None, None, 4, 4, None, 3, None)


def test_barrier_misuse_is_detected():
    """The emulation aborts on a barrier not every live thread reaches; its own sanity check
    here is that a well-formed run leaves no fiber behind (a hang would time the test out)."""
    import numpy as np
    import torch
    from pano360_b200 import _lib
    lib = harness._load_library()
    keys = torch.zeros(64 * 33, dtype=torch.int64)
    owner = torch.empty(64 * 33, dtype=torch.int32)
    assert lib.p360_owner_decode(keys.data_ptr(), owner.data_ptr(), keys.numel(), None) == 0
    assert np.all(owner.numpy() == -1)
    assert set(_lib.SIGNATURES) <= {n for n in dir(lib) if n.startswith("p360_")} | set(_lib.SIGNATURES)


# ---- the multi-rank path end to end: strips over gloo, kernels on the host -------------------
def _strip_worker(rank, world, port, golden, kind, equalize, out_path, axis="cols"):
    import os
    os.environ["P360_SEAM_MAPS"] = "1" if rank % 2 else "0"      # ranks may differ: the bytes do not
    os.environ["P360_STRIPS"] = axis

    import numpy as np
    import torch.distributed as dist

    from pano360_b200 import strips
    from .conftest import load_golden, regions_from_golden
    patcher = pytest.MonkeyPatch()
    comp = harness.install(patcher)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), P360_EMUL_THREADS="2")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        regs = regions_from_golden(load_golden(golden))
        mosaic = strips.stitch_strips(comp, regs, kind, n_levels=5, equalize=equalize)
        assert (mosaic is not None) == (rank == 0)
        if rank == 0:
            mosaic = np.array(mosaic)           # (a view of the buffer the ranks share: the next call overwrites it)
        # the device-resident result: strips gathered on rank 0 by send / recv of row bands (what a
        # system without peer-mapped memory runs; a column strip arrives contiguous and is copied in)
        gathered = strips.stitch_strips(comp, regs, kind, n_levels=5, equalize=equalize, to_host=False)
        assert (gathered is not None) == (rank == 0)
        if rank == 0:
            assert np.array_equal(gathered.numpy(), mosaic)
            np.save(out_path, mosaic)
        # measured-feedback cuts: same cuts on every rank, still a partition of the mosaic
        from pano360_b200 import geometry as geo
        plan = geo.plan_mosaic_cached(regs, kind == "multiband", 1400)
        parts = strips.tune_partition(comp, regs, plan, kind, 5, lambda p: 1.0 + 0.5 * rank, rounds=2)
        box = [strips.part_box(p, plan.shape) for p in parts]
        assert len(parts) == world and box[0][0] == 0 and box[0][2] == 0 and box[-1][1] == plan.shape[0] and box[-1][3] == plan.shape[1]
        again = strips.stitch_strips(comp, regs, kind, n_levels=5, equalize=equalize)      # (with the tuned cuts)
        if rank == 0:
            assert np.array_equal(again, mosaic)
        dist.barrier()
    finally:
        dist.destroy_process_group()
        patcher.undo()


@pytest.mark.parametrize("golden,kind,equalize,world,axis", [
    ("tiny4", "multiband", False, 2, "cols"), ("tiny4", "linear", True, 2, "cols"), ("ring12", "multiband", True, 3, "cols"),
    ("ring12", "multiband", False, 3, "rows"), ("tiny4", "none", False, 2, "rows")])
def test_strips_over_gloo_equal_single_rank(st, tmp_path, golden, kind, equalize, world, axis):
    """stitch_strips on 2-3 ranks (column or row strips with halo, each rank downloading its strip
    into the host buffer the ranks share, pair statistics all-reduced) gives the bytes of the
    single-rank stitch — SURVEY.md §8(e)."""
    import numpy as np
    import torch.multiprocessing as mp

    from .conftest import load_golden, regions_from_golden
    from .test_strips_gloo import _free_port
    regs = regions_from_golden(load_golden(golden))
    want = st.stitch(regs, blender=st.BLENDERS[kind], equalize=equalize)
    out = str(tmp_path / "mosaic.npy")
    mp.spawn(_strip_worker, args=(world, _free_port(), golden, kind, equalize, out, axis), nprocs=world, join=True)
    assert np.array_equal(np.load(out), want)


def test_random_rigs_against_the_oracle(st, comp, restore_globals):
    """A fixed slice of tools/fuzz_host.py (random view counts, odd sizes, rings across the +-pi
    seam, steep pitches, roll, band counts, projections, forced seam-band maps, random row
    windows, the blender API on external patches); the full campaign runs from the tool."""
    import importlib.util
    import os

    import numpy as np
    spec = importlib.util.spec_from_file_location(
        "fuzz_host", os.path.join(os.path.dirname(__file__), "..", "tools", "fuzz_host.py"))
    fuzz = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fuzz)
    saved = comp.seam_maps, comp.direct
    try:
        for seed in range(900, 916):
            rng = np.random.default_rng(seed)
            case = fuzz.random_case(rng)
            whole = fuzz.run_case(st, comp, case)
            if whole is not None:
                fuzz.run_windows(comp, case, whole, rng)
                if not case["equalize"]:
                    fuzz.run_blender_api(st, case)
    finally:
        comp.seam_maps, comp.direct = saved


def test_seam_plan_candidates_at_full_scale(comp):
    """K0 alone on the benchmark geometry at FULL size (one thread per tile, no pixels needed)
    against the NumPy statement of the same interval arithmetic (tools/gate_bounds.py), plus the
    case that once broke it: at the right end of the ring image 24's box ends at column 31640,
    inside the last tile column, although the image itself continues (it wraps) — it must not
    out-bid the crop of image 35 that owns the columns beyond.  Also sizes what the plan buys:
    the share of solo tiles (written straight to uint8) and of float-warp tiles per patch."""
    import importlib.util
    import os

    import numpy as np
    import torch

    from pano360_b200 import _lib
    spec = importlib.util.spec_from_file_location(
        "gate_bounds", os.path.join(os.path.dirname(__file__), "..", "tools", "gate_bounds.py"))
    gb = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gb)
    wl = gpu.synth.workload("cfg4")
    regs = gpu.synth.make_views(wl, only=set())          # the benchmark's cameras, pixel-free stubs
    plan = gpu.geo.plan_mosaic(regs, True, 1e9)
    pad = gpu.geo.coarse_band_plan(5)[0]
    crops, (ray_x, ray_z, ray_y) = comp.plan_crops(regs, plan, split_dilate=2 * (pad + 4))
    rays = torch.from_numpy(np.concatenate([ray_x, ray_z, ray_y]))
    base = rays.data_ptr()
    jobs = np.zeros(len(crops), dtype=_lib.WARP_JOB)
    for k, (i, x0, y0, x1, y1, k_r, ty0, ty1, tx0, tx1) in enumerate(crops):
        h, w = regs[i].img.shape[:2]
        jobs[k]["ray_x"], jobs[k]["ray_z"], jobs[k]["ray_y"] = base, base + 8 * len(ray_x), base + 8 * (len(ray_x) + len(ray_z))
        jobs[k]["kr"], jobs[k]["h"], jobs[k]["w"] = k_r, h, w
        jobs[k]["pw"], jobs[k]["ph"], jobs[k]["x0"], jobs[k]["y0"] = x1 - x0, y1 - y0, x0, y0
        jobs[k]["col0"], jobs[k]["row0"], jobs[k]["patch"], jobs[k]["ty0"], jobs[k]["ty1"] = x0, y0, k, ty0, ty1
        jobs[k]["tx0"], jobs[k]["tx1"] = tx0, tx1
    n = len(crops)
    table = np.zeros(n, dtype=_lib.BAND_PATCH)
    table["w4"] = table["h4"] = 1
    maps, (bits, multi) = comp._tile_maps(table, 4, plan.shape[0], plan.shape[1], pad, 0, seam_plan=True)
    dev_jobs = torch.from_numpy(jobs.view(np.uint8).reshape(-1).copy())
    _lib.call("p360_seam_plan_build", dev_jobs.data_ptr(), n, None, plan.shape[0], plan.shape[1], 0, plan.shape[0],
              maps.ctypes.data, None)
    tx, ty, words = int(maps["tiles_x"][0]), int(maps["tiles_y"][0]), int(maps["words"][0])
    planes = bits.numpy().view(np.uint32)[4 + 2 * int(maps["work_cap"][0]):].reshape(4, ty, tx, words)
    unpack = lambda plane: np.stack([((plane[..., k >> 5] >> np.uint32(k & 31)) & 1).astype(bool) for k in range(n)])
    got = unpack(planes[0])
    want = gb.candidates(regs, plan, crops=[c[:5] for c in crops])
    assert got.shape == want.shape
    assert np.mean(got != want) < 1e-4 and not np.any(want & ~got & (want.sum(0) == 1)[None])
    crop41 = [k for k, c in enumerate(crops) if c[0] == 35 and c[3] == plan.shape[1]][0]
    assert got[crop41, 198, 494]
    single = (got.sum(0) == 1).mean()
    assert single > 0.85, single            # the bound is tight: most tiles have one possible owner
    cand, need, wneed = unpack(planes[1]), unpack(planes[2]), unpack(planes[3])
    is_multi = multi.numpy().reshape(ty, tx).astype(bool)
    assert np.array_equal(is_multi, cand.sum(0) > 1)
    assert not np.any(need & ~wneed) and not np.any(cand[:, is_multi] & ~need[:, is_multi])
    solo_share = 1.0 - is_multi.mean()
    boxes = np.stack([np.zeros((ty, tx), bool)] * n)
    for k, c in enumerate(crops):
        boxes[k, c[2] // 32:-(-c[4] // 32), c[1] // 64:-(-c[3] // 64)] = True
    float_share = (wneed & boxes).sum() / boxes.sum()
    print(f"cfg4 seam plan: {solo_share:.3f} of the tiles solo, float warp on {float_share:.3f} of the patch tiles")
    assert solo_share > 0.8 and float_share < 0.35
