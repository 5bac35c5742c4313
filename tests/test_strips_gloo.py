"""Multi-rank host logic of the strip sharding on CPU:
world_size 2 over gloo.  The kernels themselves are covered by the gpu tests
(test_row_window_equals_full_mosaic); here: the row partition, which images a
rank needs, and the final gather of uint8 strips onto rank 0."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pano360_b200 import geometry as geo, strips, synth


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, shape, parts, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        h, w = shape
        full = (torch.arange(h * w * 3, dtype=torch.int64) % 251).to(torch.uint8).view(h, w, 3)
        a, b = parts[rank]
        mosaic = strips.gather_strips(full[a:b].clone(), parts, shape, dst=0)
        if rank == 0:
            assert mosaic is not None and torch.equal(mosaic, full)
            np.save(out_path, mosaic.numpy())
        else:
            assert mosaic is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("parts", [[(0, 37), (37, 90)], [(0, 90), (90, 90)], [(0, 1), (1, 90)]])
def test_gather_strips_world2(tmp_path, parts):
    shape = (90, 41)
    out = str(tmp_path / "mosaic.npy")
    mp.spawn(_worker, args=(2, _free_port(), shape, parts, out), nprocs=2, join=True)
    got = np.load(out)
    want = (np.arange(90 * 41 * 3) % 251).astype(np.uint8).reshape(90, 41, 3)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("name,n", [("cfg4", 2), ("cfg4", 4), ("cfg4", 8), ("cfg3", 8), ("cfg1", 3)])
def test_partition_covers_mosaic_and_balances(name, n):
    wl = synth.workload(name)
    plan = geo.plan_mosaic(synth.camera_only(wl), wl.blend == "multiband", 1e9)
    parts = strips.partition_rows(plan, n, wl.blend, wl.n_levels)
    assert len(parts) == n and parts[0][0] == 0 and parts[-1][1] == plan.shape[0]
    assert all(a[1] == b[0] for a, b in zip(parts, parts[1:])) and all(b >= a for a, b in parts)
    per_row = strips.row_costs(plan, wl.blend, wl.n_levels)
    halo = strips.blur_halo(wl.blend, wl.n_levels)
    costs = [per_row[max(0, a - halo):b + halo].sum() for a, b in parts if b > a]
    assert max(costs) <= 1.15 * (sum(costs) / len(costs)) + per_row.max() * 2      # balanced within 15 %
    # every image lands in at least one strip, and strips only list images that touch them
    seen = set()
    for rows in parts:
        need = strips.images_for_rows(plan, rows, halo)
        seen.update(need)
        for i in need:
            x0, y0, x1, y1 = plan.boxes[i]
            assert y0 < rows[1] + halo and y1 > rows[0] - halo
    assert seen == set(range(wl.n_views))
    # column strips: whole 64-column tiles, covering the mosaic, balanced, every image somewhere
    if -(-plan.shape[1] // 64) >= n:
        cols = strips.partition_cols(plan, n, wl.blend, wl.n_levels)
        assert len(cols) == n and cols[0][2] == 0 and cols[-1][3] == plan.shape[1]
        assert all(a[3] == b[2] and a[2] % 64 == 0 and a[3] > a[2] for a, b in zip(cols, cols[1:] + [cols[-1]]) if a is not b)
        assert all(c[:2] == (0, plan.shape[0]) for c in cols)
        per_col = strips.col_costs(plan, wl.blend, wl.n_levels)
        halo = strips.col_halo(wl.blend, wl.n_levels)
        costs = [per_col[max(0, c[2] - halo):c[3] + halo].sum() for c in cols]
        assert max(costs) <= 1.15 * (sum(costs) / len(costs)) + per_col.max() * 128
        seen = set()
        for c in cols:
            seen.update(strips.images_for_part(plan, c, wl.blend, wl.n_levels))
        assert seen == set(range(wl.n_views))
        moved = strips.rebalance(cols, [1.0 + 0.3 * (k % 2) for k in range(n)], plan.shape[0])
        assert len(moved) == n and moved[0][2] == 0 and moved[-1][3] == plan.shape[1]
        assert all(a[3] == b[2] and a[3] % 64 == 0 for a, b in zip(moved, moved[1:]))


def test_single_rank_is_identity():
    strip = torch.zeros((5, 7, 3), dtype=torch.uint8)
    assert strips.gather_strips(strip, [(0, 5)], (5, 7)) is strip
    assert strips.partition_rows(geo.plan_mosaic(synth.camera_only(synth.workload("cfg1")), True, 1400), 1) \
        == [(0, 538)]
