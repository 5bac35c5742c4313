"""Multi-GPU compositing by strips of the output.

One process per GPU (``torch.distributed``, NCCL over NVLink/NVSwitch).  The
mosaic is split into contiguous strips of equal estimated cost — column strips
on 64-column tile edges for mosaics wider than tall (a 360-degree panorama: the
horizontal seams between pitch rows, where most of the blending work sits, are
then shared by all ranks), row strips otherwise (``P360_STRIPS=rows|cols``
forces one); every rank warps and blends only the images that overlap its strip
(plus a halo of the largest blur radius, re-warped locally — the warp is
pointwise, so no halo exchange is needed) and the uint8 strips are pushed band
by band into rank 0's mosaic over NVLink (peer-mapped destination) while the
rest of the strip is still being computed.  There is no other data-path
collective.  A strip is a pair ``(a, b)`` of rows or a 4-tuple ``(ya, yb, xa, xb)``.

The reference has no distributed code; this module is the B200-side answer to
its single-process ``stitch()`` (stitcher.py:274-327) for mosaics too large or
too slow for one device.
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as dist

from . import geometry as geo


def blur_halo(kind, n_levels):
    """Rows of context a strip needs above and below: the reach of the widest
    Gaussian the blender applies (stitcher.py:218, :226) plus one collapse tile
    (``Compositor.window_halo``) plus the alignment slack of cropped tops."""
    if kind != "multiband" or n_levels < 2:
        return 0
    pad = geo.coarse_band_plan(n_levels)[0]
    # (= Compositor.window_margin: windows are widened to the whole tiles the seam plan consults)
    return max(pad + 4 + 32, 32 * -(-pad // 32) + 31) + 3


def row_costs(plan, kind="multiband", n_levels=5):
    """Estimated work per mosaic row: patch pixels touching it (SURVEY.md H8 —
    equal-height strips are unbalanced: 12/12/24/24/24/24/12/12 images)."""
    height = plan.shape[0]
    cost = np.zeros(height + 1, dtype=np.float64)
    reach = geo.coarse_band_plan(n_levels)[0] + 4 if kind == "multiband" and n_levels > 1 else 0
    for i, (x0, y0, x1, y1) in enumerate(plan.boxes):
        if x1 <= x0 or y1 <= y0:
            continue
        # columns actually warped: seam-straddling boxes span the mosaic but are mostly empty
        width = sum(b - a for a, b in geo.active_column_runs(i, (x0, y0, x1, y1), plan, dilate=2 * reach))
        cost[max(y0, 0)] += width
        cost[min(y1, height)] -= width
    per_row = np.cumsum(cost[:-1])
    return per_row + plan.shape[1] * 0.25      # mosaic-sized passes (owner, collapse)


def partition_rows(plan, n_parts, kind="multiband", n_levels=5):
    """Split [0, H) into ``n_parts`` contiguous row ranges of ~equal cost,
    counting the halo rows each strip has to recompute."""
    height = plan.shape[0]
    if n_parts <= 1:
        return [(0, height)]
    per_row = row_costs(plan, kind, n_levels)
    prefix = np.concatenate([[0.0], np.cumsum(per_row)])
    halo = blur_halo(kind, n_levels)

    def strip_cost(a, b):
        lo, hi = max(0, a - halo), min(height, b + halo)
        return prefix[hi] - prefix[lo]

    # greedy bisection on the bottleneck cost
    lo_c, hi_c = 0.0, strip_cost(0, height)
    best = None
    for _ in range(40):
        mid = 0.5 * (lo_c + hi_c)
        cuts, a, ok = [], 0, True
        for _part in range(n_parts):
            b = a
            # largest b with cost(a, b) <= mid  (monotone in b)
            lo_b, hi_b = a, height
            while lo_b < hi_b:
                m = (lo_b + hi_b + 1) // 2
                if strip_cost(a, m) <= mid:
                    lo_b = m
                else:
                    hi_b = m - 1
            b = lo_b
            if b == a and a < height:
                ok = False
                break
            cuts.append((a, b))
            a = b
        if ok and a >= height:
            best, hi_c = cuts, mid
        else:
            lo_c = mid
    if best is None:
        edges = np.linspace(0, height, n_parts + 1).astype(int)
        best = [(int(edges[i]), int(edges[i + 1])) for i in range(n_parts)]
    # trailing parts may be empty when the greedy packing finishes early: rebalance
    best = [(a, b) for a, b in best]
    return best


def col_halo(kind, n_levels):
    """Columns of context a column strip needs on each side (= ``Compositor.col_margin``)."""
    if kind != "multiband" or n_levels < 2:
        return 0
    pad = geo.coarse_band_plan(n_levels)[0]
    return max(pad + 4 + 64, 64 * -(-pad // 64) + 63) + 64 + 3


def strip_axis(plan):
    """"cols" or "rows": which way the mosaic is cut."""
    forced = os.environ.get("P360_STRIPS", "")
    if forced in ("rows", "cols"):
        return forced
    return "cols" if plan.shape[1] >= plan.shape[0] else "rows"


def window_of(part, shape):
    """(rows, cols) arguments of ``Compositor.composite`` for a strip; None = the whole axis."""
    if len(part) == 2:
        return tuple(part), None
    ya, yb, xa, xb = part
    rows = None if (ya, yb) == (0, shape[0]) else (ya, yb)
    cols = None if (xa, xb) == (0, shape[1]) else (xa, xb)
    return rows, cols


def part_box(part, shape):
    """(ya, yb, xa, xb) of a strip in either form."""
    return (part[0], part[1], 0, shape[1]) if len(part) == 2 else tuple(part)


def is_empty(part):
    return part[1] <= part[0] or (len(part) == 4 and part[3] <= part[2])


def col_costs(plan, kind="multiband", n_levels=5):
    """Estimated work per mosaic column: patch pixels touching it."""
    height, width = plan.shape
    cost = np.zeros(width + 1, dtype=np.float64)
    reach = geo.coarse_band_plan(n_levels)[0] + 4 if kind == "multiband" and n_levels > 1 else 0
    for i, (x0, y0, x1, y1) in enumerate(plan.boxes):
        if x1 <= x0 or y1 <= y0:
            continue
        for a, b in geo.active_column_runs(i, (x0, y0, x1, y1), plan, dilate=2 * reach):
            cost[max(a, 0)] += y1 - y0
            cost[min(b, width)] -= y1 - y0
    return np.cumsum(cost[:-1]) + height * 0.25


def partition_cols(plan, n_parts, kind="multiband", n_levels=5):
    """Split [0, W) into ``n_parts`` contiguous column ranges on 64-column tile edges of ~equal
    cost, counting the halo columns each strip has to recompute.  Returns 4-tuples."""
    height, width = plan.shape
    if n_parts <= 1:
        return [(0, height, 0, width)]
    prefix = np.concatenate([[0.0], np.cumsum(col_costs(plan, kind, n_levels))])
    halo = col_halo(kind, n_levels)
    tiles = -(-width // 64)
    if tiles < n_parts:
        raise ValueError(f"a {width}-column mosaic cannot be cut into {n_parts} column strips of whole tiles")

    def cost(ta, tb):
        lo, hi = max(0, 64 * ta - halo), min(width, 64 * tb + halo)
        return prefix[hi] - prefix[lo]

    lo_c, hi_c = 0.0, cost(0, tiles)
    best = None
    for _ in range(40):
        mid = 0.5 * (lo_c + hi_c)
        cuts, a, ok = [], 0, True
        for k in range(n_parts):
            left = n_parts - k - 1                       # strips still to come: each needs a tile
            lo_b, hi_b = a + 1, tiles - left
            if lo_b > hi_b or cost(a, lo_b) > mid:
                ok = False
                break
            while lo_b < hi_b:
                m = (lo_b + hi_b + 1) // 2
                if cost(a, m) <= mid:
                    lo_b = m
                else:
                    hi_b = m - 1
            cuts.append((a, lo_b))
            a = lo_b
        if ok and a >= tiles:
            best, hi_c = cuts, mid
        else:
            lo_c = mid
    if best is None:
        edges = np.linspace(0, tiles, n_parts + 1).astype(int)
        best = [(int(edges[i]), int(edges[i + 1])) for i in range(n_parts)]
    return [(0, height, 64 * a, min(64 * b, width)) for a, b in best]


def rebalance(parts, times, height, align=32, damping=0.75):
    """New cuts from the measured device time of every strip (same cuts -> same bytes, so the
    partition is free to follow the measurement): the cost per row (column) is taken as constant
    within each old strip, the new cuts sit at equal shares of the total, moved ``damping`` of the
    way and rounded to whole tiles (32 rows; 64 columns for column strips, where ``height`` is
    ignored: 4-tuples carry their own extent).  A cut inside a tile makes two ranks plan and warp it."""
    n = len(parts)
    if n < 2 or min(times) <= 0:
        return list(parts)
    by_cols = len(parts[0]) == 4
    if by_cols:
        align, height, rows = 64, parts[-1][3], parts[0][:2]
        edges = [p[2] for p in parts] + [parts[-1][3]]
    else:
        edges = [p[0] for p in parts] + [parts[-1][1]]
    cum = np.concatenate([[0.0], np.cumsum(times)])
    new = [0]
    for k in range(1, n):
        target = cum[-1] * k / n
        r = min(int(np.searchsorted(cum, target, side="right")) - 1, n - 1)
        span = edges[r + 1] - edges[r]
        y = edges[r] + (target - cum[r]) / max(times[r], 1e-12) * span
        y = edges[k] + damping * (y - edges[k])
        y = int(round(y / align)) * align
        new.append(min(max(y, new[-1] + align), height - align * (n - k)))
    new.append(height)
    if by_cols:
        return [(rows[0], rows[1], new[i], new[i + 1]) for i in range(n)]
    return [(new[i], new[i + 1]) for i in range(n)]


def partition(plan, n_parts, kind="multiband", n_levels=5):
    """The model's cuts along ``strip_axis(plan)``."""
    if strip_axis(plan) == "cols" and -(-plan.shape[1] // 64) >= n_parts:
        return partition_cols(plan, n_parts, kind, n_levels)
    return partition_rows(plan, n_parts, kind, n_levels)


def tune_partition(comp, regions, plan, kind, n_levels, step, group=None, rounds=4, log=None):
    """Measured-feedback strip cuts (collective): ``step(parts)`` runs the composite of this
    rank's strip for the given cuts (uploading what it needs; untimed by the caller) and returns
    its device time in ms; the cuts are moved up to ``rounds`` times and the best partition seen
    (smallest time of the slowest rank — the model's cuts included) is remembered in the plan,
    where ``stitch_strips`` / ``strip_cuts`` find it.  Returns the cuts."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    parts = partition(plan, world, kind, n_levels)
    best = None
    if world > 1:
        damping = 0.6
        for it in range(rounds + 1):
            mine = torch.zeros(world, dtype=torch.float64, device=comp.device)
            mine[rank] = step(parts)
            dist.all_reduce(mine, group=group)
            times = mine.cpu().tolist()
            if log is not None:
                log.append({"cuts": [list(p) for p in parts], "ms": [round(t, 3) for t in times]})
            if best is None or max(times) < max(best[1]):
                best = (parts, times)
            else:
                damping *= 0.5                       # overshot: a smaller step from the best cuts so far
            if it == rounds or min(best[1]) <= 0 or max(best[1]) <= 1.04 * (sum(best[1]) / world):
                break
            parts = rebalance(best[0], best[1], plan.shape[0], damping=damping)
            if parts == best[0]:
                break
        parts = best[0]
    if getattr(plan, "_parts", None) is None:
        plan._parts = {}
    plan._parts[(world, kind, n_levels)] = parts
    return parts


def strip_cuts(plan, world, kind, n_levels):
    """The cuts in force for this plan: tuned ones if ``tune_partition`` ran, else the model's."""
    if getattr(plan, "_parts", None) is None:
        plan._parts = {}
    key = (world, kind, n_levels)
    if key not in plan._parts:
        plan._parts[key] = partition(plan, world, kind, n_levels)
    return plan._parts[key]


def images_for_rows(plan, rows, halo):
    """Indices of images whose box intersects rows [ya - halo, yb + halo)."""
    ya, yb = rows[0] - halo, rows[1] + halo
    return [i for i, (x0, y0, x1, y1) in enumerate(plan.boxes) if y0 < yb and y1 > ya and x1 > x0]


def images_for_part(plan, part, kind, n_levels):
    """Indices of the images a strip (either form) can depend on: box rows within the row halo,
    a column run within the column halo."""
    ya, yb, xa, xb = part_box(part, plan.shape)
    if yb <= ya or xb <= xa:
        return []
    hr, hc = blur_halo(kind, n_levels), col_halo(kind, n_levels)
    reach = geo.coarse_band_plan(n_levels)[0] + 4 if kind == "multiband" and n_levels > 1 else 0
    out = []
    for i, (x0, y0, x1, y1) in enumerate(plan.boxes):
        if x1 <= x0 or not (y0 < yb + hr and y1 > ya - hr):
            continue
        if len(part) == 2 or any(a < xb + hc and b > xa - hc
                                 for a, b in geo.active_column_runs(i, (x0, y0, x1, y1), plan, dilate=2 * reach)):
            out.append(i)
    return out


def gather_strips(strip, parts, shape, dst=0, group=None):
    """Assemble the full uint8 mosaic on rank ``dst`` from one strip per rank.
    Strips land directly in their row slice of the destination buffer
    (row-major, so each slice is contiguous): one grouped send/recv."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    h, w = shape
    if world == 1:
        return strip
    if rank == dst:
        mosaic = torch.empty((h, w, 3), dtype=torch.uint8, device=strip.device)
        ops = []
        for r, (a, b) in enumerate(parts):
            if b <= a:
                continue
            if r == dst:
                mosaic[a:b].copy_(strip)
            else:
                ops.append(dist.P2POp(dist.irecv, mosaic[a:b], r, group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return mosaic
    a, b = parts[rank]
    if b > a:
        for req in dist.batch_isend_irecv([dist.P2POp(dist.isend, strip.contiguous(), dst, group)]):
            req.wait()
    return None


def final_after_warp(multi, geometry, seam_share=0.5):
    """Cut a window into rectangles that are final once the tile warp has run (no multi tile
    inside: ``early``) and the rest (``late``: final after the collapse).  ``multi``: bool
    [tiles_y, tiles_x] over the 64 x 32 tiles of the window's buffer, ``geometry`` as handed to
    ``Compositor.composite(after_warp=...)``.  Tile rows in which more than ``seam_share`` of the
    tiles are multi (a horizontal seam band) go late whole; between them, runs of tile columns
    without a multi tile go early.  Rectangles are (y0, y1, x0, x1) in BUFFER pixels, disjoint,
    and cover the window exactly."""
    (ya, yb), (xa, xb), row0 = geometry["rows"], geometry["cols"], geometry["row0"]
    ty0, ty1 = (ya - row0) // 32, (yb - 1 - row0) // 32 + 1
    tx0, tx1 = xa // 64, (xb - 1) // 64 + 1
    sub = multi[ty0:ty1, tx0:tx1]
    early, late = [], []
    if sub.size == 0:
        return early, late
    heavy = sub.mean(axis=1) > seam_share
    a = 0
    for t in range(1, len(heavy) + 1):
        if t == len(heavy) or heavy[t] != heavy[a]:
            y0, y1 = max(ya, row0 + 32 * (ty0 + a)), min(yb, row0 + 32 * (ty0 + t))
            if heavy[a]:
                late.append((y0, y1, xa, xb))
            else:
                busy = sub[a:t].any(axis=0)
                c = 0
                for u in range(1, len(busy) + 1):
                    if u == len(busy) or busy[u] != busy[c]:
                        x0, x1 = max(xa, 64 * (tx0 + c)), min(xb, 64 * (tx0 + u))
                        (late if busy[c] else early).append((y0, y1, x0, x1))
                        c = u
            a = t
    return early, late


class PeerMosaic:
    """Rank 0's mosaic buffer mapped into every rank's address space over
    NVLink / NVSwitch (torch symmetric memory): strips are written by DMA
    straight into their rows of the destination, band by band, while the next
    band is still being computed — no staging, no NCCL kernel."""

    _cache = {}

    def __init__(self, device, group):
        self.device, self.group = device, group
        self.capacity, self.handle, self.local, self.flip = 0, None, None, 0

    @classmethod
    def get(cls, device, group):
        key = (str(device), id(group))
        if key not in cls._cache:
            cls._cache[key] = cls(device, group)
        return cls._cache[key]

    def ensure(self, nbytes):
        """Collective: (re)allocate the symmetric buffer when it is too small."""
        if nbytes <= self.capacity:
            return
        import torch.distributed._symmetric_memory as symm
        cap = (int(nbytes * 1.25) + (1 << 21) - 1) >> 21 << 21
        # two mosaics: a step writes the one the previous step did not, so rank 0 may still be
        # reading the last result while the next one arrives — one barrier per step, not two
        self.local = symm.empty(2 * cap, dtype=torch.uint8, device=self.device)
        self.handle = symm.rendezvous(self.local, self.group if self.group is not None else dist.group.WORLD)
        self.capacity = cap

    def rows_of_rank0(self, h, w):
        self.flip ^= 1
        return self.handle.get_buffer(0, (h, w, 3), torch.uint8, storage_offset=self.flip * self.capacity)

    def barrier(self):
        self.handle.barrier(channel=0)


_peer_ok = True
# P360_FUSED_GATHER=1: strips are stored into rank 0's mosaic by the kernels that produce them (peer
# stores over NVLink, no copy pass).  Measured on 8 x B200, cfg4 (profiles/r02o_*): byte-identical, but
# 3.24 ms per step against 2.05 ms for the default — strips composited locally and pushed band by band
# with the copy engines (rectangle DMA): a tile row is a 192-byte segment and the collapse stores
# single bytes, which NVLink carries at a third of the rate of the DMA's long bursts.
FUSED_GATHER = os.environ.get("P360_FUSED_GATHER", "0") == "1"
# Push everything outside the seam zone right after the tile warp (final from then on), so that the
# transfer overlaps reduce / blur / collapse and only the seam zone's rectangles are left for the end.
EARLY_PUSH = os.environ.get("P360_EARLY_PUSH", "1") == "1"
_early_cache = {}


def composite_gather(comp, regions, src, plan, kind, n_levels, parts, proj=geo.SphProj, group=None,
                     bands=4):
    """Strip composite + gather, overlapped.  Preferred transport: rank 0's mosaic mapped into
    every rank (``PeerMosaic``) and copy-engine DMA into it — everything outside the seam zone
    right after the tile warp, when those tiles are final (``final_after_warp``), the seam zone's
    rectangles after the collapse; without the seam plan, row bands as they are collapsed.
    ``P360_FUSED_GATHER=1``: the kernels store into the peer mosaic themselves.  If symmetric
    memory cannot be set up: grouped NCCL send/recv of row bands.  Returns the device mosaic on
    rank 0 (valid until the call after the next), None elsewhere."""
    global _peer_ok
    from .compositor import band_edges
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world == 1:
        return comp.composite(regions, src, plan, kind, n_levels, proj)[0]
    h, w = plan.shape
    part = parts[rank]
    rows, cols = window_of(part, plan.shape)
    peer = None
    if _peer_ok and comp.device.type == "cuda":
        try:
            peer = PeerMosaic.get(comp.device, group)
            peer.ensure(h * w * 3)
        except Exception as exc:                      # no symmetric memory on this system
            import logging
            logging.getLogger(__name__).warning("peer-mapped gather unavailable (%s); using NCCL send/recv", exc)
            _peer_ok, peer = False, None

    def place(dst, piece, y0, y1):
        """strip rows [y0, y1) (mosaic coordinates) into their place in a whole-mosaic tensor: a DMA
        (rows: one contiguous copy; a column strip: a rectangle copy — an elementwise copy kernel
        would write the peer's memory byte by byte)"""
        if cols is None:
            dst[y0:y1].copy_(piece, non_blocking=True)
        elif comp.device.type != "cuda":
            dst[y0:y1, cols[0]:cols[1]].copy_(piece)
        else:
            from . import _lib
            _lib.call("p360_copy_rect", dst.data_ptr() + 3 * (y0 * w + cols[0]), 3 * w, piece.data_ptr(), piece.stride(0),
                      3 * (cols[1] - cols[0]), y1 - y0, torch.cuda.current_stream(comp.device).cuda_stream)
    if peer is not None:
        main, side = torch.cuda.current_stream(comp.device), comp.copy_stream()
        dst = peer.rows_of_rank0(h, w)                # (the buffer the previous step did not write)
        exact = kind == "multiband" and n_levels > 1 and comp.needs_exact(regions)
        if not is_empty(part):
            if FUSED_GATHER and not exact:
                # compute + gather in one: the tile warp and the collapse store the strip's bytes
                # straight into their place in rank 0's mosaic (peer stores over NVLink / NVSwitch,
                # tile by tile as they are produced) — no strip buffer, no copy pass
                comp.composite(regions, src, plan, kind, n_levels, proj, rows=rows, cols=cols,
                               out_dev=(dst.data_ptr(), w, dst))
            elif rank == 0:
                strip, _ = comp.composite(regions, src, plan, kind, n_levels, proj, rows=rows, cols=cols, exact=exact)
                ya = part_box(part, plan.shape)[0]
                place(dst, strip, ya, ya + strip.shape[0])
            else:
                state = {}

                def push_rects(buffer, rects, geometry):
                    """DMA rectangles of the window's buffer into their place in rank 0's mosaic"""
                    from . import _lib
                    done = torch.cuda.Event()
                    done.record(main)
                    side.wait_event(done)
                    top, left, pitch = geometry["top"], geometry["left"], buffer.shape[1]
                    for y0, y1, x0, x1 in rects:
                        _lib.call("p360_copy_rect", dst.data_ptr() + 3 * ((y0 + top) * w + x0 + left), 3 * w,
                                  buffer.data_ptr() + 3 * (y0 * pitch + x0), 3 * pitch, 3 * (x1 - x0), y1 - y0, side.cuda_stream)

                def after_warp(buffer, multi, geometry):
                    # everything but the seam zone is final now: it travels while reduce / blur /
                    # collapse run, and only the seam zone's rectangles are left for the end
                    key = (id(multi), geometry["top"], geometry["left"], geometry["rows"], geometry["cols"])
                    entry = _early_cache.get(key)
                    if entry is None or entry[0] is not multi:       # (an id can be reused: the map itself is kept and compared)
                        if len(_early_cache) > 64:
                            _early_cache.clear()
                        entry = _early_cache[key] = (multi,) + final_after_warp(multi, geometry)
                    state.update(buffer=buffer, geometry=geometry, late=entry[2])
                    push_rects(buffer, entry[1], geometry)

                def push_band(piece, y0, y1):
                    if state:                         # (the early pushes fired: the rest goes at the end)
                        return
                    done = torch.cuda.Event()
                    done.record(main)
                    side.wait_event(done)
                    with torch.cuda.stream(side):
                        place(dst, piece, y0, y1)
                early_ok = EARLY_PUSH and comp.device.type == "cuda" and not exact
                comp.composite(regions, src, plan, kind, n_levels, proj, rows=rows, cols=cols, on_band=push_band,
                               bands=bands, exact=exact, after_warp=after_warp if early_ok else None)
                if state:
                    push_rects(state["buffer"], state["late"], state["geometry"])
                main.wait_stream(side)
        peer.barrier()                                # every strip has landed
        return dst if rank == 0 else None
    if rank != 0:
        works = []

        def send_band(piece, y0, y1):
            works.extend(dist.batch_isend_irecv([dist.P2POp(dist.isend, piece.contiguous(), 0, group)]))
        if not is_empty(part):
            comp.composite(regions, src, plan, kind, n_levels, proj, rows=rows, cols=cols, on_band=send_band, bands=bands)
        for work in works:
            work.wait()
        return None
    mosaic = torch.empty((h, w, 3), dtype=torch.uint8, device=comp.device)
    if not is_empty(part):
        strip, _ = comp.composite(regions, src, plan, kind, n_levels, proj, rows=rows, cols=cols)
        ya = part_box(part, plan.shape)[0]
        (mosaic[ya:ya + strip.shape[0]] if cols is None else mosaic[ya:ya + strip.shape[0], cols[0]:cols[1]]).copy_(strip)
    works, landing = [], []
    for k in range(bands):
        ops = []
        for r in range(1, world):
            if is_empty(parts[r]):
                continue
            a, b, xa, xb = part_box(parts[r], plan.shape)
            y0, y1 = band_edges(a, b, bands)[k]
            if y1 > y0:
                if len(parts[r]) == 2 or (xa, xb) == (0, w):
                    ops.append(dist.P2POp(dist.irecv, mosaic[y0:y1], r, group))
                else:                                 # a column strip arrives contiguous and is copied into place
                    tmp = torch.empty((y1 - y0, xb - xa, 3), dtype=torch.uint8, device=comp.device)
                    landing.append((tmp, y0, y1, xa, xb))
                    ops.append(dist.P2POp(dist.irecv, tmp, r, group))
        if ops:
            works.extend(dist.batch_isend_irecv(ops))
    for work in works:
        work.wait()
    for tmp, y0, y1, xa, xb in landing:
        mosaic[y0:y1, xa:xb].copy_(tmp)
    return mosaic


def all_pair_statistics(comp, regions, src, group=None):
    """Exposure-gain statistics with the image pairs dealt round-robin over
    the ranks; the three sums per pair are all-reduced so that every rank
    solves the same N x N system (stitcher.py:36-66)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world == 1:
        overlaps, sizes, _ = comp.pair_statistics(regions, src)
        return overlaps, sizes
    n = len(regions)
    n_pairs = n * (n - 1) // 2
    mine = set(range(rank, n_pairs, world))
    overlaps, sizes, _ = comp.pair_statistics(regions, src, pairs=mine)
    packed = torch.from_numpy(np.stack([overlaps, sizes])).to(comp.device)
    dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    both = packed.cpu().numpy()
    return both[0], both[1]


class SharedHostMosaic:
    """One mosaic-sized buffer in host shared memory, mapped by every rank of the node, with each
    rank's own rows registered as page-locked: every rank downloads its strip over its own PCIe
    link, band by band while it is still computing, instead of funnelling the mosaic through
    rank 0's link.  (Re)allocation is collective; the mapping is kept and grown as needed."""

    _cache = {}

    def __init__(self, group):
        self.group, self.capacity, self.map, self.registered = group, 0, None, None

    @classmethod
    def get(cls, group):
        key = id(group)
        if key not in cls._cache:
            cls._cache[key] = cls(group)
        return cls._cache[key]

    def ensure(self, nbytes):
        if nbytes <= self.capacity:
            return
        import os
        rank = dist.get_rank(self.group)
        self.release()
        cap = (int(nbytes * 1.25) + (1 << 21) - 1) >> 21 << 21
        name = [f"/dev/shm/p360_mosaic_{os.getpid()}_{cap}" if rank == 0 else None]
        if rank == 0:
            with open(name[0], "wb") as fid:
                fid.truncate(cap)
        dist.broadcast_object_list(name, src=0, group=self.group)
        self.map = np.memmap(name[0], dtype=np.uint8, mode="r+", shape=(cap,))
        dist.barrier(self.group)                     # everybody has mapped it: the name can go
        if rank == 0:
            os.unlink(name[0])
        self.capacity = cap

    def rows(self, h, w):
        return self.map[:h * w * 3].reshape(h, w, 3)

    def register(self, lo, hi):
        """Page-lock bytes [lo, hi) of the mapping (rounded out to pages) for asynchronous DMA."""
        page = 4096
        lo, hi = lo // page * page, min(-(-hi // page) * page, self.capacity)
        if self.registered is not None and self.registered[0] <= lo and hi <= self.registered[1]:
            return
        self.unregister()
        rt = torch.cuda.cudart()
        err = rt.cudaHostRegister(self.map.ctypes.data + lo, hi - lo, 0)
        if int(err) != 0:
            raise RuntimeError(f"cudaHostRegister failed: {err}")
        self.registered = (lo, hi)

    def unregister(self):
        if self.registered is not None:
            torch.cuda.cudart().cudaHostUnregister(self.map.ctypes.data + self.registered[0])
            self.registered = None

    def release(self):
        self.unregister()
        self.map, self.capacity = None, 0


def stitch_strips(comp, regions, kind, n_levels=5, equalize=False, max_resolution=1400,
                  proj=geo.SphProj, group=None, to_host=True, out=None):
    """Collective: every rank calls this with the same ``regions``; rank 0
    gets the full mosaic (NumPy uint8 if ``to_host`` else a device tensor),
    other ranks get None.  Each rank uploads only the images its strip needs
    (the warp starts while the last ones are still crossing PCIe) and — with
    ``to_host`` — downloads its own strip into a host buffer shared by the
    ranks, so that N PCIe links carry the mosaic.  The array rank 0 gets back
    is ``out`` if given, else a view of that shared buffer (valid until the
    next call)."""
    from .compositor import parallel_copy
    from .stitcher import find_gains
    import time
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    phases = [("begin", time.perf_counter())] if comp.phases is not None else None

    def phase(label):
        if phases is not None:
            phases.append((label, time.perf_counter()))
    plan = geo.plan_mosaic_cached(regions, kind == "multiband", max_resolution, proj)
    parts = strip_cuts(plan, world, kind, n_levels)
    part = parts[rank]
    rows, cols = window_of(part, plan.shape)
    need = set(images_for_part(plan, part, kind, n_levels))
    if equalize:
        need = set(range(len(regions)))          # pair statistics touch every image
    # a strip reads only part of every image it meets: upload (and pack) just that
    exact = kind == "multiband" and n_levels > 1 and comp.needs_exact(regions)
    parts_of = None
    if not equalize and not is_empty(part):
        parts_of = None if exact else comp.source_rects(regions, plan, kind, n_levels, proj, rows=rows, cols=cols)
        if parts_of is None:
            parts_of = comp.source_rows(regions, plan, kind, n_levels, proj, rows=rows, cols=cols)
        need &= set(parts_of)
    phase("planned")
    src = comp.upload(regions, need=need, overlap=not equalize, rows_of=parts_of, reuse=True)
    phase("uploads queued")
    if equalize:
        overlaps, sizes = all_pair_statistics(comp, regions, src, group)
        comp.set_gains(src, find_gains(overlaps, sizes))
    if not to_host or world == 1:
        mosaic = composite_gather(comp, regions, src, plan, kind, n_levels, parts, proj, group)
        if rank != 0 or mosaic is None:
            return None
        if not to_host:
            return mosaic
        from .stitcher import _download
        host = _download(mosaic, out)
        comp.release()
        return host
    h, w = plan.shape
    shared = SharedHostMosaic.get(group)
    shared.ensure(h * w * 3)                     # collective on first use / growth
    host = shared.rows(h, w)
    if not is_empty(part):
        if comp.device.type == "cuda":
            ya, yb = part_box(part, plan.shape)[:2]
            shared.register(ya * w * 3, yb * w * 3)      # (a column strip spans every row: the whole mapping)
        comp.composite(regions, src, plan, kind, n_levels, proj, rows=rows, cols=cols, out_host=host, bands=4,
                       exact=exact)
        phase("kernels queued")
        comp.finish_download()
        phase("strip landed")
    comp.release()
    dist.barrier(group)                          # every strip has landed
    phase("barrier")
    if phases is not None:
        comp.phases.append([(label, round((t - phases[0][1]) * 1e3, 3)) for label, t in phases[1:]])
    if rank != 0:
        return None
    if out is not None:
        parallel_copy(out, host)
        return out
    return host
