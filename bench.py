#!/usr/bin/env python
"""Benchmark of the compositing hot path (warp + blend) — see DESIGN.md §Measurement.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg4] [--impl reference]

One "step" = one full composite of one synthetic panorama: every view warped
onto the spherical mosaic and blended, final uint8 mosaic resident in HBM on
rank 0 (N > 1: strip-sharded over the ranks, strips gathered over NCCL at the
end of every step).  ``value`` is output megapixels per second with the source
images already resident in HBM; ``e2e`` is the same metric through the public
drop-in API with host (pinned) buffers, H2D of every source and D2H of the
mosaic inside the timed region.

``--impl reference`` times the reference's CPU path (the live reference if its
checkout is present, else the oracle port that makes the same NumPy/OpenCV
calls) on a bounded sample of the same workload, on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from dataclasses import replace

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "output Mpix/s warp+multiband blend (device-timed)"
UNIT = "Mpix/s"

DESCRIPTIONS = {
    "cfg1": "cfg1: synthetic 4-view 640x480 spherical sequence, multiband 5 bands",
    "cfg2": "cfg2: synthetic 8-view 1920x1080 sequence, linear blend + exposure gain (-e)",
    "cfg3": "cfg3: synthetic 12-view 4000x3000 spherical pano, multiband 6 bands",
    "cfg4": "cfg4: synthetic 36-view 4000x3000 full-sphere pano (~30k x 8k mosaic), multiband 5 bands, strip-sharded",
    "cfg5": "cfg5: batch of 64 synthetic 6-view 1920x1080 panoramas (8 distinct scenes x 8), multiband 5 bands, one pano per GPU at a time",
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="cfg4", choices=sorted(DESCRIPTIONS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink views (debug only; invalid as a bench number)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the short cfg1/cfg2/cfg3/cfg5 runs after cfg4")
    ap.add_argument("--cpu-budget-s", type=float, default=300.0,
                    help="time budget of the CPU arm: it picks the largest sample whose steps + warm-up fit")
    return ap.parse_args()


# ----------------------------------------------------------------------------
# CPU arm: the reference's own NumPy/OpenCV path on a bounded sample
# ----------------------------------------------------------------------------
def cpu_sample(wl, per_step_s=1e9):
    """A sub-panorama of the workload small enough for the CPU arm's budget: neighbouring views
    at full resolution, same blender — the largest of the candidates whose step fits
    ``per_step_s`` (the arm's time budget / the steps it was asked for; step times as measured on
    the GPU box's 16 host cores).  cfg4: 2 yaw columns x 2 pitch rows (vertical and horizontal
    seams and a four-image corner, ~30 s per step; the whole workload is ~20 min and needs
    ~85 GiB), else two neighbours of one pitch row (~12 s).  P/M of the sample is reported next to
    the workload's."""
    from pano360_b200 import synth  # noqa: F401
    if wl.name == "cfg4":
        candidates = [([5, 6, 17, 18], 31.0), ([17, 18], 12.0)]
    elif wl.name == "cfg3":
        candidates = [([2, 3, 8, 9], 36.0), ([2, 3], 14.0)]
    else:
        candidates = [(list(range(wl.n_views)), 0.0)]
    pick = next((p for p, cost in candidates if cost <= per_step_s), candidates[-1][0])
    sample = replace(wl, yaws=tuple(wl.yaws[i] for i in pick), pitches=tuple(wl.pitches[i] for i in pick))
    what = (f"{len(pick)} views of {wl.name} ({wl.width}x{wl.height}, views {pick}), "
            f"{wl.blend}" + (f" {wl.n_levels} bands" if wl.blend == "multiband" else "")
            + (" + gains" if wl.equalize else ""))
    return sample, what


def cpu_runner(wl):
    """Callable running the CPU path once on ``regions`` -> mosaic, and its kind."""
    from oracle import ref_harness
    if ref_harness.available():
        def run(regions):
            return ref_harness.ref_stitch(regions, wl.blend, equalize=wl.equalize, n_levels=wl.n_levels,
                                          max_resolution=wl.max_resolution)
        return run, "reference"
    from oracle import restate

    def run(regions):
        return restate.stitch(regions, wl.blend, wl.equalize, wl.n_levels, wl.max_resolution)
    return run, "port"


def time_cpu(wl, steps, warmup, budget_s):
    import cv2
    from pano360_b200 import synth
    sample, what = cpu_sample(wl, budget_s / max(steps + warmup, 1))
    regions = synth.make_views(sample)
    run, kind = cpu_runner(sample)
    times, mpix = [], None
    t_begin = time.perf_counter()
    done_warm = 0
    for _ in range(warmup):
        if time.perf_counter() - t_begin > budget_s * 0.4 and done_warm >= 1:
            break
        mpix = np.prod(run(regions).shape[:2]) / 1e6
        done_warm += 1
    for _ in range(steps):
        t0 = time.perf_counter()
        mpix = np.prod(run(regions).shape[:2]) / 1e6
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_begin > 1.25 * budget_s:      # (a slower box than the estimates: stay bounded)
            break
    sec = float(np.mean(times))
    from pano360_b200 import geometry as geo
    ratio = []
    for w in (sample, wl):          # P/M: patch-box pixels per mosaic pixel, the reference's own boxes
        cams = synth.make_views(w, only=set())
        pl = geo.plan_mosaic(cams, w.blend == "multiband", w.max_resolution)
        ratio.append(sum((x1 - x0) * (y1 - y0) for x0, y0, x1, y1 in pl.boxes) / (pl.shape[0] * pl.shape[1]))
    return {"value": mpix / sec, "unit": UNIT, "cores": int(cv2.getNumThreads()), "kind": kind,
            "sample": f"{what}; mosaic {mpix:.1f} Mpix in {sec:.2f} s/step, {len(times)} timed steps, {done_warm} warm-up; "
                      f"P/M of the sample {ratio[0]:.2f} vs {ratio[1]:.2f} for the workload "
                      f"(OpenCV pool {cv2.getNumThreads()} threads of {os.cpu_count()} cores, NumPy single-threaded)"}, \
        sec, len(times), done_warm


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from pano360_b200 import synth
    wl = synth.workload(args.workload, scale=args.scale)
    base, sec, n_timed, n_warm = time_cpu(wl, args.steps, args.warmup, args.cpu_budget_s)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT,
            "n_gpus": args.gpus, "steps": n_timed, "warmup": n_warm, "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": DESCRIPTIONS[wl.name], "sample": base["sample"]},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------
class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region: NVML from a
    background thread every 50 ms (cheaper than polling an nvidia-smi process,
    which measurably stalls kernel launches); nvidia-smi -lms as a fallback."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
               "sw_power_cap": 0x4}

    def __init__(self, index):
        self.index, self.proc, self.thread = index, None, None
        self.samples, self.reasons, self.max_mhz, self.stop = [], set(), None, False
        self.path = tempfile.mktemp(prefix="p360_clocks_", suffix=".csv")

    def _nvml_loop(self, nvml, handle):
        while not self.stop:
            try:
                self.samples.append(float(nvml.nvmlDeviceGetClockInfo(handle, nvml.NVML_CLOCK_SM)))
                mask = nvml.nvmlDeviceGetCurrentClocksThrottleReasons(handle)
                for name, bit in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def __enter__(self):
        try:
            import threading
            import pynvml as nvml
            nvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[self.index]) if visible and visible.replace(",", "").isdigit() else self.index
            handle = nvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(nvml.nvmlDeviceGetMaxClockInfo(handle, nvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._nvml_loop, args=(nvml, handle), daemon=True)
            self.thread.start()
            return self
        except Exception:
            self.thread = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None
        return self

    def __exit__(self, *exc):
        self.stop = True
        if self.thread is not None:
            self.thread.join(timeout=2)
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        out = {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0, "source": "nvml"}
        sm, reasons = list(self.samples), set(self.reasons)
        if self.thread is None:
            out["source"] = "nvidia-smi"
            try:
                rows = [r.split(",") for r in open(self.path).read().strip().splitlines() if r.strip()]
                os.unlink(self.path)
            except OSError:
                rows = []
            for r in rows:
                r = [c.strip() for c in r]
                try:
                    sm.append(float(r[1]))
                    out["sm_max_mhz"] = float(r[2])
                except (ValueError, IndexError):
                    continue
                for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        if sm:
            busy = sorted(sm)[len(sm) // 2:]          # upper half = samples under load
            out["sm_mhz"] = float(np.median(busy))
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


def model_bytes(wl, plan, n_src_bytes):
    """Algorithmic bytes of one step, stage-materialised model of SURVEY.md §8(d), on the
    reference's own patch boxes."""
    m = plan.shape[0] * plan.shape[1]
    p = sum((x1 - x0) * (y1 - y0) for x0, y0, x1, y1 in plan.boxes)
    if wl.blend == "multiband":
        lv = wl.n_levels
        return n_src_bytes + 17 * p + (4 * p + m) + 32 * p * (lv - 1) + 64 * p * lv + m * (40 * lv + 16), p, m
    if wl.blend == "linear":
        return n_src_bytes + 17 * p + 48 * p + 19 * m, p, m
    return n_src_bytes + 17 * p + 17 * p + 3 * m, p, m


def mosaic_checksum(mosaic):
    """CRC-32 of a mosaic (device tensor or host array), outside every timed region.  Equal values
    at N = 1, 2, 4, 8 — and for the device-timed and the end-to-end legs — say that the strips
    reassemble the single-GPU mosaic byte for byte."""
    import zlib
    host = mosaic.cpu().numpy() if hasattr(mosaic, "cpu") else np.ascontiguousarray(mosaic)
    return f"{zlib.crc32(host.reshape(-1).data):08x}"


def algorithmic_bytes(trace_name, wl, crop_px, src_bytes, m_px, h2d_px=0):
    """SURVEY.md §8(d) bytes of the stage(s) a traced kernel stands for, on the patch pixels the
    launch really covers (crop_px: after the seam split), None for implementation-only kernels."""
    lv = wl.n_levels
    table = {
        "K1t_warp_tiles": src_bytes + 17 * crop_px + 4 * crop_px + m_px,      # K1 + K2 fused: 3S + 17P + (4P + M)
        "K1_warp": src_bytes + 17 * crop_px + (4 * crop_px + m_px if wl.blend == "multiband" else 0),
        "K3a_pyramid_reduce": None, "K3_gauss_blur": 32 * crop_px * (lv - 1),
        "K4_multiband_collapse": 64 * crop_px * lv + m_px * (40 * lv + 16),
        "K6_linear_collapse": 48 * crop_px + 19 * m_px, "K7_paste_collapse": 17 * crop_px + 3 * m_px,
    }
    return table.get(trace_name)


def run_batch_of_panoramas(args, wl, comp, world, rank, n_panos=64, n_scenes=8, steps=None):
    """cfg5: independent panoramas, pano_id % N -> GPU, no communication
    (replicas only).  One step = the whole batch of 64."""
    import torch
    import torch.distributed as dist
    from pano360_b200 import _lib, geometry as geo, stitcher, synth
    scenes = []
    for seed in range(n_scenes):
        regs = synth.make_views(replace(wl, env_seed=seed, view_seed=7 + seed))
        for reg in regs:
            t = torch.empty(reg.img.shape, dtype=torch.uint8, pin_memory=True)
            t.numpy()[...] = reg.img
            reg.img, reg._pin = t.numpy(), t
        plan = geo.plan_mosaic(regs, True, wl.max_resolution)
        scenes.append((regs, plan, comp.upload(regs, pack=False),
                       torch.empty(plan.shape + (3,), dtype=torch.uint8, pin_memory=True)))
    mine = list(range(rank, n_panos, world))
    stitcher.MAX_RESOLUTION = wl.max_resolution

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step():
        for pano in mine:
            regs, plan, raw, _ = scenes[pano % n_scenes]
            comp.composite(regs, comp.pack_sources(raw), plan, wl.blend, wl.n_levels)

    def e2e_step():
        for pano in mine:
            regs, plan, _, out = scenes[pano % n_scenes]
            stitcher.stitch(regs, blender=stitcher.multiband_blend, n_levels=wl.n_levels, out=out.numpy())

    def timed(fn, steps, trace=False):
        barrier()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0, t0 = _lib.launch_count, time.perf_counter()
        comp.trace = [] if trace else None
        start.record(torch.cuda.current_stream())
        for _ in range(steps):
            fn()
        end.record(torch.cuda.current_stream())
        barrier()
        traced["events"], comp.trace = comp.trace, None
        host_ms, dev_ms, launched = (time.perf_counter() - t0) * 1e3, start.elapsed_time(end), _lib.launch_count - n0
        if world > 1:
            t = torch.tensor([dev_ms, host_ms], dtype=torch.float64, device=comp.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dev_ms, host_ms = t.tolist()
            n = torch.tensor([launched], dtype=torch.int64, device=comp.device)
            dist.all_reduce(n)
            launched = int(n.item())
        return dev_ms / steps, host_ms / steps, launched

    steps = args.steps if steps is None else steps
    traced = {}
    for _ in range(args.warmup):
        device_step()
    with ClockSampler(comp.device.index or 0) as clocks:
        ms, host_ms, launches = timed(device_step, steps)
    # per-kernel times of the batch and the roofline of its dominant kernel (this rank's panoramas;
    # SURVEY 8(d) bytes on the patch pixels after the seam split) from ONE more, traced pass: two
    # events around each of the batch's ~450 calls cost a tenth of a step this short, so the timed
    # steps above run untraced
    timed(device_step, 1, trace=True)
    trace_steps = 1
    per_kernel = {}
    for name, _, ev0, ev1 in traced.get("events") or []:
        agg = per_kernel.setdefault(name, [0.0, 0])
        agg[0] += ev0.elapsed_time(ev1)
        agg[1] += 1
    roofline = None
    if per_kernel:
        reach = comp.blur_reach(wl.blend, wl.n_levels)
        crop_px = src_b = m_px = 0
        for pano in mine:
            regs, plan, _, _ = scenes[pano % n_scenes]
            crops, _ = comp.plan_crops(regs, plan, split_dilate=2 * reach)
            crop_px += sum((c[3] - c[1]) * (c[4] - c[2]) for c in crops)
            src_b += sum(int(np.prod(r.img.shape)) for r in regs)
            m_px += plan.shape[0] * plan.shape[1]
        top = max(per_kernel, key=lambda k: per_kernel[k][0])
        nbytes = algorithmic_bytes(top, wl, crop_px, src_b, m_px) or 0
        peak = 6650.0
        try:
            peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", peak))
        except OSError:
            pass
        t_ms = max(per_kernel[top][0] / trace_steps, 1e-9)
        roofline = {"kernel": top, "bound": "hbm", "achieved": nbytes / t_ms / 1e6, "peak": peak, "unit": "GB/s",
                    "frac": nbytes / t_ms / 1e6 / peak, "traffic": None, "launch_ms": t_ms / (per_kernel[top][1] / trace_steps),
                    "algorithmic_bytes_per_launch": nbytes / (per_kernel[top][1] / trace_steps),
                    "share_of_step": t_ms / max(ms, 1e-9),
                    "bytes_model": "SURVEY 8(d) bytes of the stage the kernel stands for, summed over this rank's panoramas"}
    e2e_step()
    _, e2e_ms, _ = timed(e2e_step, steps)
    mpix = sum(np.prod(scenes[p % n_scenes][1].shape) for p in range(n_panos)) / 1e6
    if rank == 0:
        src_bytes = sum(int(np.prod(r.img.shape)) for p in mine for r in scenes[p % n_scenes][0])
        out_bytes = sum(int(np.prod(scenes[p % n_scenes][1].shape)) * 3 for p in mine)
        return {
            "metric": METRIC, "value": mpix / (ms / 1e3), "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": DESCRIPTIONS["cfg5"], "panoramas": n_panos, "distinct_scenes": n_scenes,
                       "views": wl.n_views, "view_size": [wl.width, wl.height], "n_levels": wl.n_levels,
                       "mosaic_mpix_total": mpix, "parallelism": "replicas: pano_id % n_gpus, no collective",
                       "l2": "every panorama streams ~1 GB; 8 distinct scenes cycle, nothing survives in the 126 MB L2"},
            "clocks": clocks.summary(),
            "e2e": {"value": mpix / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": src_bytes,
                    "d2h_bytes_per_step": out_bytes, "ms_per_step": e2e_ms, "api": "pano360_b200.stitcher.stitch"},
            "gpu_launches": launches, "roofline": roofline,
            "kernels": {k: {"ms_per_step": v[0] / trace_steps, "launches_per_step": v[1] / trace_steps}
                        for k, v in sorted(per_kernel.items())},
            "host_ms_per_step": host_ms}
    return None


class Bench:
    """One workload on this rank's GPU: device-timed leg, end-to-end legs, roofline."""

    def __init__(self, args, wl, comp, world, rank):
        import torch
        from pano360_b200 import geometry as geo, strips, synth
        self.args, self.wl, self.comp, self.world, self.rank = args, wl, comp, world, rank
        self.torch, self.strips, self.synth = torch, strips, synth
        kind, levels = wl.blend, wl.n_levels
        self.regions = synth.make_views(wl, only=set())     # cameras only: nothing rendered yet
        self.rendered, self.pageable = set(), [None] * wl.n_views
        self.plan = plan = geo.plan_mosaic_cached(self.regions, kind == "multiband", wl.max_resolution)
        self.halo = strips.blur_halo(kind, levels)
        self.tuned = None
        if world > 1:
            # measured-feedback strip cuts (untimed): the model's cuts first, then moved by the
            # device time every rank measures for its own strip
            def step(parts):
                self.place(parts)
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
                best = float("inf")
                for k in range(4):                        # (one warm run, then the best of three)
                    torch.cuda.synchronize()
                    ev[0].record(torch.cuda.current_stream())
                    if not strips.is_empty(parts[rank]):
                        rows, cols = strips.window_of(parts[rank], plan.shape)
                        comp.composite(self.regions, comp.pack_sources(self.raw), plan, kind, levels, rows=rows, cols=cols)
                    ev[1].record(torch.cuda.current_stream())
                    torch.cuda.synchronize()
                    if k:
                        best = min(best, ev[0].elapsed_time(ev[1]))
                return best
            self.tuned = {"rounds": []}
            self.place(strips.tune_partition(comp, self.regions, plan, kind, levels, step, log=self.tuned["rounds"]))
        else:
            self.place(strips.strip_cuts(plan, 1, kind, levels))
        self.src_bytes_all = sum(int(np.prod(r.img.shape)) for r in self.regions)
        self.out_pinned = torch.empty(plan.shape + (3,), dtype=torch.uint8, pin_memory=True) if rank == 0 and world == 1 else None

    def place(self, parts):
        """Adopt row cuts: render the views this rank's strip needs (those not rendered yet), keep
        pinned host copies (the e2e leg reads these every step) and pageable ones (what main() / a
        PKL user hands to stitch()), and make the rows the strip reads resident in HBM as uploaded."""
        torch, comp, wl, plan, rank = self.torch, self.comp, self.wl, self.plan, self.rank
        self.parts = parts
        part = parts[rank]
        rows, cols = self.strips.window_of(part, plan.shape)
        need = set(self.strips.images_for_part(plan, part, wl.blend, wl.n_levels))
        if wl.equalize:
            need = set(range(wl.n_views))
        missing = need - self.rendered
        if missing:
            fresh = self.synth.make_views(wl, only=missing)
            for i in missing:
                self.pageable[i] = fresh[i].img
                t = torch.empty(fresh[i].img.shape, dtype=torch.uint8, pin_memory=True)
                t.numpy()[...] = fresh[i].img
                self.regions[i]._pin, self.regions[i].img = t, t.numpy()     # numpy view of pinned memory
            self.rendered |= missing
        self.need = need
        self.rows_of = None
        if not wl.equalize and not self.strips.is_empty(part):
            # the part of every image this rank's strip reads (the seam plan's rectangles, else rows) —
            # what stitch() / stitch_strips() upload
            self.rows_of = comp.source_rects(self.regions, plan, wl.blend, wl.n_levels, rows=rows, cols=cols)
            if self.rows_of is None and self.world > 1:
                self.rows_of = comp.source_rows(self.regions, plan, wl.blend, wl.n_levels, rows=rows, cols=cols)
            if self.rows_of is not None:
                need &= set(self.rows_of)
        self.raw = comp.upload(self.regions, need=need, pack=False, rows_of=self.rows_of)   # u8 x 3 as uploaded, resident in HBM
        self.h2d_bytes = self.raw.bytes_up

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        self.torch.cuda.synchronize()

    def device_step(self):
        """value leg: the uploaded u8 x 3 images are resident in HBM; everything stitch() does
        with them is inside (RGBX packing, gains, warp, blend, gather of the strips)."""
        comp, wl = self.comp, self.wl
        src = comp.pack_sources(self.raw)
        if wl.equalize:
            from pano360_b200.stitcher import find_gains
            overlaps, sizes = self.strips.all_pair_statistics(comp, self.regions, src)
            comp.set_gains(src, find_gains(overlaps, sizes))
        return self.strips.composite_gather(comp, self.regions, src, self.plan, wl.blend, wl.n_levels, self.parts)

    def e2e_step(self, pageable=False):
        """e2e leg: public API, host buffers in, host mosaic out."""
        wl = self.wl
        regions = self.regions
        if pageable:
            from pano360_b200.camera import Image
            regions = [Image(img if img is not None else r.img, r.rot, r.intr) for r, img in zip(self.regions, self.pageable)]
        if self.world == 1:
            from pano360_b200 import stitcher
            stitcher.MAX_RESOLUTION = wl.max_resolution
            return stitcher.stitch(regions, blender=stitcher.BLENDERS[wl.blend], equalize=wl.equalize,
                                   n_levels=wl.n_levels, out=None if pageable else self.out_pinned.numpy())
        # (N ranks: rank 0 gets a view of the host buffer the ranks share — every rank downloads its own strip)
        return self.strips.stitch_strips(self.comp, regions, wl.blend, wl.n_levels, wl.equalize, wl.max_resolution)

    def timed(self, step_fn, n_steps, trace=False):
        import torch.distributed as dist
        from pano360_b200 import _lib
        torch, comp = self.torch, self.comp
        self.barrier()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = _lib.launch_count
        comp.trace = [] if trace else None
        t_host = time.perf_counter()
        start.record(torch.cuda.current_stream())
        for _ in range(n_steps):
            result = step_fn()
        end.record(torch.cuda.current_stream())
        enqueue_s = time.perf_counter() - t_host          # host time to issue the steps (no sync)
        self.barrier()
        host_s = time.perf_counter() - t_host
        ms = start.elapsed_time(end)
        launched = _lib.launch_count - launches0
        trace_out, comp.trace = comp.trace, None
        if self.world > 1:
            t = torch.tensor([ms, host_s * 1e3], dtype=torch.float64, device=comp.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, host_ms = t.tolist()
            host_s = host_ms / 1e3
            n = torch.tensor([launched], dtype=torch.int64, device=comp.device)
            dist.all_reduce(n, op=dist.ReduceOp.SUM)
            launched = int(n.item())
        return {"ms": ms, "host_s": host_s, "launches": launched, "trace": trace_out, "result": result,
                "enqueue_ms": enqueue_s / n_steps * 1e3}

    def run(self, steps, warmup, local_gpu, full=True):
        """-> the result record (rank 0) or None.  ``full``: also the roofline / per-kernel tables,
        the pageable e2e leg and the block statistics (the main workload)."""
        import torch.distributed as dist
        torch, comp, wl, plan, world, rank = self.torch, self.comp, self.wl, self.plan, self.world, self.rank
        for _ in range(warmup):
            self.device_step()
        with ClockSampler(local_gpu) as clocks:
            dev = self.timed(self.device_step, steps, trace=True)
        clock_summary = clocks.summary()
        checksum = mosaic_checksum(dev["result"]) if rank == 0 and dev["result"] is not None else None
        stats = self.block_stats() if full and world == 1 else None
        for _ in range(min(warmup, 2)):
            self.e2e_step()
        e2e = self.timed(self.e2e_step, steps)
        e2e_crc = mosaic_checksum(e2e["result"]) if rank == 0 and e2e["result"] is not None else None
        self.h2d_bytes = int(comp.last_upload_bytes)       # what the last e2e step's uploads really copied ...
        if world > 1:                                      # ... summed over the ranks
            t = torch.tensor([self.h2d_bytes], dtype=torch.int64, device=comp.device)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            self.h2d_bytes = int(t.item())
        phases = None
        if world > 1:                      # one more call, host-side phase times of rank 0
            comp.phases = []
            self.barrier()
            self.e2e_step()
            self.barrier()
            phases, comp.phases = (comp.phases[-1] if comp.phases else None), None
        e2e_page = None
        if full and world == 1:
            self.e2e_step(pageable=True)
            e2e_page = self.timed(lambda: self.e2e_step(pageable=True), max(1, min(steps, 3)))

        mpix = plan.shape[0] * plan.shape[1] / 1e6
        ms_per_step = max(dev["ms"], 1e-9) / steps
        e2e_ms = e2e["host_s"] / steps * 1e3
        crop_px = self.crop_pixels()
        # ---- per-kernel table + roofline of the dominant kernel, from CUDA events in the timed steps
        per_kernel = {}
        for name, _, ev0, ev1 in dev["trace"] or []:
            agg = per_kernel.setdefault(name, [0.0, 0])
            agg[0] += ev0.elapsed_time(ev1)
            agg[1] += 1
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_kind = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        roofline, shares = None, {}
        m_px = plan.shape[0] * plan.shape[1]
        if per_kernel:
            traced_ms = sum(v[0] for v in per_kernel.values())
            for k, v in sorted(per_kernel.items()):
                nbytes = algorithmic_bytes(k, wl, crop_px, self.src_bytes_all, m_px)
                shares[k] = {"ms_per_step": v[0] / steps, "launches_per_step": v[1] / steps,
                             "share_of_traced": v[0] / max(traced_ms, 1e-9),
                             # SURVEY 8(d) stage-materialised bytes / time: the pipeline decimates and fuses, so for
                             # the blur and collapse stages this exceeds the HBM peak and is no roofline
                             "model_GBps": None if nbytes is None else nbytes * steps / max(v[0], 1e-9) / 1e6}
            if stats is not None:          # bytes of the blocks the list kernels really ran
                for k, key in (("K3a_pyramid_reduce", "reduce_bytes"), ("K3_gauss_blur", "blur_bytes")):
                    if k in shares:
                        shares[k]["run_GBps"] = stats[key] / (shares[k]["ms_per_step"] / 1e3) / 1e9
            top = max(per_kernel, key=lambda k: per_kernel[k][0])
            t_ms, count = per_kernel[top]
            nbytes = algorithmic_bytes(top, wl, crop_px, self.src_bytes_all, m_px) or 0
            per_launch = nbytes * steps / count
            t_ms = max(t_ms, 1e-9)
            achieved = nbytes * steps / t_ms / 1e6                     # GB/s
            traffic = None
            try:           # ncu-measured DRAM bytes per launch of this kernel (1 GPU, same workload)
                table = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
                if world == 1 and self.args.scale == 1.0:
                    traffic = table.get(wl.name, {}).get(top)
            except OSError:
                pass
            roofline = {"kernel": top, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                        "frac": achieved / peak, "traffic": traffic, "peak_source": peak_kind,
                        "launch_ms": t_ms / count, "algorithmic_bytes_per_launch": per_launch,
                        "bytes_model": "SURVEY 8(d): 3S + 17P [K1] + 4P + M [K2] for the fused tile warp, P = patch pixels "
                                       "after the seam split (%.1f Mpix), S = %.1f Mpix, M = %.1f Mpix"
                                       % (crop_px / 1e6, self.src_bytes_all / 3e6, m_px / 1e6),
                        "share_of_step": t_ms / max(dev["ms"], 1e-9)}
            if "K1p_pack_rgbx" in per_kernel and top.startswith("K1"):     # the same bytes over warp + RGBX packing
                both = t_ms + per_kernel["K1p_pack_rgbx"][0]
                roofline["frac_with_pack"] = nbytes * steps / both / 1e6 / peak
        my_kernel_ms = sum(v[0] for v in per_kernel.values()) / steps if per_kernel else 0.0
        per_rank = [my_kernel_ms]
        if world > 1:
            t = torch.zeros(world, dtype=torch.float64, device=comp.device)
            t[rank] = my_kernel_ms
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            per_rank = [round(v, 3) for v in t.tolist()]
        total_bytes, p_px, _ = model_bytes(wl, plan, self.src_bytes_all)
        if rank != 0:
            return None
        line = {
            "metric": METRIC, "value": mpix / (ms_per_step / 1e3), "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": DESCRIPTIONS[wl.name] + (f" (DEBUG scale 1/{self.args.scale})" if self.args.scale != 1 else ""),
                       "views": wl.n_views, "view_size": [wl.width, wl.height], "blend": wl.blend,
                       "n_levels": wl.n_levels if wl.blend == "multiband" else None, "equalize": wl.equalize,
                       "mosaic": list(plan.shape), "mosaic_mpix": mpix, "patch_mpix_reference_boxes": p_px / 1e6,
                       "patch_mpix_after_seam_split": crop_px / 1e6,
                       "strips": [list(p) for p in self.parts], "strip_axis": "cols" if len(self.parts[0]) == 4 else "rows",
                       "strip_cuts": self.tuned or "model", "halo_rows": self.halo,
                       "halo_cols": self.strips.col_halo(wl.blend, wl.n_levels),
                       "timed_region": "sources resident in HBM as uploaded (u8 x 3; of each image the rectangle the seam plan "
                                       "reads, as stitch() uploads it); RGBX packing, gains, seam plan, warp, blend and the "
                                       "gather of the strips are all inside",
                       "l2": "no flush: each step streams >> 126 MB (model bytes %.1f GB) so nothing survives in L2 between steps"
                             % (total_bytes / 1e9)},
            "clocks": clock_summary,
            "e2e": {"value": mpix / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": self.h2d_bytes,
                    "d2h_bytes_per_step": int(np.prod(plan.shape)) * 3, "ms_per_step": e2e_ms,
                    "api": "pano360_b200.stitcher.stitch" if world == 1 else "pano360_b200.strips.stitch_strips",
                    "host_buffers": "pinned", "mosaic_checksum": e2e_crc, "rank0_phases_ms": phases},
            "gpu_launches": dev["launches"],
            "mosaic_checksum": checksum,
            "roofline": roofline,
            "kernels": shares,
            "per_rank_kernel_ms": per_rank,
            "host_ms_per_step": dev["host_s"] / steps * 1e3,
            "host_enqueue_ms_per_step": dev["enqueue_ms"],
        }
        if full:
            line["pipeline"] = {"model_bytes_per_step": total_bytes, "GBps": total_bytes / (ms_per_step / 1e3) / 1e9,
                                "note": "whole-step SURVEY 8(d) model bytes / time: the pipeline decimates and fuses, so "
                                        "this exceeds the HBM peak and is not a roofline"}
        if e2e_page is not None:
            n = max(1, min(steps, 3))
            line["e2e_pageable"] = {"value": mpix / (e2e_page["host_s"] / n), "unit": UNIT,
                                    "ms_per_step": e2e_page["host_s"] / n * 1e3,
                                    "note": "same call with plain (pageable) NumPy inputs and a returned array, as main() makes it"}
        if stats is not None:
            line["blocks_run"] = stats
        return line

    def crop_pixels(self):
        """Patch pixels this rank's composite covers (after the seam split; its row window only)."""
        comp, wl = self.comp, self.wl
        part = self.parts[self.rank]
        if self.strips.is_empty(part):
            return 0
        rows, cols = (None, None) if self.world == 1 else self.strips.window_of(part, self.plan.shape)
        from pano360_b200 import geometry as geo
        crops = comp._window_geometry(self.regions, self.plan, wl.blend, wl.n_levels, geo.SphProj, rows, cols)[0]
        return int(sum((c[3] - c[1]) * (c[4] - c[2]) for c in crops))

    def block_stats(self):
        """Blocks the list kernels really ran in the last composite and the bytes they moved
        (reduce: 32 x 32 px of RGBA + keys in, d2 + d4 out; blur: staged cells in, cells out)."""
        keep = self.comp._keep.get("bands")
        if not keep or keep[3] is None:
            return None
        self.torch.cuda.synchronize()
        bits = keep[4][0]
        n_red, n_h, n_v = [int(v) for v in bits[1:4].tolist()]
        h_rows = int(keep[3]["h_rows"][0])
        seg, rows_h = (64, 16) if h_rows == 4 else (256, 4)
        ks = 25                                   # widest coarse tap set (level L-2 on the f = 4 grid)
        red_b = n_red * (1024 * 24 + 256 * 16 + 64 * 16)
        h_b = n_h * ((seg + ks - 1) * rows_h * 16 + seg * rows_h * 16)
        v_b = n_v * (32 * (64 + ks - 1) * 16 + 32 * 64 * 16)
        return {"reduce_blocks": n_red, "blur_h_blocks": n_h, "blur_v_blocks": n_v,
                "reduce_bytes": red_b, "blur_bytes": h_b + v_b}


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from pano360_b200 import synth
    from pano360_b200.compositor import Compositor

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comp = Compositor(torch.device("cuda", local))

    wl = synth.workload(args.workload, scale=args.scale)
    if wl.name == "cfg5":
        line = run_batch_of_panoramas(args, wl, comp, world, rank)
    else:
        bench = Bench(args, wl, comp, world, rank)
        line = bench.run(args.steps, args.warmup, local)
        bench = None
        comp.release(everything=True)
        torch.cuda.empty_cache()
    # the other BASELINE configs, short runs, so that they are in the driver's record too
    others = {}
    if args.workload == "cfg4" and args.scale == 1.0 and not args.no_other_configs:
        for name in ("cfg1", "cfg2", "cfg3"):
            if world > 1:
                break                      # single-GPU configurations
            sub = Bench(args, synth.workload(name), comp, 1, 0).run(max(3, args.steps), 3, local, full=False)
            others[name] = {k: sub[k] for k in ("value", "unit", "ms_per_step", "e2e", "gpu_launches", "roofline",
                                                "mosaic_checksum", "host_enqueue_ms_per_step")}
            others[name]["config"] = sub["config"]["workload"]
            comp.release(everything=True)
            torch.cuda.empty_cache()
        sub = run_batch_of_panoramas(args, synth.workload("cfg5"), comp, world, rank, steps=max(2, min(args.steps, 3)))
        if sub is not None:
            others["cfg5"] = {k: sub[k] for k in ("value", "unit", "ms_per_step", "e2e", "gpu_launches", "n_gpus", "roofline",
                                                  "kernels", "host_ms_per_step")}
            others["cfg5"]["config"] = sub["config"]["workload"] + "; replicas: pano_id % n_gpus"
    if rank == 0 and line is not None:
        if others:
            line["other_configs"] = others
        if world == 1 and not args.no_cpu_baseline and wl.name != "cfg5":
            base, _, _, _ = time_cpu(wl, 1, 0, args.cpu_budget_s / 4)     # (one step of the larger sample)
            line["cpu_baseline"] = base
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
