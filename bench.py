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
    ap.add_argument("--cpu-budget-s", type=float, default=240.0)
    return ap.parse_args()


# ----------------------------------------------------------------------------
# CPU arm: the reference's own NumPy/OpenCV path on a bounded sample
# ----------------------------------------------------------------------------
def cpu_sample(wl):
    """A sub-panorama of the workload small enough for ~10-25 s of CPU work:
    adjacent views of the same ring at full resolution, same blender."""
    from pano360_b200 import synth  # noqa: F401
    if wl.name == "cfg4":
        pick = [17, 18]          # two neighbouring views of the pitch-0 row
    elif wl.name == "cfg3":
        pick = [2, 3]
    else:
        pick = list(range(wl.n_views))
    sample = replace(wl, yaws=tuple(wl.yaws[i] for i in pick), pitches=tuple(wl.pitches[i] for i in pick))
    what = (f"{len(pick)} adjacent views of {wl.name} ({wl.width}x{wl.height}, views {pick}), "
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
    sample, what = cpu_sample(wl)
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
        if time.perf_counter() - t_begin > budget_s:
            break
    sec = float(np.mean(times))
    return {"value": mpix / sec, "unit": UNIT, "cores": int(cv2.getNumThreads()), "kind": kind,
            "sample": f"{what}; mosaic {mpix:.1f} Mpix in {sec:.2f} s/step, {len(times)} timed steps "
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
    """Algorithmic bytes of one step, stage-materialised model of SURVEY.md §8(d)."""
    m = plan.shape[0] * plan.shape[1]
    p = sum((x1 - x0) * (y1 - y0) for x0, y0, x1, y1 in plan.boxes)
    if wl.blend == "multiband":
        lv = wl.n_levels
        return n_src_bytes + 17 * p + (4 * p + m) + 32 * p * (lv - 1) + 64 * p * lv + m * (40 * lv + 16), p, m
    if wl.blend == "linear":
        return n_src_bytes + 17 * p + 48 * p + 19 * m, p, m
    return n_src_bytes + 17 * p + 17 * p + 3 * m, p, m


def run_batch_of_panoramas(args, wl, comp, world, rank, n_panos=64, n_scenes=8):
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
        scenes.append((regs, plan, comp.upload(regs),
                       torch.empty(plan.shape + (3,), dtype=torch.uint8, pin_memory=True)))
    mine = list(range(rank, n_panos, world))
    stitcher.MAX_RESOLUTION = wl.max_resolution

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step():
        for pano in mine:
            regs, plan, src, _ = scenes[pano % n_scenes]
            comp.composite(regs, src, plan, wl.blend, wl.n_levels)

    def e2e_step():
        for pano in mine:
            regs, plan, _, out = scenes[pano % n_scenes]
            stitcher.stitch(regs, blender=stitcher.multiband_blend, n_levels=wl.n_levels, out=out.numpy())

    def timed(fn, steps):
        barrier()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0, t0 = _lib.launch_count, time.perf_counter()
        start.record(torch.cuda.current_stream())
        for _ in range(steps):
            fn()
        end.record(torch.cuda.current_stream())
        barrier()
        host_ms, dev_ms, launched = (time.perf_counter() - t0) * 1e3, start.elapsed_time(end), _lib.launch_count - n0
        if world > 1:
            t = torch.tensor([dev_ms, host_ms], dtype=torch.float64, device=comp.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dev_ms, host_ms = t.tolist()
            n = torch.tensor([launched], dtype=torch.int64, device=comp.device)
            dist.all_reduce(n)
            launched = int(n.item())
        return dev_ms / steps, host_ms / steps, launched

    for _ in range(args.warmup):
        device_step()
    with ClockSampler(comp.device.index or 0) as clocks:
        ms, _, launches = timed(device_step, args.steps)
    e2e_step()
    _, e2e_ms, _ = timed(e2e_step, args.steps)
    mpix = sum(np.prod(scenes[p % n_scenes][1].shape) for p in range(n_panos)) / 1e6
    if rank == 0:
        src_bytes = sum(int(np.prod(r.img.shape)) for p in mine for r in scenes[p % n_scenes][0])
        out_bytes = sum(int(np.prod(scenes[p % n_scenes][1].shape)) * 3 for p in mine)
        print(json.dumps({
            "metric": METRIC, "value": mpix / (ms / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": DESCRIPTIONS["cfg5"], "panoramas": n_panos, "distinct_scenes": n_scenes,
                       "views": wl.n_views, "view_size": [wl.width, wl.height], "n_levels": wl.n_levels,
                       "mosaic_mpix_total": mpix, "parallelism": "replicas: pano_id % n_gpus, no collective",
                       "l2": "every panorama streams ~1 GB; 8 distinct scenes cycle, nothing survives in the 126 MB L2"},
            "clocks": clocks.summary(),
            "e2e": {"value": mpix / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": src_bytes,
                    "d2h_bytes_per_step": out_bytes, "ms_per_step": e2e_ms, "api": "pano360_b200.stitcher.stitch"},
            "gpu_launches": launches, "roofline": None}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from pano360_b200 import _lib, geometry as geo, strips, synth
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
        return run_batch_of_panoramas(args, wl, comp, world, rank)
    kind, levels = wl.blend, wl.n_levels
    cameras = synth.make_views(wl, only=set())          # cameras only: nothing rendered yet
    plan = geo.plan_mosaic(cameras, kind == "multiband", wl.max_resolution)
    parts = strips.partition_rows(plan, world, kind, levels)
    halo = strips.blur_halo(kind, levels)
    rows = parts[rank]
    need = set(strips.images_for_rows(plan, rows, halo)) if rows[1] > rows[0] else set()
    if wl.equalize:
        need = set(range(len(cameras)))
    regions = synth.make_views(wl, only=need)           # each rank renders only the views its strip needs

    # pinned host copies of the inputs (the e2e leg reads these every step)
    pinned = []
    for i, reg in enumerate(regions):
        if i not in need:
            pinned.append(None)
            continue
        t = torch.empty(reg.img.shape, dtype=torch.uint8, pin_memory=True)
        t.numpy()[...] = reg.img
        reg.img = t.numpy()                      # numpy view of pinned memory
        pinned.append(t)
    src = comp.upload(regions, need=need)
    src_bytes_all = sum(int(np.prod(r.img.shape)) for r in regions)
    h2d_bytes = sum(int(np.prod(regions[i].img.shape)) for i in need)
    out_pinned = torch.empty(plan.shape + (3,), dtype=torch.uint8, pin_memory=True) if rank == 0 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step():
        """value leg: sources resident in HBM."""
        if wl.equalize:
            overlaps, sizes = strips.all_pair_statistics(comp, regions, src)
            from pano360_b200.stitcher import find_gains
            comp.set_gains(src, find_gains(overlaps, sizes))
        return strips.composite_gather(comp, regions, src, plan, kind, levels, parts)

    def e2e_step():
        """e2e leg: public API, host buffers in, host mosaic out."""
        if world == 1:
            from pano360_b200 import stitcher
            stitcher.MAX_RESOLUTION = wl.max_resolution
            return stitcher.stitch(regions, blender=stitcher.BLENDERS[kind], equalize=wl.equalize,
                                   n_levels=levels, out=out_pinned.numpy())
        return strips.stitch_strips(comp, regions, kind, levels, wl.equalize, wl.max_resolution,
                                    out=None if out_pinned is None else out_pinned.numpy())

    def timed(step_fn, n_steps, trace=False):
        barrier()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = _lib.launch_count
        comp.trace = [] if trace else None
        t_host = time.perf_counter()
        start.record(torch.cuda.current_stream())
        for _ in range(n_steps):
            result = step_fn()
        end.record(torch.cuda.current_stream())
        enqueue_s = time.perf_counter() - t_host          # host time to issue the steps (no sync)
        barrier()
        host_s = time.perf_counter() - t_host
        ms = start.elapsed_time(end)
        launched = _lib.launch_count - launches0
        trace_out, comp.trace = comp.trace, None
        if world > 1:
            t = torch.tensor([ms, host_s * 1e3], dtype=torch.float64, device=comp.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, host_ms = t.tolist()
            host_s = host_ms / 1e3
            n = torch.tensor([launched], dtype=torch.int64, device=comp.device)
            dist.all_reduce(n, op=dist.ReduceOp.SUM)
            launched = int(n.item())
        timed.enqueue_ms = enqueue_s / n_steps * 1e3
        return ms, host_s, launched, trace_out, result

    for _ in range(args.warmup):
        device_step()
    with ClockSampler(local) as clocks:
        ms, host_s, launches, trace, mosaic = timed(device_step, args.steps, trace=True)
    clock_summary = clocks.summary()
    enqueue_ms = timed.enqueue_ms
    # e2e: host wall clock (includes the blocking D2H), max over ranks
    for _ in range(min(args.warmup, 2)):
        e2e_step()
    _, e2e_host_s, _, _, _ = timed(e2e_step, args.steps)

    mpix = plan.shape[0] * plan.shape[1] / 1e6
    ms_per_step = ms / args.steps
    value = mpix / (ms_per_step / 1e3)
    e2e_value = mpix / (e2e_host_s / args.steps)

    # ---- roofline of the dominant kernel, from CUDA events in the timed region
    per_kernel = {}
    for name, nbytes, ev0, ev1 in trace or []:
        agg = per_kernel.setdefault(name, [0.0, 0, 0])
        agg[0] += ev0.elapsed_time(ev1)
        agg[1] += nbytes
        agg[2] += 1
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_kind = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    roofline, shares = None, {}
    if per_kernel:
        traced_ms = sum(v[0] for v in per_kernel.values())
        shares = {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[2] / args.steps,
                      "share_of_traced": v[0] / traced_ms, "GBps": v[1] / v[0] / 1e6}
                  for k, v in sorted(per_kernel.items())}
        top = max(per_kernel, key=lambda k: per_kernel[k][0])
        t_ms, nbytes, count = per_kernel[top]
        achieved = nbytes / t_ms / 1e6                     # GB/s
        traffic = None
        try:           # ncu-measured DRAM bytes per launch of this kernel (1 GPU, same workload)
            table = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            if world == 1 and args.scale == 1.0:
                traffic = table.get(wl.name, {}).get(top)
        except OSError:
            pass
        roofline = {"kernel": top, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "peak_source": peak_kind,
                    "launch_ms": t_ms / count, "algorithmic_bytes_per_launch": nbytes / count,
                    "share_of_step": t_ms / ms}
    # per-rank load (strip balance): traced kernel time of every rank, gathered on rank 0
    my_kernel_ms = sum(v[0] for v in per_kernel.values()) / args.steps if per_kernel else 0.0
    per_rank = [my_kernel_ms]
    if world > 1:
        t = torch.zeros(world, dtype=torch.float64, device=comp.device)
        t[rank] = my_kernel_ms
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        per_rank = [round(v, 3) for v in t.tolist()]
    total_bytes, p_px, m_px = model_bytes(wl, plan, src_bytes_all)
    pipeline = {"model_bytes_per_step": total_bytes, "GBps": total_bytes / (ms_per_step / 1e3) / 1e9,
                "frac_of_hbm_peak": total_bytes / (ms_per_step / 1e3) / 1e9 / peak}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": DESCRIPTIONS[wl.name] + (f" (DEBUG scale 1/{args.scale})" if args.scale != 1 else ""),
                   "views": wl.n_views, "view_size": [wl.width, wl.height], "blend": kind,
                   "n_levels": levels if kind == "multiband" else None, "equalize": wl.equalize,
                   "mosaic": list(plan.shape), "mosaic_mpix": mpix, "patch_mpix": p_px / 1e6,
                   "strips": [list(p) for p in parts], "halo_rows": halo,
                   "l2": "no flush: each step streams >> 126 MB (model bytes %.1f GB) so nothing survives in L2 between steps"
                         % (total_bytes / 1e9)},
        "clocks": clock_summary,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": int(np.prod(plan.shape)) * 3, "ms_per_step": e2e_host_s / args.steps * 1e3,
                "api": "pano360_b200.stitcher.stitch" if world == 1 else "pano360_b200.strips.stitch_strips"},
        "gpu_launches": launches,
        "roofline": roofline,
        "pipeline": pipeline,
        "kernels": shares,
        "per_rank_kernel_ms": per_rank,
        "host_ms_per_step": host_s / args.steps * 1e3,
        "host_enqueue_ms_per_step": enqueue_ms,
    }
    if world == 1 and not args.no_cpu_baseline:
        base, _, _, _ = time_cpu(wl, 1, 1, args.cpu_budget_s / 4)
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
