"""cProfile of the host side of one composite (run on the GPU box)."""
import sys, os, cProfile, pstats, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pano360_b200 import synth, geometry as geo, strips
from pano360_b200.compositor import Compositor
wl = synth.workload("cfg4")
cams = synth.make_views(wl, only=set())
plan = geo.plan_mosaic(cams, True, 1e9)
parts = strips.partition_rows(plan, 8, "multiband", 5)
rows = parts[3]
comp = Compositor()
need = set(strips.images_for_rows(plan, rows, strips.blur_halo("multiband", 5)))
regs = synth.make_views(wl, only=need)
src = comp.upload(regs, need=need)
for _ in range(3): comp.composite(regs, src, plan, "multiband", 5, rows=rows)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10): comp.composite(regs, src, plan, "multiband", 5, rows=rows)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"strip rows {rows}: enqueue {(t1-t0)/10*1e3:.2f} ms/step, total {(t2-t0)/10*1e3:.2f} ms/step, images {len(need)}")
pr = cProfile.Profile(); pr.enable()
for _ in range(10): comp.composite(regs, src, plan, "multiband", 5, rows=rows)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
