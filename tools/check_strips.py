"""torchrun check: strip-sharded composite == single-GPU composite, byte for byte.
   python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/check_strips.py"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pano360_b200 import geometry as geo, strips, synth
from pano360_b200.compositor import Compositor

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
comp = Compositor(torch.device("cuda", local))
ok = True
for name, scale, noise in (("cfg1", 1.0, 20.0), ("cfg2", 2.0, 0.0), ("cfg4", 8.0, 5.0), ("cfg3", 4.0, 5.0)):
    wl = synth.workload(name, scale=scale)
    regs = synth.make_views(wl, noise=noise)
    got = strips.stitch_strips(comp, regs, wl.blend, wl.n_levels, wl.equalize, wl.max_resolution)
    if rank == 0:
        from pano360_b200 import stitcher
        stitcher.MAX_RESOLUTION = wl.max_resolution
        want = stitcher.stitch(regs, blender=stitcher.BLENDERS[wl.blend], equalize=wl.equalize, n_levels=wl.n_levels)
        same = got.shape == want.shape and np.array_equal(got, want)
        d = np.abs(got.astype(int) - want.astype(int)).max() if got.shape == want.shape else -1
        print(f"{name} x{world} ranks: mosaic {got.shape} identical={same} max|d|={d}", flush=True)
        ok &= bool(same) or (wl.equalize and d <= 1)
    # the device-resident result (strips stored into rank 0's mosaic by the kernels / pushed by DMA)
    for fused in (True, False):
        strips.FUSED_GATHER = fused
        dev = strips.stitch_strips(comp, regs, wl.blend, wl.n_levels, wl.equalize, wl.max_resolution, to_host=False)
        if rank == 0:
            same = np.array_equal(dev.cpu().numpy(), got)
            print(f"{name} x{world} ranks, device gather fused={fused}: identical={same}", flush=True)
            ok &= bool(same)
    strips.FUSED_GATHER = False
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
