import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "active-orchard-slam_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
from aos_gpu import lib, synth
spec = synth.config("C3", n_points=50_000_000)
pts = synth.make_orchard_torch(spec, "cuda")
ctx = lib.Context(0)
ctx.set_profiling(True)
P = lib.SeedParams(grid_resolution=spec.grid_resolution, inflation_radius=spec.inflation_radius, polygon=spec.polygon)
ctx.seed_stage(P, pts)
s = ctx.seed_summary(); w, h = s.info.width, s.info.height
for gid, nm in ((lib.GRID_RAW, "raw"), (lib.GRID_SKELETON_FRAMED, "skeleton")):
    ptr, pitch = ctx.grid_device_bits(gid)
    near = torch.empty((h, w), dtype=torch.int32, device="cuda"); d2 = torch.empty((h, w), dtype=torch.int32, device="cuda")
    for it in range(3):
        torch.cuda.synchronize(); t = time.perf_counter()
        assert ctx.L.aos_edt_bits(ctx.h, ptr, w, h, near.data_ptr(), d2.data_ptr()) == 0
        torch.cuda.synchronize(); print(nm, "edt ms", round((time.perf_counter() - t) * 1e3, 2), [(n, round(m, 2)) for n, m in ctx.stage_times()])
