import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "active-orchard-slam_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
from aos_gpu import lib, synth
name = sys.argv[1] if len(sys.argv) > 1 else "C2"
npts = int(sys.argv[2]) if len(sys.argv) > 2 else None
spec = synth.config(name, n_points=npts)
t = time.time(); pts = synth.make_orchard_torch(spec, "cuda"); torch.cuda.synchronize(); print("gen", time.time() - t, pts.shape)
ctx = lib.Context(0); ctx.set_profiling(True)
P = lib.SeedParams(grid_resolution=spec.grid_resolution, inflation_radius=spec.inflation_radius, polygon=spec.polygon)
for it in range(3):
    t = time.time(); s = ctx.seed_stage(P, pts); dt = time.time() - t
    print(f"iter {it}: {dt*1e3:.2f} ms  grid {s.info.width}x{s.info.height} kept {s.n_points_in} clusters {s.n_clusters} rows {s.n_rows} thin launches {s.thinning_launches}")
    print("   ", "  ".join(f"{n}={ms:.3f}" for n, ms in ctx.stage_times()))
cl = ctx.clusters()
print("classes: bbox words", ); print("flags", np.bincount(cl["reserved"], minlength=8), "max size", cl["size"].max() if len(cl) else 0)
