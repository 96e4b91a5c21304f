import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "active-orchard-slam_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
from aos_gpu import lib, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000_000
spec = synth.config("C3", n_points=n)
pts = synth.make_orchard_torch(spec, "cuda")
ctx = lib.Context(0); ctx.set_profiling(True)
for it in range(3):
    torch.cuda.synchronize(); t = time.perf_counter()
    dp = ctx.radius_outlier_removal(pts, 0.2, 2)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(f"ror {n} pts -> {dp.shape[0]} kept in {dt*1e3:.1f} ms", [(a, round(b, 2)) for a, b in ctx.stage_times()])
