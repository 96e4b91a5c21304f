import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "active-orchard-slam_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
from aos_gpu import lib, synth
spec = synth.config("C3")
pts = synth.make_orchard_torch(spec, "cuda")
ctx = lib.Context(0)
P = lib.SeedParams(grid_resolution=spec.grid_resolution, inflation_radius=spec.inflation_radius, polygon=spec.polygon)
ctx.seed_stage(P, pts)
t = time.time(); seeds, counts, rows = ctx.select_seeds(); print("select_seeds", time.time() - t, counts, len(rows))
os.environ["AOS_SEEDS_ON_HOST"] = "1"
np.save(os.path.join(ROOT, "gpurun_out", "c3_seeds.npy"), seeds.astype(np.float64))
np.save(os.path.join(ROOT, "gpurun_out", "c3_rows.npy"), rows)
