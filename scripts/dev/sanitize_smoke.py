"""Small end-to-end pass over every kernel family for compute-sanitizer (memcheck): TINY map through
aos_map_to_graph, ROR, EDT + clearance, band API with world 1, generic point layout."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "active-orchard-slam_b200")); sys.path.insert(0, ROOT)
import numpy as np
from aos_gpu import lib, synth, bands
spec = synth.config("TINY", seed=7)
pts = synth.make_orchard(spec)
P = lib.SeedParams(grid_resolution=spec.grid_resolution, inflation_radius=spec.inflation_radius, polygon=spec.polygon)
ctx = lib.Context(0)
dp = ctx.radius_outlier_removal(pts, 0.2, 2)
print("ror kept", dp.shape[0], "of", len(pts))
ctx.set_clearance(True)
info = ctx.map_to_graph(P, pts, fetch=True)
print("graph", info["graph"])
g = ctx.graph(); print("clearance min/max", float(g["edge_clearances"].min()), float(g["edge_clearances"].max()))
rec = np.zeros((len(pts), 8), np.float32); rec[:, 1], rec[:, 2], rec[:, 5] = pts[:, 0], pts[:, 1], pts[:, 2]
ctx.seed_stage(P, rec, point_step=32, offsets=(4, 8, 20))
ctx.labels()
gi = lib.grid_geometry(P)
band = bands.band_for(gi.height, 1, 0, ctx.band_halo_rows(P))
ctx.band_raster(P, band, pts)
n = 0
while ctx.band_thin_launch():
    n += 1
print("band thin launches", n + 1)
P2 = lib.SeedParams(grid_resolution=0.02, inflation_radius=1.6, polygon=spec.polygon)   # R = 80: EDT inflation
ctx.seed_stage(P2, pts)
print("ok")
