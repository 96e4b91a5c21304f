"""Stage times of one config-3 map (library-side CUDA events):  python scripts/dev/time_map.py [C3] [device|replay]"""
import sys, os, time
sys.path[:0] = [os.getcwd(), os.path.join(os.getcwd(), "active-orchard-slam_b200")]
import torch
from aos_gpu import lib, synth
name = sys.argv[1] if len(sys.argv) > 1 else "C3"
mode = sys.argv[2] if len(sys.argv) > 2 else "device"
spec = synth.config(name, seed=0)
pts = synth.make_orchard_torch(spec, "cuda")
p = lib.SeedParams(grid_resolution=spec.grid_resolution, inflation_radius=spec.inflation_radius, polygon=spec.polygon)
ctx = lib.Context(0)
ctx.set_voronoi_mode(mode == "device")
ctx.set_profiling(True)
for _ in range(4):
    torch.cuda.synchronize(); t = time.time(); ctx.map_to_graph(p, pts); torch.cuda.synchronize(); ms = (time.time() - t) * 1e3
print("map ms (wall)", round(ms, 3), "nodes", ctx.graph()["n_nodes"], "launches", ctx.launch_count() if hasattr(ctx, "launch_count") else "?")
tot = 0.0
for k, v in ctx.stage_times():
    tot += v
    print(f"  {k:24s} {v:8.3f}")
print("  sum", round(tot, 3))
