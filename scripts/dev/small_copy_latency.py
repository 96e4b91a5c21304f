"""Latency of a small pinned H2D copy (4 MB) + sync on its own stream while the library uploads clouds (aos_seed_stage
with host points: 16 MB pieces, two queued) from another thread."""
import os, sys, threading, time
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "active-orchard-slam_b200"))
from aos_gpu import lib, synth

spec = synth.config("C3", seed=0)
dev = torch.device("cuda", 0)
pts = synth.make_orchard_torch(spec, dev)
params = lib.SeedParams(grid_resolution=spec.grid_resolution, inflation_radius=spec.inflation_radius, polygon=spec.polygon,
                        exclusion=spec.exclusion)
h = torch.empty(pts.shape, dtype=torch.float32, pin_memory=True); h.copy_(pts)
host_np = h.numpy()
ctx = lib.Context(0)
ctx.seed_stage(params, host_np)
small_h = torch.empty(1_000_000, dtype=torch.float32, pin_memory=True)
small_d = torch.empty(1_000_000, dtype=torch.float32, device=dev)
s = torch.cuda.Stream()
k_d = torch.empty(1_000_000, dtype=torch.float32, device=dev)

def lat(n=60, what="copy"):
    out = []
    for _ in range(n):
        t = time.perf_counter()
        with torch.cuda.stream(s):
            if what == "copy":
                small_d.copy_(small_h, non_blocking=True)
            elif what == "kernel":
                k_d.add_(1.0)
            else:
                small_h.copy_(small_d, non_blocking=True)
        s.synchronize()
        out.append((time.perf_counter() - t) * 1e3)
        time.sleep(0.005)
    out.sort()
    return f"min {out[0]:.2f} median {out[len(out)//2]:.2f} p90 {out[int(len(out)*0.9)]:.2f} max {out[-1]:.2f} ms"

for what in ("copy", "kernel", "d2h"):
    print(f"idle      {what:6s}", lat(what=what), flush=True)
stop = threading.Event()
def up():
    torch.cuda.set_device(0)
    while not stop.is_set():
        ctx.seed_stage(params, host_np)
t = threading.Thread(target=up); t.start()
time.sleep(0.3)
for what in ("copy", "kernel", "d2h"):
    print(f"uploading {what:6s}", lat(what=what), flush=True)
stop.set(); t.join()
