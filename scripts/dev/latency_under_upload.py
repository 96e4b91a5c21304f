"""How much do the kernel phases of a map stretch while ANOTHER stream uploads a cloud over PCIe?
One context runs device-resident C3 maps back to back; a second thread uploads 3.2 GB pinned clouds in a loop:
(0) nothing, (1) one monolithic copy, (2) 16 MB pieces, two queued at a time, (3) as (2) on a low-priority stream."""
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
ctx = lib.Context(0)
ctx.map_to_graph(params, pts)
h = torch.empty(pts.shape, dtype=torch.float32, pin_memory=True)
d = torch.empty_like(pts)
hb, db = h.view(-1), d.view(-1)
piece = 4 * 1024 * 1024   # floats = 16 MB

def seed_time(reps=4):
    out = []
    for _ in range(reps):
        t = time.perf_counter(); ctx.seed_stage(params, pts); a = time.perf_counter(); ctx.select_seeds(); b = time.perf_counter()
        out.append(((a - t) * 1e3, (b - a) * 1e3))
    return "seed stage ms " + " ".join(f"{x[0]:.1f}" for x in out) + " | select ms " + " ".join(f"{x[1]:.1f}" for x in out)

stop = threading.Event()
def uploader(mode):
    torch.cuda.set_device(0)
    lo, hi = torch.cuda.Stream.priority_range()
    s = torch.cuda.Stream(priority=lo if mode == 3 else hi) if mode == 3 else torch.cuda.Stream()
    ev = [torch.cuda.Event(), torch.cuda.Event()]
    n = 0
    while not stop.is_set():
        with torch.cuda.stream(s):
            if mode == 1:
                d.copy_(h, non_blocking=True)
            else:
                k = 0
                for off in range(0, hb.numel(), piece):
                    if k >= 2:
                        ev[k & 1].synchronize()
                    db[off:off + piece].copy_(hb[off:off + piece], non_blocking=True)
                    ev[k & 1].record(s)
                    k += 1
        s.synchronize()
        n += 1

print("no upload:      ", seed_time(), flush=True)
for mode, name in ((1, "monolithic     "), (2, "16 MB paced    "), (3, "paced, low prio")):
    stop.clear()
    t = threading.Thread(target=uploader, args=(mode,)); t.start()
    time.sleep(0.3)
    print(name + ":", seed_time(), flush=True)
    stop.set(); t.join()
