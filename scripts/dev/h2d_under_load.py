"""What slows the 3.2 GB cloud uploads of the e2e bench from 55 GB/s (alone) to ~47 GB/s?  H2D rate of repeated pinned
copies (a) alone, (b) with 15 host threads running Subdiv2D replays, (c) with D2H copies of 132 MB in the other direction,
(d) with kernels streaming HBM on another stream."""
import os, sys, threading, time
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "active-orchard-slam_b200"))
from aos_gpu import lib

n = 800_000_000
h = torch.empty(n, dtype=torch.float32, pin_memory=True); h.fill_(1.0)
d = torch.empty(n, dtype=torch.float32, device="cuda")
s_up = torch.cuda.Stream()

def rate(reps=6):
    out = []
    for _ in range(reps):
        torch.cuda.synchronize(); t = time.perf_counter()
        with torch.cuda.stream(s_up):
            d.copy_(h, non_blocking=True)
        s_up.synchronize()
        out.append(n * 4 / (time.perf_counter() - t) / 1e9)
    return f"min {min(out):.1f} median {sorted(out)[len(out)//2]:.1f} max {max(out):.1f} GB/s"

print("alone:", rate(), flush=True)

# (b) CPU load: Subdiv2D replays of 240 k orchard-like seeds on 15 threads (ctypes releases the GIL)
rng = np.random.default_rng(0)
rows = []
for r in range(245):
    x = np.arange(980) * 1.0 + rng.uniform(0, 1)
    rows.append(np.stack([x + rng.normal(0, 0.02, 980), 3 + 4 * r + 0.01 * x * rng.uniform(-1, 1)], 1))
seeds = np.concatenate(rows)
stop = threading.Event()
def burn():
    while not stop.is_set():
        lib.voronoi_facets(seeds, 0.0, 1000.0, 0.0, 1000.0)
th = [threading.Thread(target=burn) for _ in range(15)]
for t in th: t.start()
time.sleep(1.0)
print("with 15 replay threads:", rate(), flush=True)
stop.set()
for t in th: t.join()

# (c) D2H traffic in the other direction
ho = torch.empty(33_000_000, dtype=torch.float32, pin_memory=True)
dd = torch.empty(33_000_000, dtype=torch.float32, device="cuda")
s_dn = torch.cuda.Stream()
stop = threading.Event()
def down():
    while not stop.is_set():
        with torch.cuda.stream(s_dn):
            ho.copy_(dd, non_blocking=True)
        s_dn.synchronize()
t = threading.Thread(target=down); t.start()
print("with back-to-back 132 MB D2H:", rate(), flush=True)
stop.set(); t.join()

# (d) HBM-streaming kernels on another stream
a = torch.empty(400_000_000, dtype=torch.float32, device="cuda"); b = torch.empty_like(a)
s_k = torch.cuda.Stream()
stop = threading.Event()
def kern():
    while not stop.is_set():
        with torch.cuda.stream(s_k):
            for _ in range(20):
                b.copy_(a)
        s_k.synchronize()
t = threading.Thread(target=kern); t.start()
print("with HBM-streaming kernels:", rate(), flush=True)
stop.set(); t.join()

# (e) small H2D copies from other threads (pageable 4 MB + pinned 27 MB) interleaving on the copy engine
hp = torch.empty(1_000_000, dtype=torch.float32)
hq = torch.empty(7_000_000, dtype=torch.float32, pin_memory=True)
dq = torch.empty(7_000_000, dtype=torch.float32, device="cuda")
stop = threading.Event()
def small():
    s = torch.cuda.Stream()
    while not stop.is_set():
        with torch.cuda.stream(s):
            dq[:1_000_000].copy_(hp, non_blocking=True)
            dq.copy_(hq, non_blocking=True)
        s.synchronize()
        time.sleep(0.02)
th = [threading.Thread(target=small) for _ in range(8)]
for t in th: t.start()
print("with 8 threads of small H2D copies:", rate(), flush=True)
stop.set()
for t in th: t.join()
