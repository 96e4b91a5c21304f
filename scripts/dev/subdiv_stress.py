"""Randomised stress of the Subdiv2D replay (csrc/host_subdiv.cu) against the real cv2.Subdiv2D on seed sets chosen to
hit the rare paths: exact lattices (points on edges, co-circular quadruples, duplicates), one exactly collinear line,
clusters of near-duplicates, co-circular points, pairs 1 um apart, monotone curves (long fans), map-order rows.
    python scripts/dev/subdiv_stress.py <rng seed> <seconds>
185 k sets, 0 mismatches on the round-2 replay (four processes x 90 s)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, 'active-orchard-slam_b200'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
from aos_gpu import lib
from oracle import subdiv
import ctypes
L = lib.load()
L.aos_set_subdiv_outer_factor.argtypes = [ctypes.c_float]
L.aos_set_subdiv_outer_factor(ctypes.c_float(subdiv.outer_factor()))
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
def gen(trial):
    n = int(rng.integers(1, 600))
    m = trial % 9
    if m == 0: return rng.uniform(0, 50, (n, 2))
    if m == 1: return np.stack([rng.integers(0, 25, n) * 2.0, rng.integers(0, 20, n) * 2.0], 1) + 1.0   # exact lattice: on-edge, co-circular, duplicates
    if m == 2: return np.stack([np.arange(n) * 0.05 % 45 + 1, np.full(n, 7.0)], 1)                       # one exactly collinear line, 5 cm steps
    if m == 3:
        c = rng.uniform(5, 45, (8, 2)); return c[rng.integers(0, 8, n)] + rng.normal(0, 1e-4, (n, 2))      # tight clusters (near-duplicates)
    if m == 4:
        t = rng.uniform(0, 2 * np.pi, n); return np.stack([25 + 10 * np.cos(t), 20 + 10 * np.sin(t)], 1)   # co-circular
    if m == 5:
        s = rng.uniform(1, 49, (n, 2)); s[1::2] = s[::2][: len(s[1::2])] + 1e-6; return s                   # pairs 1 um apart
    if m == 6:
        x = np.sort(rng.uniform(1, 49, n)); return np.stack([x, 5 + 0.3 * np.sin(x)], 1)                     # monotone curve (long fans)
    if m == 7:
        g = np.stack(np.meshgrid(np.arange(1, 30), np.arange(1, 20)), -1).reshape(-1, 2).astype(float); rng.shuffle(g); return g[:n]
    rows = []
    for r in range(int(rng.integers(1, 8))):
        x = np.arange(int(rng.integers(5, 45))) * 1.0 + rng.uniform(0, 1)
        a = np.stack([x, 3 + 4 * r + 1.95 + 0 * x], 1); b = np.stack([x + 0.006, 3 + 4 * r - 1.95 + 0 * x], 1)
        rows.append(np.stack([a, b], 1).reshape(-1, 2))
    return np.concatenate(rows)
t0 = time.time(); trial = 0; bad = 0
while time.time() - t0 < float(sys.argv[2]) if len(sys.argv) > 2 else 60:
    s = gen(trial)
    b = (0.0, 50.0, 0.0, 40.0) if trial % 3 else (-4.5, 72.8, -2.4, 42.4)
    fx, fo, _ = subdiv.voronoi_facets(s, *b)
    gx, go = lib.voronoi_facets(s, *b)
    if not (np.array_equal(fo, go) and np.array_equal(fx.view(np.uint32), gx.view(np.uint32))):
        bad += 1; print("MISMATCH trial", trial, "mode", trial % 9, "n", len(s)); np.save(f"/tmp/subdiv_bad_{trial}.npy", s)
    trial += 1
print("trials", trial, "mismatches", bad)
