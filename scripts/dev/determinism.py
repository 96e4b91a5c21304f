"""Run-to-run determinism probe: the same cloud through aos_map_to_graph on several contexts at once, several times;
every result digest must be identical.  Usage: python scripts/dev/determinism.py [workload] [contexts] [repeats]"""
import concurrent.futures as cf
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "active-orchard-slam_b200")]
import torch  # noqa: E402
from aos_gpu import lib, synth  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "C3"
T = int(sys.argv[2]) if len(sys.argv) > 2 else 4
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda", 0)
spec = synth.config(wl, seed=0)
pts = synth.make_orchard_torch(spec, dev)
params = lib.SeedParams(grid_resolution=spec.grid_resolution, inflation_radius=spec.inflation_radius, polygon=spec.polygon)
ctxs = [lib.Context(0) for _ in range(T)]
pool = cf.ThreadPoolExecutor(T)
seen = {}
for r in range(reps):
    t0 = time.time()
    def work(t):
        torch.cuda.set_device(0)
        ctxs[t].map_to_graph(params, pts)
        return ctxs[t].result_digest(parts=True)
    out = list(pool.map(work, range(T)))
    for t, (tot, per) in enumerate(out):
        seen.setdefault(tot, []).append((r, t))
        if len(seen) > 1 and tot != next(iter(seen)):
            first = None
            for k, v in per.items():
                if v != base_per[k]:
                    first = k if first is None else first
                    print("DIFF rep", r, "ctx", t, k, flush=True)
        else:
            base_per = per
    print(f"rep {r}: {time.time() - t0:.1f}s digests {sorted(set(o[0][:12] for o in out))}", flush=True)
print(json.dumps({"workload": wl, "contexts": T, "reps": reps, "distinct": len(seen), "digests": {k[:16]: v for k, v in seen.items()}}))
