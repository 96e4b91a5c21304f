"""Generates tests/golden/fullsize_digests.json: per-array sha256 digests of the CPU oracle's results (fast mode, proven
equal to the literal loops by tests/test_oracle_fast_cpu.py and pinned to the compiled reference by tests/test_ref_cpu.py)
on full-size seeded clouds -- BASELINE config 3 (20000 x 20000 cells, 200 M points) and friends.  The GPU tests
(tests/test_fullsize_gpu.py) regenerate the same numpy cloud (synth.make_orchard_strips: IEEE-exact operations only) and compare aos_gpu.lib.digest_of() part by part.

    python tests/golden/make_fullsize_digests.py [C3 [C3H ...]]        (about 2 minutes and 12 GB per config 3 map)
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "active-orchard-slam_b200")]
import numpy as np  # noqa: E402
from aos_gpu import lib, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

CASES = {   # name -> (workload, seed, points)
    "C3": ("C3", 0, None),
    "C3_40M": ("C3", 3, 40_000_000),
    "C4": ("C4", 0, None),          # 40000 x 40000 cells @ 0.025 m, R = 32 (about 20 GB and 2 minutes)
}


def run(case):
    wl, seed, npts = CASES[case]
    spec = synth.config(wl, seed=seed, n_points=npts)
    t0 = time.time()
    pts = synth.make_orchard_strips(spec)
    t_gen = time.time() - t0
    p = O.SeedParams(grid_resolution=spec.grid_resolution, inflation_radius=spec.inflation_radius, polygon=spec.polygon)
    O.set_fast(True, skip_labels=True)
    t0 = time.time()
    r = O.seed_stage(p, pts)
    t_seed = time.time() - t0
    t0 = time.time()
    g = O.gvd_stage(r["seeds"], r["skel_framed"], r["origin_x"], r["origin_y"], r["res"], r["rows_info"])
    t_gvd = time.time() - t0
    total, per = lib.digest_of(O.result_artefacts(r, g), parts=True)
    return {"workload": wl, "seed": seed, "points": int(len(pts)), "width": r["w"], "height": r["h"], "n_clusters": int(r["n_clusters"]),
            "n_rows": int(r["n_rows"]), "n_seeds": int(len(r["seeds"])), "n_nodes": int(len(g["nodes"])), "n_edges": int(len(g["edges"])),
            "digest": total, "parts": per,
            "cpu_seconds": {"generate": round(t_gen, 1), "seed_stage": round(t_seed, 1), "gvd_stage": round(t_gvd, 1),
                            "threads": os.cpu_count()}}


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fullsize_digests.json")
    out = json.load(open(path)) if os.path.exists(path) else {}
    for case in (sys.argv[1:] or list(CASES)):
        out[case] = run(case)
        print(case, {k: v for k, v in out[case].items() if k != "parts"}, flush=True)
        json.dump(out, open(path, "w"), indent=1, sort_keys=True)
