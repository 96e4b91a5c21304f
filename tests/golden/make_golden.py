#!/usr/bin/env python
"""Regenerates tests/golden/*.npz.

The reference ships no tests, fixtures or sample data and cannot be built in this
image (ROS 2, PCL, Eigen and OpenCV C++ are absent), so these vectors come from the CPU oracle
(oracle/aos_oracle_*.c, a line-by-line restatement of the reference loops) driving the REAL OpenCV
cv2.Subdiv2D / cv2.morphologyEx of this container (cv2 4.13.0).  They pin (a) the oracle against silent
drift, (b) the library against known answers without running the oracle on the GPU box.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "active-orchard-slam_b200")]

from aos_gpu import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

CASES = {
    # name: (spec kwargs, params overrides)
    "tiny_seed7": (dict(extent_x=20.0, extent_y=12.0, row_pitch=4.0, n_points=40_000, outlier_count=3, seed=7), {}),
    "small_seed4_minlen0": (dict(extent_x=40.0, extent_y=24.0, n_points=120_000, outlier_count=6, seed=4),
                            dict(cluster_min_length=0.0)),
    "refpoly_seed5": (dict(extent_x=77.0, extent_y=14.0, origin_x=-4.5, origin_y=-2.4, row_pitch=3.5, n_points=150_000,
                           seed=5, exclusion=synth.REFERENCE_EXCLUSION_DISCS), dict(polygon=synth.REFERENCE_POLYGON)),
}
GRIDS = ("occ_raw", "occ_inflated", "occ_border", "opened", "skel", "skel_framed")
SEED_KEYS = ("cl_first", "cl_size", "cl_sumx", "cl_sumy", "cl_cx", "cl_cy", "cl_maxd2", "cl_len", "row_cluster", "rows",
             "rows_info", "seeds")
GRAPH_KEYS = ("merged_seeds", "facets_xy", "facet_off", "nodes", "node_labels", "node_cluster_indices", "node_label_counts",
              "node_label_clusters", "node_label_types", "edges", "edge_lengths", "edge_clearances", "corner_points")


def case_inputs(name):
    kw, over = CASES[name]
    spec = synth.OrchardSpec(**kw)
    pts = synth.make_orchard(spec)
    pk = dict(grid_resolution=spec.grid_resolution, inflation_radius=spec.inflation_radius, polygon=spec.polygon,
              exclusion=spec.exclusion)
    pk.update(over)
    return spec, pts, pk


def main():
    for name in CASES:
        spec, pts, pk = case_inputs(name)
        r = O.seed_stage(O.SeedParams(**pk), pts)
        g = O.gvd_stage(r["seeds"], r["skel_framed"], r["origin_x"], r["origin_y"], r["res"], r["rows_info"])
        out = {"w": r["w"], "h": r["h"], "origin_x": r["origin_x"], "origin_y": r["origin_y"], "res": r["res"],
               "counts": np.array([r["n_virtual"], r["n_ray"], r["n_endpoint"]]),
               "points_sha": np.frombuffer(__import__("hashlib").sha256(pts.tobytes()).digest(), np.uint8)}
        for k in GRIDS:
            out[k] = np.packbits(r[k] == 100, axis=1, bitorder="little")
        for k in SEED_KEYS:
            out[k] = r[k]
        for k in GRAPH_KEYS:
            out["g_" + k] = g[k]
        out["g_counts"] = np.array([g["n_voro_edges"], g["n_boundary_points_precrop"]])
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, r["w"], r["h"], "clusters", r["n_clusters"], "rows", r["n_rows"], "seeds", len(r["seeds"]), "nodes",
              len(g["nodes"]), "edges", len(g["edges"]), os.path.getsize(path), "bytes")


def trim_paths(name, g):
    """Deterministic paths over a golden case's grid: random chords (some cross tree rows) and one along the frame."""
    rng = np.random.default_rng(sum(map(ord, name)))
    w, h, res, ox, oy = int(g["w"]), int(g["h"]), float(g["res"]), float(g["origin_x"]), float(g["origin_y"])
    paths = []
    for k in range(12):
        a = np.array([ox + rng.uniform(1, w * res - 1), oy + rng.uniform(1, h * res - 1)])
        b = np.array([ox + rng.uniform(1, w * res - 1), oy + rng.uniform(1, h * res - 1)])
        t = np.linspace(0, 1, int(rng.integers(2, 600)))[:, None]
        paths.append(a + (b - a) * t)
    paths.append(np.stack([np.linspace(ox + 0.3, ox + w * res - 0.3, 200), np.full(200, oy + 0.12)], 1))   # hugs the frame
    return paths


def main_trim():
    """tests/golden/trim_paths.npz: trimPathNearOccupiedRegions (path_gen:1570-1630) on the golden skeletons."""
    out = {}
    for name in CASES:
        g = np.load(os.path.join(HERE, name + ".npz"))
        skel = np.unpackbits(g["skel_framed"], axis=1, bitorder="little")[:, :int(g["w"])].astype(np.int8) * 100
        kept = [O.trim_path(p, skel, float(g["origin_x"]), float(g["origin_y"]), g["res"], 0.2) for p in trim_paths(name, g)]
        out[name] = np.array(kept, np.int32)
        print(name, "trim kept", kept)
    np.savez_compressed(os.path.join(HERE, "trim_paths.npz"), **out)


if __name__ == "__main__":
    if "--trim-only" not in sys.argv:
        main()
    main_trim()
