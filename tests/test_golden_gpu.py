"""The library against the committed golden vectors (tests/golden/*.npz, made by make_golden.py from the
oracle + the real cv2.Subdiv2D): known answers, no oracle involved at run time."""
import os
import sys

import numpy as np
import pytest

from aos_gpu import lib

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden  # noqa: E402

pytestmark = pytest.mark.gpu

_GRID_IDS = dict(occ_raw=lib.GRID_RAW, occ_inflated=lib.GRID_INFLATED, occ_border=lib.GRID_OCCUPANCY,
                 opened=lib.GRID_OPENED, skel=lib.GRID_SKELETON, skel_framed=lib.GRID_SKELETON_FRAMED)


@pytest.mark.parametrize("name", sorted(make_golden.CASES))
def test_library_reproduces_golden_vectors(gpu_ctx, name):
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    spec, pts, pk = make_golden.case_inputs(name)
    gpu_ctx.map_to_graph(lib.SeedParams(**pk), pts)
    s = gpu_ctx.seed_summary()
    assert (s.info.width, s.info.height) == (int(g["w"]), int(g["h"]))
    assert s.info.origin_x == float(g["origin_x"]) and s.info.origin_y == float(g["origin_y"])
    for k, gid in _GRID_IDS.items():
        got = gpu_ctx.grid_int8(gid)
        assert set(np.unique(got)) <= {0, 100}
        assert np.array_equal(np.packbits(got == 100, axis=1, bitorder="little"), g[k]), k
    cl = gpu_ctx.clusters()
    assert np.array_equal(cl["label"], g["cl_first"]) and np.array_equal(cl["size"], g["cl_size"])
    assert np.array_equal(cl["sum_x"], g["cl_sumx"]) and np.array_equal(cl["sum_y"], g["cl_sumy"])
    assert np.array_equal(cl["center_x"], g["cl_cx"]) and np.array_equal(cl["center_y"], g["cl_cy"])
    assert np.array_equal(cl["max_d2"], g["cl_maxd2"]) and np.array_equal(cl["length"], g["cl_len"])
    rows = gpu_ctx.tree_rows()
    got_rows = np.stack([rows[k] for k in ("center_x", "center_y", "start_x", "start_y", "end_x", "end_y", "length")], 1) \
        if len(rows) else np.zeros((0, 7))
    assert np.array_equal(got_rows, g["rows"]) and np.array_equal(rows["cluster"], g["row_cluster"])
    seeds, counts, rows_info = gpu_ctx.select_seeds()
    assert list(counts) == list(g["counts"])
    assert np.array_equal(seeds, g["seeds"]) and np.array_equal(rows_info, g["rows_info"])
    gr = gpu_ctx.graph()
    for k in ("nodes", "node_labels", "node_cluster_indices", "node_label_counts", "node_label_clusters", "node_label_types",
              "edges", "edge_lengths", "edge_clearances", "corner_points"):
        assert np.array_equal(gr[k], g["g_" + k]), k
    assert gr["n_voronoi_edges"] == int(g["g_counts"][0]) and gr["n_boundary_points"] == int(g["g_counts"][1])
    assert gr["n_merged_seeds"] == len(g["g_merged_seeds"])


@pytest.mark.parametrize("name", sorted(make_golden.CASES))
def test_device_steps_reproduce_golden_vectors(gpu_ctx, name):
    """The device-side pieces of the gvd half on their own, against the golden vectors (no oracle at run time): the
    0.5 m seed merge, circumcentres + facet walks (cv2.Subdiv2D's facets), and trimPathNearOccupiedRegions."""
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    g = np.load(os.path.join(here, name + ".npz"))
    merged = gpu_ctx.merge_seeds_device(g["seeds"])
    assert np.array_equal(merged.view(np.uint64), g["g_merged_seeds"].view(np.uint64))
    minx, miny = float(g["origin_x"]), float(g["origin_y"])
    maxx = minx + float(np.float32(np.float32(int(g["w"])) * np.float32(g["res"])))
    maxy = miny + float(np.float32(np.float32(int(g["h"])) * np.float32(g["res"])))
    xy, nxt = gpu_ctx.voronoi_facets_device(g["g_merged_seeds"], minx, maxx, miny, maxy)
    off, fxy = g["g_facet_off"], g["g_facets_xy"]
    keep = [(int(off[f]), int(off[f + 1])) for f in range(len(off) - 1) if off[f + 1] - off[f] >= 2]
    want_xy = np.concatenate([fxy[a:b] for a, b in keep]) if keep else np.zeros((0, 2), np.float32)
    assert np.array_equal(xy.view(np.uint32), want_xy.view(np.uint32))
    base, want_next = 0, []
    for a, b in keep:
        n = np.arange(base + 1, base + (b - a) + 1, dtype=np.int32)
        n[-1] = base
        want_next.append(n)
        base += b - a
    assert np.array_equal(nxt, np.concatenate(want_next) if want_next else np.zeros(0, np.int32))
    trim = np.load(os.path.join(here, "trim_paths.npz"))[name]
    bits = np.ascontiguousarray(lib.pack_bits(np.unpackbits(g["skel_framed"], axis=1, bitorder="little")[:, :int(g["w"])] != 0))
    info = (float(g["res"]), minx, miny, int(g["w"]))
    got = [gpu_ctx.trim_path(p, 0.2, skeleton_bits=bits, info=info) for p in make_golden.trim_paths(name, g)]
    assert got == list(trim)
