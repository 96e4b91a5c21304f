"""The C oracle against oracle/_ref -- the reference's OWN node classes compiled unmodified from /root/reference
(oracle/ref_seed.cpp, ref_gvd.cpp, ref_vd.cpp, ref_path.cpp; stand-in headers under oracle/ref_shim/) -- bit for bit.

This is what pins the oracle (and through it every GPU parity test) to the reference's compiled code for
generateOccupancyGrid, the PassThrough crop + exclusion discs, applyInflation, markBoundariesAsOccupied,
skeletonizeOccupancyGrid's conversions, clusterOccupiedCells, the length / polygon filters, convertClustersToTreeRows,
generateVirtualSeeds and its ray casts, generateRayPointsFromEndpoints, the endpoint seeds, the sorted rows message,
markPolygonBoundaryAsOccupied, voronoiSeedsCallback's merge, VoronoiDiagram::compute / extractBoundaryPoints,
buildGraphFromBoundaryPoints, edgePassesThroughOccupiedPixels, filterNodesAndEdgesOutsideGrid, the TL/TR/BL/BR search with
its ray-cast fallback, publishGraph, and trimPathNearOccupiedRegions.  Third-party calls inside the reference run the real
cv2 (morphologyEx, Subdiv2D) or, where the library is absent from the image (ximgproc::thinning, PCL), the restatement.

The reference hard-codes its 11 exclusion discs (seed_gen:487-499), so every case passes those discs to the oracle.
The library is built in the authoring container (where /root/reference exists) and travels as a prebuilt file."""
import numpy as np
import pytest

from aos_gpu import synth
from oracle import ref as R

R.build()
pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref/libaos_ref.so not built (no /root/reference here)")

GRAPH_KEYS = ("nodes", "edges", "edge_lengths", "edge_clearances", "node_labels", "node_cluster_indices", "node_label_counts",
              "node_label_clusters", "node_label_types")


def _compare(oracle, spec, pts, **over):
    kw = dict(grid_resolution=spec.grid_resolution, inflation_radius=spec.inflation_radius, polygon=spec.polygon,
              exclusion=synth.REFERENCE_EXCLUSION_DISCS)
    kw.update(over)
    p = oracle.SeedParams(**kw)
    a = oracle.seed_stage(p, pts)
    b = R.seed_stage(p, pts)
    assert (a["w"], a["h"], a["origin_x"], a["origin_y"], a["res"]) == (b["w"], b["h"], b["origin_x"], b["origin_y"], b["res"])
    for k in ("occ_border", "skel", "skel_framed"):
        assert np.array_equal(a[k], b[k]), f"{k}: {int((a[k] != b[k]).sum())} cells differ"
    # clusterOccupiedCells: discovery order, BFS cell order, float32 centres, lengths
    assert a["n_clusters"] == b["n_clusters"]
    for k in ("cl_size", "cl_first", "cl_cx", "cl_cy", "cl_len", "cl_cell_off", "cl_cells"):
        assert np.array_equal(a[k], b[k]), k
    assert a["seeds"].shape == b["seeds"].shape and np.array_equal(a["seeds"], b["seeds"])
    assert np.array_equal(a["rows_info"], b["rows_info"])
    ga = oracle.gvd_stage(a["seeds"], a["skel_framed"], a["origin_x"], a["origin_y"], a["res"], a["rows_info"])
    gb = R.gvd_stage(b["seeds"], b["skel_framed"], b["origin_x"], b["origin_y"], b["res"], b["rows_info"])
    assert np.array_equal(ga["merged_seeds"], gb["merged_seeds"])
    if len(ga["merged_seeds"]) == 0:
        assert gb["published"] == 0
        return a, None
    assert gb["published"] >= 1 and ga["n_voro_edges"] == gb["n_voro_edges"]
    assert (float(ga["resolution"]), ga["origin_x"], ga["origin_y"]) == (gb["resolution"], gb["origin_x"], gb["origin_y"])
    for k in GRAPH_KEYS:
        assert ga[k].shape == gb[k].shape, f"{k}: {ga[k].shape} vs {gb[k].shape}"
        assert np.array_equal(ga[k], gb[k]), f"{k} differs at {np.argwhere(ga[k] != gb[k])[:4].tolist()}"
    assert np.all(gb["nodes_xyz"][:, 2] == 0.0)
    return a, ga


@pytest.mark.parametrize("name,seed", [("TINY", 0), ("TINY", 1), ("TINY", 7), ("SMALL", 0), ("SMALL", 3), ("SMALL", 4)])
def test_oracle_equals_reference_small(oracle, name, seed):
    spec = synth.config(name, seed=seed)
    _compare(oracle, spec, synth.make_orchard(spec))


def test_oracle_equals_reference_c1(oracle):
    spec = synth.config("C1", n_points=400_000)
    a, g = _compare(oracle, spec, synth.make_orchard(spec))
    assert (a["w"], a["h"]) == (1000, 600) and g["nodes"].shape[0] > 100


def test_oracle_equals_reference_c2(oracle):
    spec = synth.config("C2", n_points=600_000)
    a, g = _compare(oracle, spec, synth.make_orchard(spec))
    assert (a["w"], a["h"]) == (2000, 1200) and a["n_rows"] >= 8 and g["nodes"].shape[0] > 500


def test_reference_default_polygon(oracle):
    """The node's built-in polygon (seed_gen:193-215: no /aos_planner/exploration_area message sent to the reference)
    and its exclusion discs on a cloud laid over that field; the oracle gets the same polygon as a parameter."""
    spec = synth.OrchardSpec(extent_x=77.0, extent_y=14.0, origin_x=-4.5, origin_y=-2.4, row_pitch=3.5, n_points=300_000, seed=5)
    pts = synth.make_orchard(spec)
    p = oracle.SeedParams(grid_resolution=0.05, inflation_radius=0.8, polygon=synth.REFERENCE_POLYGON,
                          exclusion=synth.REFERENCE_EXCLUSION_DISCS)
    a = oracle.seed_stage(p, pts)
    p_ref = oracle.SeedParams(grid_resolution=0.05, inflation_radius=0.8, polygon=np.zeros((0, 2)))   # n_poly = 0: built-in polygon
    b = R.seed_stage(p_ref, pts)
    assert (a["w"], a["h"]) == (b["w"], b["h"]) == (1546, 296)
    for k in ("occ_border", "skel", "skel_framed"):
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(a["seeds"], b["seeds"]) and np.array_equal(a["rows_info"], b["rows_info"])
    assert np.array_equal(a["cl_cells"], b["cl_cells"]) and np.array_equal(a["cl_cx"], b["cl_cx"])
    ga = oracle.gvd_stage(a["seeds"], a["skel_framed"], a["origin_x"], a["origin_y"], a["res"], a["rows_info"])
    gb = R.gvd_stage(b["seeds"], b["skel_framed"], b["origin_x"], b["origin_y"], b["res"], b["rows_info"])
    for k in GRAPH_KEYS:
        assert np.array_equal(ga[k], gb[k]), k


@pytest.mark.parametrize("rot,seed", [(12.0, 3), (-31.0, 4), (90.0, 5)])
def test_rotated_orchards(oracle, rot, seed):
    spec = synth.OrchardSpec(extent_x=60.0, extent_y=50.0, row_pitch=5.0, n_points=350_000, seed=seed, rotation_deg=rot)
    _compare(oracle, spec, synth.make_orchard(spec))


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_min_length_zero_tie_breaks(oracle, seed):
    spec = synth.config("TINY", seed=seed)
    spec.outlier_count = 12
    _compare(oracle, spec, synth.make_orchard(spec), cluster_min_length=0.0)


@pytest.mark.parametrize("res,infl", [(0.1, 0.8), (0.05, 0.6), (0.05, 1.0), (0.025, 0.8), (0.07, 0.5)])
def test_resolutions_and_radii(oracle, res, infl):
    spec = synth.OrchardSpec(extent_x=30.0, extent_y=20.0, row_pitch=5.0, n_points=120_000, seed=9, grid_resolution=res,
                             inflation_radius=infl)
    _compare(oracle, spec, synth.make_orchard(spec))


def test_long_rows_float32_sums(oracle):
    """400 m rows at x up to 8000 cells: the float32 running sums of clusterOccupiedCells pass 2^24 (seed_gen:1053-1059)."""
    spec = synth.OrchardSpec(extent_x=400.0, extent_y=12.0, row_pitch=4.0, n_points=250_000, gap_prob=0.0, jitter=0.0,
                             outlier_count=4, seed=11)
    a, _ = _compare(oracle, spec, synth.make_orchard(spec))
    assert a["cl_sumx"].max() >= (1 << 24)


def test_empty_cloud_and_nan_points(oracle):
    spec = synth.config("TINY", seed=0)
    pts = synth.make_orchard(spec)
    pts[::97, 0] = np.nan
    pts[5::101, 2] = np.inf
    _compare(oracle, spec, pts)
    _compare(oracle, spec, np.full((1, 4), np.nan, np.float32))


def test_raw_grid_without_discs(oracle):
    """generateOccupancyGrid itself (seed_gen:581-622): compared on a map whose points avoid every exclusion disc."""
    spec = synth.OrchardSpec(extent_x=40.0, extent_y=24.0, origin_x=100.0, origin_y=50.0, n_points=150_000, seed=2)
    pts = synth.make_orchard(spec)
    p = oracle.SeedParams(grid_resolution=0.05, inflation_radius=0.8, polygon=spec.polygon)
    a = oracle.seed_stage(p, pts)
    b = R.seed_stage(p, pts)
    assert np.array_equal(a["occ_raw"], b["occ_raw_nodisc"])
    assert np.array_equal(R.inflate(a["occ_raw"], 0.05, 0.8), a["occ_inflated"])
    assert np.array_equal(R.mark_borders(a["occ_inflated"]), a["occ_border"])
    assert np.array_equal(R.skeletonize(a["occ_inflated"]), a["skel"])


@pytest.mark.parametrize("seed", range(4))
def test_steps_on_random_grids(oracle, seed):
    rng = np.random.default_rng(seed)
    h, w = int(rng.integers(3, 90)), int(rng.integers(3, 130))
    g = np.where(rng.random((h, w)) < 0.03, 100, 0).astype(np.int8)
    import ctypes as C
    P8 = C.POINTER(C.c_int8)
    for res, infl in ((0.05, 0.8), (0.1, 0.35), (0.05, 0.0)):
        cells = int(np.float32(infl) / np.float32(res))
        out = np.empty_like(g)
        oracle.lib().orc_inflate(g.ctypes.data_as(P8), w, h, cells, out.ctypes.data_as(P8))
        assert np.array_equal(out, R.inflate(g, res, infl))
    out = np.empty_like(g)
    oracle.lib().orc_mark_borders(g.ctypes.data_as(P8), w, h, out.ctypes.data_as(P8))
    assert np.array_equal(out, R.mark_borders(g))


def test_point_in_polygon(oracle):
    rng = np.random.default_rng(1)
    poly = synth.REFERENCE_POLYGON
    for _ in range(4000):
        x, y = rng.uniform(-5, 75), rng.uniform(-3, 12)
        assert bool(oracle.lib().orc_point_in_polygon(x, y, poly.ctypes.data_as(__import__("ctypes").POINTER(__import__("ctypes").c_double)), 4)) == \
            R.point_in_polygon(x, y, poly)


def test_trim_path(oracle):
    spec = synth.config("SMALL", seed=1)
    p = oracle.SeedParams(grid_resolution=0.05, inflation_radius=0.8, polygon=spec.polygon)
    a = oracle.seed_stage(p, synth.make_orchard(spec))
    rng = np.random.default_rng(3)
    for _ in range(40):
        s, e = rng.uniform([1, 1], [39, 23]), rng.uniform([1, 1], [39, 23])
        n = int(rng.integers(2, 200))
        path = s[None, :] + (e - s)[None, :] * np.linspace(0, 1, n)[:, None]
        assert oracle.trim_path(path, a["skel_framed"], a["origin_x"], a["origin_y"], a["res"]) == \
            R.trim_path(path, a["skel_framed"], a["origin_x"], a["origin_y"], a["res"])
