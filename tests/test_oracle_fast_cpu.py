"""The oracle's fast mode (hash grids, convex hulls, threads: oracle/aos_oracle_fast.c) against its literal loops.
Fast mode exists so that the oracle finishes BASELINE config 3 on a CPU (full-size parity tests, the same-config CPU arm
of bench.py); it is only trusted because every artefact is identical to the literal restatement, which in turn is pinned
to the compiled reference (tests/test_ref_cpu.py)."""
import numpy as np
import pytest

from aos_gpu import synth

GRAPH_KEYS = ("merged_seeds", "nodes", "edges", "edge_lengths", "edge_clearances", "node_labels", "node_cluster_indices",
              "node_label_counts", "node_label_clusters", "node_label_types", "corner_points")
SEED_KEYS = ("occ_raw", "occ_inflated", "occ_border", "opened", "skel", "skel_framed", "labels", "cl_first", "cl_size", "cl_sumx",
             "cl_sumy", "cl_cx", "cl_cy", "cl_maxd2", "cl_len", "cl_cell_off", "cl_cells", "row_cluster", "rows", "rows_info",
             "seeds")


def _both(oracle, p, pts, threads=3):
    oracle.set_fast(False)
    a = oracle.seed_stage(p, pts)
    ga = oracle.gvd_stage(a["seeds"], a["skel_framed"], a["origin_x"], a["origin_y"], a["res"], a["rows_info"])
    try:
        oracle.set_fast(True, threads=threads)
        b = oracle.seed_stage(p, pts)
        gb = oracle.gvd_stage(b["seeds"], b["skel_framed"], b["origin_x"], b["origin_y"], b["res"], b["rows_info"])
    finally:
        oracle.set_fast(False)
    for k in SEED_KEYS:
        assert np.array_equal(a[k], b[k]), k
    assert (a["n_virtual"], a["n_ray"], a["n_endpoint"]) == (b["n_virtual"], b["n_ray"], b["n_endpoint"])
    assert set(ga) == set(gb)
    for k in GRAPH_KEYS:
        if k in ga:
            assert ga[k].shape == gb[k].shape and np.array_equal(ga[k], gb[k]), k
    return a, ga


@pytest.mark.parametrize("name,seed,npts", [("TINY", 0, None), ("TINY", 7, None), ("SMALL", 1, None), ("SMALL", 4, None),
                                            ("C1", 0, 400_000), ("C2", 0, 600_000)])
def test_fast_equals_literal(oracle, name, seed, npts):
    spec = synth.config(name, seed=seed, n_points=npts)
    p = oracle.SeedParams(grid_resolution=spec.grid_resolution, inflation_radius=spec.inflation_radius, polygon=spec.polygon)
    _both(oracle, p, synth.make_orchard(spec))


def test_fast_reference_polygon_and_discs(oracle):
    spec = synth.OrchardSpec(extent_x=77.0, extent_y=14.0, origin_x=-4.5, origin_y=-2.4, row_pitch=3.5, n_points=300_000, seed=5)
    p = oracle.SeedParams(grid_resolution=0.05, inflation_radius=0.8, polygon=synth.REFERENCE_POLYGON,
                          exclusion=synth.REFERENCE_EXCLUSION_DISCS)
    _both(oracle, p, synth.make_orchard(spec))


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_fast_min_length_zero(oracle, seed):
    spec = synth.config("TINY", seed=seed)
    spec.outlier_count = 12
    p = oracle.SeedParams(grid_resolution=spec.grid_resolution, inflation_radius=spec.inflation_radius, polygon=spec.polygon,
                          cluster_min_length=0.0)
    _both(oracle, p, synth.make_orchard(spec), threads=5)


@pytest.mark.parametrize("rot,seed", [(17.0, 3), (-40.0, 4)])
def test_fast_rotated(oracle, rot, seed):
    spec = synth.OrchardSpec(extent_x=60.0, extent_y=50.0, row_pitch=5.0, n_points=350_000, seed=seed, rotation_deg=rot)
    p = oracle.SeedParams(grid_resolution=spec.grid_resolution, inflation_radius=spec.inflation_radius, polygon=spec.polygon)
    _both(oracle, p, synth.make_orchard(spec), threads=8)


@pytest.mark.parametrize("seed", range(6))
def test_fast_steps_random(oracle, seed):
    """inflation / opening / thinning / diameter on random images of odd shapes, every radius class."""
    import ctypes as C
    P8 = C.POINTER(C.c_int8)
    rng = np.random.default_rng(100 + seed)
    h, w = int(rng.integers(1, 120)), int(rng.integers(1, 150))
    g = np.where(rng.random((h, w)) < rng.choice([0.02, 0.3, 0.7]), 100, 0).astype(np.int8)
    L = oracle.lib()
    L.orc_fast_inflate.argtypes = [P8, C.c_int, C.c_int, C.c_int, P8]
    L.orc_fast_open_cross.argtypes = [P8, C.c_int, C.c_int, P8]
    L.orc_fast_thin_zhangsuen.argtypes = [P8, C.c_int, C.c_int]
    L.orc_fast_thin_zhangsuen.restype = C.c_int
    oracle.set_fast(True, threads=4)
    try:
        for cells in (0, 1, 2, 5, 16):
            a, b = np.empty_like(g), np.empty_like(g)
            L.orc_inflate(g.ctypes.data_as(P8), w, h, cells, a.ctypes.data_as(P8))
            L.orc_fast_inflate(g.ctypes.data_as(P8), w, h, cells, b.ctypes.data_as(P8))
            assert np.array_equal(a, b), cells
        a, b = np.empty_like(g), np.empty_like(g)
        L.orc_open_cross(g.ctypes.data_as(P8), w, h, a.ctypes.data_as(P8))
        L.orc_fast_open_cross(g.ctypes.data_as(P8), w, h, b.ctypes.data_as(P8))
        assert np.array_equal(a, b)
        blob = np.empty_like(g)
        L.orc_inflate(g.ctypes.data_as(P8), w, h, 3, blob.ctypes.data_as(P8))
        t1, t2 = blob.copy(), blob.copy()
        n1 = L.orc_thin_zhangsuen(t1.ctypes.data_as(P8), w, h)
        n2 = L.orc_fast_thin_zhangsuen(t2.ctypes.data_as(P8), w, h)
        assert np.array_equal(t1, t2) and n1 == n2
    finally:
        oracle.set_fast(False)


def test_fast_merge_seeds_random(oracle):
    rng = np.random.default_rng(5)
    for n in (0, 1, 2, 50, 2000):
        s = rng.uniform(0, 12, (n, 2))
        if n > 10:
            s[3] = s[7]
            s[11] = [np.nan, 1.0]
        oracle.set_fast(False)
        a = oracle.merge_seeds(s)
        oracle.set_fast(True, threads=2)
        try:
            b = oracle.merge_seeds(s)
        finally:
            oracle.set_fast(False)
        assert np.array_equal(a, b, equal_nan=True)
