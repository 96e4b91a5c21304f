"""Host Voronoi step of the gvd half (csrc/host_subdiv.cu: replay of cv::Subdiv2D) pinned bit-for-bit
against the real OpenCV implementation (cv2.Subdiv2D through oracle/subdiv.py), and the 0.5 m seed merge
(gvd:84-128) against the oracle's literal O(S^2) loop.  No GPU needed: these C-ABI entry points are host code."""
import os
import sys

import numpy as np
import pytest

from aos_gpu import lib

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden  # noqa: E402


def _seed_sets(rng, trial):
    n = int(rng.integers(1, 400))
    mode = trial % 5
    if mode == 0:    # uniform
        return rng.uniform(0, 50, (n, 2))
    if mode == 1:    # lattice: four co-circular points everywhere, with / without a tiny perturbation
        s = np.stack([rng.integers(0, 30, n) * 1.0, rng.integers(0, 20, n) * 2.0], 1)
        return s + rng.normal(0, 1e-3, (n, 2)) * (trial % 10 < 5)
    if mode == 2:    # exactly collinear rows, 0.7 m spacing (points on existing edges)
        return np.stack([np.arange(n) * 0.7 % 45, (np.arange(n) // 60) * 4.0 + 3], 1)
    if mode == 3:    # duplicates and points outside the rectangle (clipped with the 0.1 m margin)
        s = rng.uniform(-3, 53, (n, 2))
        s[::7] = s[0]
        return s
    # orchard-like: rows along x, 1 m seeds, slight slope
    rows = []
    for r in range(int(rng.integers(1, 8))):
        x = np.arange(int(rng.integers(5, 45))) * 1.0 + rng.uniform(0, 1)
        rows.append(np.stack([x, 3 + 4 * r + 0.01 * x * rng.uniform(-1, 1)], 1))
    return np.concatenate(rows)


@pytest.mark.parametrize("block", range(8))
def test_voronoi_facets_match_cv2_bit_for_bit(oracle, block):
    from oracle import subdiv
    rng = np.random.default_rng(100 + block)
    for trial in range(40):
        s = _seed_sets(rng, trial)
        b = (0.0, 50.0, 0.0, 40.0) if trial % 3 else (-4.5, 72.8, -2.4, 12.4)   # second: non-integer rectangle
        fx, fo, _ = subdiv.voronoi_facets(s, *b)
        gx, go = lib.voronoi_facets(s, *b)
        assert np.array_equal(fo, go), f"facet sizes differ (block {block} trial {trial})"
        assert np.array_equal(fx.view(np.uint32), gx.view(np.uint32)), f"facet vertices differ (block {block} trial {trial})"


def _hard_seed_sets(rng, trial):
    """Sets aimed at the replay's rare paths (scripts/dev/subdiv_stress.py runs the same generators by the 100 k)."""
    n = int(rng.integers(1, 500))
    mode = trial % 6
    if mode == 0:    # exact lattice in random order: points on existing edges, co-circular quadruples, duplicates
        return np.stack([rng.integers(0, 25, n) * 2.0, rng.integers(0, 20, n) * 2.0], 1) + 1.0
    if mode == 1:    # one exactly collinear line, 5 cm steps, wrapping around (every point on an edge)
        return np.stack([np.arange(n) * 0.05 % 45 + 1, np.full(n, 7.0)], 1)
    if mode == 2:    # tight clusters: near-duplicates 0.1 mm apart
        c = rng.uniform(5, 45, (8, 2))
        return c[rng.integers(0, 8, n)] + rng.normal(0, 1e-4, (n, 2))
    if mode == 3:    # co-circular points
        t = rng.uniform(0, 2 * np.pi, n)
        return np.stack([25 + 10 * np.cos(t), 20 + 10 * np.sin(t)], 1)
    if mode == 4:    # pairs one micrometre apart (LOC_VERTEX / tiny triangles)
        s = rng.uniform(1, 49, (n, 2))
        s[1::2] = s[::2][: len(s[1::2])] + 1e-6
        return s
    x = np.sort(rng.uniform(1, 49, n))   # a monotone curve: every insertion takes over a long fan
    return np.stack([x, 5 + 0.3 * np.sin(x)], 1)


def test_voronoi_facets_match_cv2_on_degenerate_sets(oracle):
    from oracle import subdiv
    rng = np.random.default_rng(4242)
    for trial in range(240):
        s = _hard_seed_sets(rng, trial)
        b = (0.0, 50.0, 0.0, 40.0) if trial % 3 else (-4.5, 72.8, -2.4, 42.4)
        fx, fo, _ = subdiv.voronoi_facets(s, *b)
        gx, go = lib.voronoi_facets(s, *b)
        assert np.array_equal(fo, go), f"facet sizes differ (trial {trial})"
        assert np.array_equal(fx.view(np.uint32), gx.view(np.uint32)), f"facet vertices differ (trial {trial})"


def test_voronoi_facets_empty_and_invalid_bounds():
    xy, off = lib.voronoi_facets(np.zeros((0, 2)), 0, 10, 0, 10)
    assert len(xy) == 0 and list(off) == [0]
    xy, off = lib.voronoi_facets(np.array([[1.0, 2.0]]), float("nan"), 10, 0, 10)
    assert len(xy) == 0
    # non-finite seeds are skipped (vd:66-70), degenerate bounds are widened to 1 m (vd:36-47)
    from oracle import subdiv
    s = np.array([[0.2, 0.3], [np.nan, 1.0], [0.4, 0.1], [np.inf, 0.0], [0.3, 0.6]])
    fx, fo, _ = subdiv.voronoi_facets(s, 0.0, 0.5, 0.0, 0.5)
    gx, go = lib.voronoi_facets(s, 0.0, 0.5, 0.0, 0.5)
    assert np.array_equal(fo, go) and np.array_equal(fx.view(np.uint32), gx.view(np.uint32))


def test_merge_seeds_matches_reference_loop(oracle):
    rng = np.random.default_rng(7)
    for trial in range(30):
        n = int(rng.integers(0, 600))
        s = rng.uniform(0, 12, (n, 2))
        if n > 10:
            s[5] = s[4] + [0.5, 0.0]      # exactly at the merge distance (<=)
            s[9] = [np.nan, 1.0]
        want = oracle.merge_seeds(s)
        got = lib.merge_seeds(s)
        want = want[np.isfinite(want).all(axis=1)]   # processGraph drops non-finite merged seeds (gvd:266-270)
        assert got.shape == want.shape and np.array_equal(got, want)


def test_host_voronoi_reproduces_golden_facets():
    """Host step only (no kernel): facets of the golden merged seeds, bit for bit."""
    for name in sorted(make_golden.CASES):
        g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
        minx, miny = float(g["origin_x"]), float(g["origin_y"])
        maxx = minx + float(np.float32(np.float32(int(g["w"])) * np.float32(g["res"])))
        maxy = miny + float(np.float32(np.float32(int(g["h"])) * np.float32(g["res"])))
        xy, off = lib.voronoi_facets(g["g_merged_seeds"], minx, maxx, miny, maxy)
        assert np.array_equal(off, g["g_facet_off"]) and np.array_equal(xy.view(np.uint32), g["g_facets_xy"].view(np.uint32))


def test_literal_splice_path_gives_the_same_facets(oracle):
    """The literal form of insert() (new_edge / splice / connect_edges, swapEdges as four splices: the switch
    aos_set_subdiv_literal_splices) == the replay's closed-form slot updates == cv2."""
    from oracle import subdiv
    L = lib.load()
    rng = np.random.default_rng(321)
    try:
        for trial in range(25):
            s = _seed_sets(rng, trial)
            b = (0.0, 50.0, 0.0, 40.0)
            fx, fo, _ = subdiv.voronoi_facets(s, *b)
            L.aos_set_subdiv_literal_splices(1)
            gx, go = lib.voronoi_facets(s, *b)
            L.aos_set_subdiv_literal_splices(0)
            hx, ho = lib.voronoi_facets(s, *b)
            assert np.array_equal(fo, go) and np.array_equal(fx.view(np.uint32), gx.view(np.uint32))
            assert np.array_equal(fo, ho) and np.array_equal(fx.view(np.uint32), hx.view(np.uint32))
    finally:
        L.aos_set_subdiv_literal_splices(0)


def test_scalar_and_avx2_flip_loops_give_the_same_facets(oracle):
    """aos_set_subdiv_simd: the scalar flip loop and the AVX2 one (four triangle areas per register) == cv2."""
    from oracle import subdiv
    L = lib.load()
    rng = np.random.default_rng(99)
    modes = [0] + ([1] if L.aos_set_subdiv_simd(1) == 0 else [])
    try:
        for trial in range(60):
            s = _hard_seed_sets(rng, trial) if trial % 2 else _seed_sets(rng, trial)
            b = (0.0, 50.0, 0.0, 40.0)
            fx, fo, _ = subdiv.voronoi_facets(s, *b)
            for m in modes:
                assert L.aos_set_subdiv_simd(m) == 0
                gx, go = lib.voronoi_facets(s, *b)
                assert np.array_equal(fo, go) and np.array_equal(fx.view(np.uint32), gx.view(np.uint32)), (trial, m)
        # the map-order pattern (long fans) in the scalar loop; the default mode runs it in the tests below
        assert L.aos_set_subdiv_simd(0) == 0
        s = _orchard_seed_order(rng, 12, 300)
        b = (0.0, 400.0, 0.0, 60.0)
        fx, fo, _ = subdiv.voronoi_facets(s, *b)
        gx, go = lib.voronoi_facets(s, *b)
        assert np.array_equal(fo, go) and np.array_equal(fx.view(np.uint32), gx.view(np.uint32))
        assert L.aos_set_subdiv_simd(2) != 0
    finally:
        L.aos_set_subdiv_simd(-1)


def _orchard_seed_order(rng, n_rows, per_row, pitch=4.0, half=1.95):
    """Seeds in the order a real map produces them: row after row, walking along the row, alternating between the two
    sides of the tree line (the insertion pattern with ~55 Lawson flips per seed that the replay's hot loop is tuned for)."""
    out = []
    for r in range(n_rows):
        x0 = rng.uniform(0, 1)
        slope = 0.002 * rng.uniform(-1, 1)
        for i in range(per_row):
            x = x0 + i * 0.98 + rng.normal(0, 0.01)
            yc = 3.0 + pitch * r + slope * x
            out.append((x, yc + half + rng.normal(0, 0.03)))
            out.append((x + 0.006, yc - half + rng.normal(0, 0.03)))
    return np.array(out)


def test_replay_matches_cv2_on_a_map_sized_seed_order(oracle):
    """16 k seeds in map order (rows of alternating sides), then 6 k uniform ones: the facets of the replay -- fused
    flips and literal splices -- equal cv2.Subdiv2D's bit for bit."""
    from oracle import subdiv
    rng = np.random.default_rng(2024)
    s = np.concatenate([_orchard_seed_order(rng, 20, 400), rng.uniform(1, 395, (6000, 2)) * [1.0, 0.2]])
    b = (0.0, 400.0, 0.0, 90.0)
    fx, fo, _ = subdiv.voronoi_facets(s, *b)
    gx, go = lib.voronoi_facets(s, *b)
    assert np.array_equal(fo, go) and np.array_equal(fx.view(np.uint32), gx.view(np.uint32))
    L = lib.load()
    try:
        L.aos_set_subdiv_literal_splices(1)
        hx, ho = lib.voronoi_facets(s[:6000], *b)
    finally:
        L.aos_set_subdiv_literal_splices(0)
    kx, ko = lib.voronoi_facets(s[:6000], *b)
    assert np.array_equal(ho, ko) and np.array_equal(hx.view(np.uint32), kx.view(np.uint32))


def test_replay_matches_cv2_at_c3_size(oracle):
    """240 k seeds in map order over 1 km x 1 km (the size of a C3 map's merged seed set): 1.4 M facet vertices,
    bit for bit against cv2.Subdiv2D."""
    from oracle import subdiv
    rng = np.random.default_rng(77)
    s = _orchard_seed_order(rng, 245, 490)
    b = (0.0, 1000.0, 0.0, 1000.0)
    fx, fo, _ = subdiv.voronoi_facets(s, *b)
    gx, go = lib.voronoi_facets(s, *b)
    assert len(go) - 1 == len(s) and len(gx) > 1_300_000
    assert np.array_equal(fo, go) and np.array_equal(fx.view(np.uint32), gx.view(np.uint32))


def test_outer_factor_changes_vertex_bits_not_topology(oracle):
    """cv::Subdiv2D::initDelaunay's outer triangle sits 3 x (OpenCV <= 4.5.x, the reference's platform and the library's
    default) or 6 x (4.13, this image) the rectangle away.  The far vertices change the flip history, so the factor is
    part of the arithmetic: on an orchard seed set the graph topology is the same, but a few per cent of the node
    coordinates move by float32 ulps.  (Hence aos_set_subdiv_outer_factor and the one-line probe a node runs against
    the OpenCV it links; conftest.py does the same with cv2.)"""
    import ctypes as C
    from aos_gpu import synth
    from oracle import subdiv
    L = lib.load()
    assert subdiv.outer_factor() in (3.0, 6.0)
    spec = synth.config("C2", seed=0, n_points=600_000)
    p = oracle.SeedParams(grid_resolution=spec.grid_resolution, inflation_radius=spec.inflation_radius, polygon=spec.polygon)
    oracle.set_fast(True)
    try:
        a = oracle.seed_stage(p, synth.make_orchard(spec))
        merged = oracle.merge_seeds(a["seeds"])
        minx, miny = float(a["origin_x"]), float(a["origin_y"])
        maxx = minx + float(np.float32(np.float32(a["w"]) * a["res"]))
        maxy = miny + float(np.float32(np.float32(a["h"]) * a["res"]))
        skel = np.ascontiguousarray(a["skel_framed"])
        rows = np.ascontiguousarray(a["rows_info"], np.float64)
        out = {}
        for f in (3.0, 6.0):
            L.aos_set_subdiv_outer_factor(C.c_float(f))
            xy, off = lib.voronoi_facets(merged, minx, maxx, miny, maxy)
            gr = oracle._Graph()
            rc = oracle.lib().orc_gvd_graph(xy.ctypes.data_as(oracle._PF), off.ctypes.data_as(oracle._P32), len(off) - 1,
                                            skel.ctypes.data_as(oracle._P8), skel.shape[1], skel.shape[0], minx, miny,
                                            C.c_float(float(a["res"])), rows.ctypes.data_as(oracle._PD), len(rows), C.byref(gr))
            assert rc == 0
            out[f] = (oracle._arr(gr.nodes, 2 * gr.n_nodes, np.float64).reshape(-1, 2),
                      oracle._arr(gr.edges, 2 * gr.n_edges, np.int32).reshape(-1, 2),
                      oracle._arr(gr.node_labels, gr.n_nodes, np.int32))
            oracle.lib().orc_graph_free(C.byref(gr))
            if f == subdiv.outer_factor():   # the factor of this image's cv2 reproduces the real Subdiv2D bit for bit
                fx, fo, _ = subdiv.voronoi_facets(merged, minx, maxx, miny, maxy)
                assert np.array_equal(fo, off) and np.array_equal(fx, xy)
    finally:
        oracle.set_fast(False)
        L.aos_set_subdiv_outer_factor(C.c_float(subdiv.outer_factor()))
    (n3, e3, l3), (n6, e6, l6) = out[3.0], out[6.0]
    assert n3.shape == n6.shape and np.array_equal(e3, e6) and np.array_equal(l3, l6)      # same graph ...
    d = np.abs(n3 - n6).max(axis=1)
    assert 0 < (d > 0).sum() < 0.2 * len(n3) and d.max() < 1e-4                              # ... different low bits
