"""GPU parity of the gvd half (aos_gvd_stage / aos_map_to_graph) against the oracle (C restatement of
aos_gvd_node around the real cv2.Subdiv2D): GvdGraph.msg arrays bit-exact, in the reference's order."""
import numpy as np
import pytest

from aos_gpu import lib, synth
from helpers import assert_graph_parity, params_pair

pytestmark = pytest.mark.gpu


def _oracle_graph(oracle, r):
    return oracle.gvd_stage(r["seeds"], r["skel_framed"], r["origin_x"], r["origin_y"], r["res"], r["rows_info"])


@pytest.mark.parametrize("name,seed,npts", [("TINY", 0, None), ("TINY", 3, None), ("SMALL", 1, None), ("C1", 0, 400_000),
                                            ("C2", 0, 600_000)])
def test_map_to_graph(gpu_ctx, oracle, name, seed, npts):
    spec = synth.config(name, seed=seed, n_points=npts)
    pts = synth.make_orchard(spec)
    po, pl = params_pair(spec, oracle)
    r = oracle.seed_stage(po, pts)
    ref = _oracle_graph(oracle, r)
    info = gpu_ctx.map_to_graph(pl, pts)
    assert info["graph"] is not None
    assert_graph_parity(gpu_ctx.graph(), ref)


def test_gvd_stage_from_messages(gpu_ctx, oracle):
    """The aos_gvd_node side of the drop-in: seeds, rows and the int8 skeleton arrive as messages
    (a separate process, no seed stage on this context)."""
    spec = synth.config("SMALL", seed=2)
    pts = synth.make_orchard(spec)
    po, _ = params_pair(spec, oracle)
    r = oracle.seed_stage(po, pts)
    ref = _oracle_graph(oracle, r)
    ctx = lib.Context(0)
    got = ctx.gvd_stage(r["seeds"], r["rows_info"], skeleton=r["skel_framed"], info=(r["res"], r["origin_x"], r["origin_y"]))
    assert_graph_parity(got, ref)
    # bit-packed side channel (row F4): same graph from the 8x smaller AOS_FMT_BITS grid
    bits = lib.pack_bits(r["skel_framed"] == 100)
    got2 = ctx.gvd_stage(r["seeds"], r["rows_info"], skeleton=bits, info=(r["res"], r["origin_x"], r["origin_y"], r["w"]))
    assert_graph_parity(got2, ref)
    ctx.close()


def test_reference_polygon_graph(gpu_ctx, oracle):
    """Negative, non-integer origin (the reference's default polygon): Subdiv2D's integer rectangle, clipping
    margin and the crop bounds all differ from the grid extent."""
    spec = synth.OrchardSpec(extent_x=77.0, extent_y=14.0, origin_x=-4.5, origin_y=-2.4, row_pitch=3.5,
                             n_points=300_000, seed=5, exclusion=synth.REFERENCE_EXCLUSION_DISCS)
    pts = synth.make_orchard(spec)
    po, pl = params_pair(spec, oracle, polygon=synth.REFERENCE_POLYGON)
    r = oracle.seed_stage(po, pts)
    ref = _oracle_graph(oracle, r)
    gpu_ctx.map_to_graph(pl, pts)
    assert_graph_parity(gpu_ctx.graph(), ref)


def test_gvd_stage_random_seeds_and_obstacles(gpu_ctx, oracle):
    """Seeds that do not come from rows: random seeds (many within 0.5 m of each other -> merges, near-duplicate
    Voronoi vertices, proximity edges), a random obstacle raster, random rows near the border (ray-cast corner
    fallback, gvd:558-684)."""
    rng = np.random.default_rng(3)
    for trial in range(4):
        w, h, res = 640, 400, np.float32(0.05)
        skel = np.zeros((h, w), np.int8)
        for _ in range(30):   # random line obstacles
            x0, y0 = rng.integers(0, w), rng.integers(0, h)
            L = rng.integers(5, 120)
            if rng.random() < 0.5:
                skel[y0, x0:x0 + L] = 100
            else:
                skel[y0:y0 + L, x0] = 100
        skel[0, :] = skel[-1, :] = 100
        skel[:, 0] = skel[:, -1] = 100
        ox, oy = -3.25 + trial, 1.5 * trial
        n = 400
        seeds = np.stack([ox + rng.uniform(0, w * 0.05, n), oy + rng.uniform(0, h * 0.05, n)], 1)
        seeds[::9] += 30.0          # some seeds outside the grid (clipped by Subdiv2D's rectangle margin)
        rows = np.zeros((6, 4))
        rows[:, 0] = ox + rng.uniform(0.2, 5, 6)
        rows[:, 2] = ox + rng.uniform(20, 31.5, 6)
        rows[:, 1] = rows[:, 3] = oy + rng.uniform(0.3, h * 0.05 - 0.3, 6)
        rows[5] = rows[5, [2, 3, 0, 1]]   # start.x > end.x: swapped by the node (gvd:140-145)
        ref = oracle.gvd_stage(seeds, skel, ox, oy, res, rows)
        got = gpu_ctx.gvd_stage(seeds, rows, skeleton=skel, info=(res, ox, oy))
        assert_graph_parity(got, ref)


def test_gvd_stage_needs_inputs(oracle):
    ctx = lib.Context(0)
    with pytest.raises(lib.AosError):   # no skeleton on this context (gvd:257)
        ctx.gvd_stage(np.zeros((3, 2)), np.zeros((0, 4)))
    with pytest.raises(lib.AosError):   # no valid seeds (gvd:273)
        ctx.gvd_stage(np.full((2, 2), np.nan), np.zeros((0, 4)), skeleton=np.zeros((8, 8), np.int8), info=(0.05, 0.0, 0.0))
    ctx.close()


def test_sweep_of_maps_in_flight(oracle):
    """BASELINE config 5 in miniature: a sweep over row pitch / inflation radius / resolution, all maps in flight at
    once through aos_map_to_graph_batch (one context and host thread per map), each bit-exact against the oracle."""
    sweep = [(3.5, 0.6, 0.05), (4.0, 0.8, 0.05), (5.0, 1.0, 0.05), (6.0, 0.8, 0.1), (4.5, 0.7, 0.1), (8.0, 1.0, 0.05),
             (4.0, 0.8, 0.025), (6.0, 0.8, 0.2)]   # config 4's resolution (R = 32 cells) and a coarse grid (R = 4)
    ctxs, prms, clouds, refs = [], [], [], []
    for i, (pitch, infl, res) in enumerate(sweep):
        spec = synth.OrchardSpec(extent_x=30.0, extent_y=24.0, row_pitch=pitch, n_points=90_000, outlier_count=4, seed=20 + i,
                                 grid_resolution=res, inflation_radius=infl)
        pts = synth.make_orchard(spec)
        po, pl = params_pair(spec, oracle)
        r = oracle.seed_stage(po, pts)
        refs.append((r, _oracle_graph(oracle, r)))
        ctxs.append(lib.Context(0))
        prms.append(pl)
        clouds.append(pts)
    status = lib.map_to_graph_batch(ctxs, prms, clouds)
    from helpers import assert_seed_parity
    for ctx, st, (r, g) in zip(ctxs, status, refs):
        assert st == 0
        assert_seed_parity(ctx, r)
        assert_graph_parity(ctx.graph(), g)
        ctx.close()


def test_blocking_host_waits_give_the_same_graph(oracle):
    """aos_set_host_wait(device, 1): threads waiting for the GPU sleep instead of spinning (what bench.py runs its
    throughput legs with); the switch is about scheduling only -- same graph, and it can be switched back."""
    L = lib.load()
    spec = synth.config("SMALL", seed=5)
    pts = synth.make_orchard(spec)
    po, pl = params_pair(spec, oracle)
    r = oracle.seed_stage(po, pts)
    ref = _oracle_graph(oracle, r)
    assert L.aos_set_host_wait(99, 1) != 0            # no such device
    try:
        assert L.aos_set_host_wait(0, 1) == 0
        ctx = lib.Context(0)
        ctx.map_to_graph(pl, pts)
        assert_graph_parity(ctx.graph(), ref)
        d_blocking = ctx.result_digest()
        assert L.aos_set_host_wait(0, 0) == 0
        ctx.map_to_graph(pl, pts)
        assert ctx.result_digest() == d_blocking
        ctx.close()
    finally:
        L.aos_set_host_wait(0, 0)


def test_inflation_radius_beyond_the_stencil(gpu_ctx, oracle):
    """R = 80 cells (> 64): the seed stage switches to the EDT threshold for applyInflation; still bit-exact."""
    spec = synth.OrchardSpec(extent_x=30.0, extent_y=20.0, row_pitch=7.0, tree_spacing=4.0, n_points=60_000, outlier_count=2,
                             seed=4, grid_resolution=0.02, inflation_radius=1.6)
    pts = synth.make_orchard(spec)
    po, pl = params_pair(spec, oracle)
    r = oracle.seed_stage(po, pts)
    gpu_ctx.seed_stage(pl, pts)
    from helpers import assert_seed_parity
    assert_seed_parity(gpu_ctx, r)


def test_cpp_example_runs(tmp_path, oracle):
    """The C++ caller of examples/ (no Python in the loop) reports the oracle's counts."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "map_to_graph"
    subprocess.check_call(["g++", "-std=c++17", "-I", os.path.join(root, "include"), os.path.join(root, "examples", "map_to_graph.cpp"),
                           "-L", os.path.dirname(lib.LIB_PATH), "-laos_gpu", "-Wl,-rpath," + os.path.dirname(lib.LIB_PATH), "-o", str(exe)])
    spec = synth.config("SMALL", seed=1)
    pts = synth.make_orchard(spec)
    cloud = tmp_path / "cloud.bin"
    pts.tofile(cloud)
    po, _ = params_pair(spec, oracle)
    r = oracle.seed_stage(po, pts)
    ref = _oracle_graph(oracle, r)
    x0, y0 = spec.polygon[0]
    x1, y1 = spec.polygon[2]
    out = subprocess.run([str(exe), str(cloud), str(x0), str(y0), str(x1), str(y1)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert f"grid {r['w']} x {r['h']}" in out.stdout
    assert f"{r['n_clusters']} clusters, {r['n_rows']} tree rows" in out.stdout
    assert f"GvdGraph: {len(ref['nodes'])} nodes" in out.stdout and f"{len(ref['edges'])} edges" in out.stdout


def test_long_rows_graph(gpu_ctx, oracle):
    """1 km rows: endpoint rays run for thousands of 0.1 m steps (the warp-parallel replay of `cur += 0.1`), float32
    cluster sums pass 2^24 (BFS-order replay), and the graph has a few thousand nodes -- all bit-exact."""
    spec = synth.OrchardSpec(extent_x=1000.0, extent_y=16.0, row_pitch=4.0, n_points=500_000, gap_prob=0.01,
                             jitter=0.05, outlier_count=4, seed=12)
    pts = synth.make_orchard(spec)
    po, pl = params_pair(spec, oracle)
    r = oracle.seed_stage(po, pts)
    assert r["cl_sumx"].max() >= (1 << 24) and r["n_rows"] >= 4
    ref = _oracle_graph(oracle, r)
    gpu_ctx.map_to_graph(pl, pts)
    from helpers import assert_seed_selection_parity
    assert_seed_selection_parity(gpu_ctx, r)
    assert_graph_parity(gpu_ctx.graph(), ref)
    assert len(ref["nodes"]) > 2000


@pytest.mark.parametrize("rot,seed", [(17.0, 0), (45.0, 1), (90.0, 2), (-63.0, 3)])
def test_rotated_orchards(gpu_ctx, oracle, rot, seed):
    """Rows that are not parallel to the grid: diagonal skeletons (8-connected steps, BFS order), rays and row end
    points in arbitrary directions, Voronoi cells of slanted seed lines -- the whole path against the oracle."""
    spec = synth.OrchardSpec(extent_x=44.0, extent_y=40.0, row_pitch=5.0, n_points=260_000, outlier_count=5, seed=seed,
                             rotation_deg=rot)
    pts = synth.make_orchard(spec)
    po, pl = params_pair(spec, oracle)
    r = oracle.seed_stage(po, pts)
    assert r["n_rows"] >= 3
    ref = _oracle_graph(oracle, r)
    gpu_ctx.map_to_graph(pl, pts)
    from helpers import assert_seed_parity
    assert_seed_parity(gpu_ctx, r)
    assert_graph_parity(gpu_ctx.graph(), ref)
