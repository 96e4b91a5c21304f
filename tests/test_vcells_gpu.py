"""The opt-in device Voronoi (aos_set_voronoi_mode(AOS_VORONOI_DEVICE), csrc/k_vcells.cu) against the default Subdiv2D
replay on the same maps.  The device path builds the same Voronoi diagram in parallel but cannot reproduce Subdiv2D's
history-dependent choices (circumcentre edge pair, facet start), so the contract here is NOT bit-exactness: the node sets
must agree within 1e-4 m for all but a small fraction of nodes (the winners of the 5 cm first-come merge), the edges
must agree under that matching, the labelled corner nodes must be the same points, and the path must be deterministic.
The measured fractions are printed (pytest -s) and recorded in DESIGN.md."""
import numpy as np
import pytest

from aos_gpu import lib, synth
from aos_gpu.compare import compare_graphs

pytestmark = pytest.mark.gpu


def run_both(spec, pts):
    params = lib.SeedParams(grid_resolution=spec.grid_resolution, inflation_radius=spec.inflation_radius, polygon=spec.polygon,
                            exclusion=spec.exclusion)
    ctx = lib.Context(0)
    try:
        ctx.map_to_graph(params, pts)
        g_ref = ctx.graph()
        ctx.set_voronoi_mode(True)
        ctx.map_to_graph(params, pts)
        g_dev = ctx.graph()
        d1 = ctx.result_digest()
        ctx.map_to_graph(params, pts)
        d2 = ctx.result_digest()
        stages = dict(ctx.stage_times()) if hasattr(ctx, "stage_times") else {}
        ctx.set_voronoi_mode(False)
        ctx.map_to_graph(params, pts)
        g_back = ctx.graph()
    finally:
        ctx.close()
    assert d1 == d2, "device Voronoi is not deterministic"
    for k in ("nodes", "edges", "node_labels"):   # switching back restores the bit-exact path
        assert np.array_equal(g_ref[k], g_back[k])
    return g_ref, g_dev, stages


@pytest.mark.parametrize("name,seed,npts", [("SMALL", 1, None), ("C2", 0, 600_000), ("C2", 3, None), ("C1", 0, 400_000)])
def test_device_voronoi_agrees_with_replay(name, seed, npts):
    spec = synth.config(name, seed=seed, n_points=npts)
    g_ref, g_dev, _ = run_both(spec, synth.make_orchard(spec))
    r = compare_graphs(g_ref, g_dev)
    print(name, seed, r)
    assert r["ref_nodes_matched"] >= 0.95 and r["dev_nodes_matched"] >= 0.95
    assert r["ref_edges_found"] >= 0.90
    assert r["labelled_same"] >= 0.9
    assert abs(r["nodes_dev"] - r["nodes_ref"]) <= 0.03 * r["nodes_ref"] + 3


def test_device_voronoi_rotated_and_reference_polygon():
    spec = synth.OrchardSpec(extent_x=60.0, extent_y=50.0, row_pitch=5.0, n_points=350_000, seed=3, rotation_deg=17.0)
    g_ref, g_dev, _ = run_both(spec, synth.make_orchard(spec))
    r = compare_graphs(g_ref, g_dev)
    print("rotated", r)
    assert r["ref_nodes_matched"] >= 0.95 and r["ref_edges_found"] >= 0.90
    spec = synth.OrchardSpec(extent_x=77.0, extent_y=14.0, origin_x=-4.5, origin_y=-2.4, row_pitch=3.5, n_points=300_000, seed=5,
                             exclusion=synth.REFERENCE_EXCLUSION_DISCS)
    params = lib.SeedParams(grid_resolution=0.05, inflation_radius=0.8, polygon=synth.REFERENCE_POLYGON,
                            exclusion=synth.REFERENCE_EXCLUSION_DISCS)
    ctx = lib.Context(0)
    try:
        pts = synth.make_orchard(spec)
        ctx.map_to_graph(params, pts)
        g_ref = ctx.graph()
        ctx.set_voronoi_mode(True)
        ctx.map_to_graph(params, pts)
        g_dev = ctx.graph()
    finally:
        ctx.close()
    r = compare_graphs(g_ref, g_dev)
    print("reference polygon", r)
    assert r["ref_nodes_matched"] >= 0.95 and r["ref_edges_found"] >= 0.90


def test_device_voronoi_full_size():
    """config 3 grid with 40 M points: divergence and the stage time at map size"""
    spec = synth.config("C3", seed=3, n_points=40_000_000)
    pts = synth.make_orchard_strips(spec)
    params = lib.SeedParams(grid_resolution=spec.grid_resolution, inflation_radius=spec.inflation_radius, polygon=spec.polygon)
    ctx = lib.Context(0)
    try:
        ctx.set_profiling(True)
        ctx.map_to_graph(params, pts)
        g_ref = ctx.graph()
        t_ref = dict(ctx.stage_times())
        ctx.set_voronoi_mode(True)
        ctx.map_to_graph(params, pts)
        ctx.set_profiling(True)
        ctx.map_to_graph(params, pts)
        g_dev = ctx.graph()
        t_dev = dict(ctx.stage_times())
    finally:
        ctx.close()
    r = compare_graphs(g_ref, g_dev)
    print("C3 grid:", r, "device voronoi ms:", t_dev.get("gvd_device_voronoi"), "host replay ms:", t_ref.get("gvd_host_voronoi"))
    # at 1 km one float32 ulp is 6e-5 m, so the 1e-4 m match also counts plain rounding differences; what must hold at any
    # size: (almost) the same number of nodes, every reference node has a device node within the 5 cm merge radius, and the
    # large majority coincide to a millimetre
    assert abs(r["nodes_dev"] - r["nodes_ref"]) <= 1e-3 * r["nodes_ref"]
    assert r["worst_unmatched_m"] <= 0.0505
    assert r["ref_nodes_within_1mm"] >= 0.85 and r["labelled_same"] >= 0.9
