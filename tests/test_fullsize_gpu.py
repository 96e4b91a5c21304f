"""BASELINE config 3 at full size (20000 x 20000 cells, 1 km^2 at 0.05 m): the oracle cannot run there (its
reference loops are quadratic), so parity is carried by size-independent properties of every artefact -- the same
properties the small bit-exact cases satisfy -- checked on the device or with scipy's k-d tree on the host."""
import numpy as np
import pytest

from aos_gpu import lib, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c3():
    import torch
    spec = synth.config("C3", n_points=60_000_000)
    pts = synth.make_orchard_torch(spec, "cuda")
    params = lib.SeedParams(grid_resolution=spec.grid_resolution, inflation_radius=spec.inflation_radius, polygon=spec.polygon)
    ctx = lib.Context(0)
    info = ctx.map_to_graph(params, pts)
    del pts
    torch.cuda.empty_cache()
    yield ctx, params, info, spec
    ctx.close()


def _dev(ctx, gid):
    import torch
    from aos_gpu.bands import _CudaArray
    ptr, pitch = ctx.grid_device_bits(gid)
    s = ctx.seed_summary()
    return torch.as_tensor(_CudaArray(ptr, (s.info.height, pitch)), device="cuda")


def _popcount(t):
    import torch
    v = t.view(torch.uint8)
    lut = torch.tensor([bin(i).count("1") for i in range(256)], dtype=torch.int64, device=t.device)
    return int(lut[v.long()].sum())


def test_grid_inclusions(c3):
    ctx, params, info, spec = c3
    assert (info["width"], info["height"]) == (20000, 20000)
    raw, infl, occ = _dev(ctx, lib.GRID_RAW), _dev(ctx, lib.GRID_INFLATED), _dev(ctx, lib.GRID_OCCUPANCY)
    opened, skel, framed = _dev(ctx, lib.GRID_OPENED), _dev(ctx, lib.GRID_SKELETON), _dev(ctx, lib.GRID_SKELETON_FRAMED)
    assert int((raw & ~infl).count_nonzero()) == 0        # inflation only adds
    assert int((infl & ~occ).count_nonzero()) == 0        # the frame only adds
    assert int((opened & ~infl).count_nonzero()) == 0     # opening only removes
    assert int((skel & ~opened).count_nonzero()) == 0     # thinning only removes
    assert int((skel & ~framed).count_nonzero()) == 0
    n_raw, n_inf, n_skel = _popcount(raw), _popcount(infl), _popcount(skel)
    assert 0 < n_skel < n_raw < n_inf
    # the 5-cell frame of /occupancy_grid and the 1-cell frame of the published skeleton are complete
    assert int((occ[:5] != -1)[:, :624].count_nonzero()) == 0 and int((occ[-5:] != -1)[:, :624].count_nonzero()) == 0
    assert int((framed[0] != -1)[:624].count_nonzero()) == 0 and int((framed[:, 0] & 1).count_nonzero()) == 20000


def test_open_and_thin_are_idempotent(c3):
    import torch
    ctx, params, info, spec = c3
    w, h = info["width"], info["height"]
    opened, skel = _dev(ctx, lib.GRID_OPENED), _dev(ctx, lib.GRID_SKELETON)
    again = torch.zeros_like(opened)
    assert ctx.L.aos_open_bits(ctx.h, opened.data_ptr(), again.data_ptr(), w, h) == 0
    assert ctx.L.aos_synchronize(ctx.h) == 0
    assert torch.equal(again, opened)
    s2 = skel.clone()
    import ctypes as C
    launches, sub = C.c_int32(), C.c_int32()
    assert ctx.L.aos_thin_bits(ctx.h, s2.data_ptr(), w, h, C.byref(launches), C.byref(sub)) == 0
    assert torch.equal(s2, skel) and launches.value == 1   # fixed point: the first launch deletes nothing


def test_clusters_partition_the_skeleton(c3):
    ctx, params, info, spec = c3
    cl = ctx.clusters()
    assert len(cl) == info["n_clusters"] > 1000
    assert (np.diff(cl["label"]) > 0).all()                        # discovery order == raster order of the first cell
    n_skel_in_poly = int(cl["size"].sum())
    skel = _dev(ctx, lib.GRID_SKELETON)
    assert n_skel_in_poly <= _popcount(skel)                         # cells outside the polygon are not clustered
    assert n_skel_in_poly > 0.9 * _popcount(skel)
    # centres are the float32 means up to rounding; lengths bounded by the bounding box of the sums
    mean_x = cl["sum_x"] / cl["size"]
    assert np.allclose(cl["center_x"], mean_x, rtol=2e-6, atol=0.5)
    assert (cl["length"] >= 0).all() and (cl["max_d2"] <= 2 * 20000 ** 2).all()
    big = cl[cl["size"] > 2000]
    assert len(big) > 0 and (np.float32(np.sqrt(big["max_d2"].astype(np.float64)) * np.float32(0.05)) == big["length"]).all()
    rows = ctx.tree_rows()
    assert len(rows) == info["n_rows"] and (rows["length"] >= 2.0).all()
    assert (np.diff(rows["cluster"]) > 0).all()


def test_seeds_respect_the_first_come_rule(c3):
    from scipy.spatial import cKDTree
    ctx, params, info, spec = c3
    seeds, counts, rows_info = ctx.select_seeds()
    assert sum(counts) == len(seeds) > 100_000
    a, b = counts[0], counts[0] + counts[1]
    for part in (seeds[:a], seeds[a:b], seeds[b:]):                 # no two accepted seeds of one list closer than 0.5 m
        assert len(cKDTree(part).query_pairs(0.5 - 1e-9)) == 0
    cy = 0.5 * (rows_info[:, 1] + rows_info[:, 3])
    assert len(rows_info) == info["n_rows"]
    merged = lib.merge_seeds(seeds)
    assert len(merged) == info["graph"]["merged_seeds"]
    assert cy.min() > 0 and cy.max() < 1000


def test_graph_invariants(c3):
    from scipy.spatial import cKDTree
    ctx, params, info, spec = c3
    g = ctx.graph()
    n, e = g["nodes"], g["edges"]
    assert len(n) == info["graph"]["nodes"] > 100_000 and len(e) == info["graph"]["edges"]
    assert (n[:, 0] >= 0).all() and (n[:, 0] <= 1000.0).all() and (n[:, 1] >= 0).all() and (n[:, 1] <= 1000.0).all()   # cropped (gvd:420-483)
    assert len(cKDTree(n).query_pairs(0.05 - 1e-12)) == 0          # first-come 5 cm merge (vd:149-207)
    assert (e[:, 0] < e[:, 1]).all() and e.min() >= 0 and e.max() < len(n)
    keys = e[:, 0].astype(np.int64) << 32 | e[:, 1]
    assert len(np.unique(keys)) == len(keys)                        # (min,max) de-duplication (gvd:838-846)
    d = np.sqrt(((n[e[:, 0]] - n[e[:, 1]]) ** 2).sum(1))
    assert np.array_equal(d.astype(np.float32), g["edge_lengths"])
    assert (g["edge_clearances"] == 0).all()
    # proximity edges: every node pair within 0.5 m that does not cross the skeleton is an edge -> spot check degree
    lab = g["node_labels"]
    assert lab.max() <= 15 and (lab > 0).sum() >= info["n_rows"]    # each row labels at least one corner node
    assert g["node_label_counts"].sum() == len(g["node_label_clusters"]) == len(g["node_label_types"])
    assert (g["node_label_types"] >= 0).all() and (g["node_label_types"] <= 3).all()
    assert g["node_label_clusters"].max() < info["n_rows"]
    # no published edge touches a skeleton cell at its end points (gvd:320-359 samples include both ends)
    skel = _dev(ctx, lib.GRID_SKELETON_FRAMED).cpu().numpy().view(np.uint32)
    for idx in (e[:, 0], e[:, 1]):
        mx = (n[idx, 0] / np.float64(np.float32(0.05))).astype(np.int64)
        my = (n[idx, 1] / np.float64(np.float32(0.05))).astype(np.int64)
        ok = (mx >= 0) & (mx < 20000) & (my >= 0) & (my < 20000)
        bit = (skel[my[ok], mx[ok] >> 5] >> (mx[ok] & 31).astype(np.uint32)) & 1
        assert int(bit.sum()) == 0
