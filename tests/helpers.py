"""Shared helpers of the parity tests."""
import numpy as np

from aos_gpu import lib, synth


def params_pair(spec, oracle_mod, **over):
    """(oracle SeedParams, library SeedParams) for an OrchardSpec."""
    kw = dict(grid_resolution=spec.grid_resolution, inflation_radius=spec.inflation_radius,
              polygon=spec.polygon, exclusion=spec.exclusion)
    kw.update(over)
    return oracle_mod.SeedParams(**kw), lib.SeedParams(**kw)


def canonical_clusters_oracle(r):
    """Oracle cluster table keyed by canonical label (min linear index)."""
    return {int(r["cl_first"][i]): dict(size=int(r["cl_size"][i]), sumx=int(r["cl_sumx"][i]), sumy=int(r["cl_sumy"][i]),
                                        cx=np.float32(r["cl_cx"][i]), cy=np.float32(r["cl_cy"][i]),
                                        maxd2=int(r["cl_maxd2"][i]), length=np.float32(r["cl_len"][i]))
            for i in range(r["n_clusters"])}


def assert_seed_parity(ctx, r, check_labels=True):
    """Bit-exact comparison of every seed-stage artefact of the library against the oracle dict r."""
    names = [(lib.GRID_RAW, "occ_raw"), (lib.GRID_INFLATED, "occ_inflated"), (lib.GRID_OCCUPANCY, "occ_border"),
             (lib.GRID_OPENED, "opened"), (lib.GRID_SKELETON, "skel"), (lib.GRID_SKELETON_FRAMED, "skel_framed")]
    s = ctx.seed_summary()
    assert (s.info.width, s.info.height) == (r["w"], r["h"])
    assert s.info.origin_x == r["origin_x"] and s.info.origin_y == r["origin_y"]
    for gid, key in names:
        got = ctx.grid_int8(gid)
        assert got.shape == r[key].shape
        diff = np.argwhere(got != r[key])
        assert len(diff) == 0, f"{key}: {len(diff)} cells differ, first at (y,x)={diff[:5].tolist()}"
    cl = ctx.clusters()
    ref = canonical_clusters_oracle(r)
    assert len(cl) == len(ref), f"clusters: {len(cl)} vs {len(ref)}"
    # discovery order == raster order of first cell
    assert [int(c["label"]) for c in cl] == [int(x) for x in r["cl_first"]]
    for c in cl:
        o = ref[int(c["label"])]
        assert int(c["size"]) == o["size"]
        assert int(c["sum_x"]) == o["sumx"] and int(c["sum_y"]) == o["sumy"]
        assert int(c["max_d2"]) == o["maxd2"], f"cluster {c['label']}: max_d2 {c['max_d2']} vs {o['maxd2']}"
        assert np.float32(c["length"]) == o["length"]
        assert np.float32(c["center_x"]) == o["cx"] and np.float32(c["center_y"]) == o["cy"], \
            f"cluster {c['label']}: centre ({c['center_x']},{c['center_y']}) vs ({o['cx']},{o['cy']})"
    if check_labels:
        lab = ctx.labels()
        assert np.array_equal(lab, r["labels"])
    rows = ctx.tree_rows()
    assert len(rows) == r["n_rows"]
    for i, row in enumerate(rows):
        o = r["rows"][i]
        got = [row["center_x"], row["center_y"], row["start_x"], row["start_y"], row["end_x"], row["end_y"], row["length"]]
        assert list(o) == got, f"row {i}: {got} vs {list(o)}"
        assert int(row["cluster"]) == int(r["row_cluster"][i])
    assert_seed_selection_parity(ctx, r)


def assert_seed_selection_parity(ctx, r):
    """Host seed selection (/voronoi_seeds, /exploration_tree_rows_info) against the oracle, bit-exact."""
    seeds, counts, rows_info = ctx.select_seeds()
    assert counts == (r["n_virtual"], r["n_ray"], r["n_endpoint"]), f"{counts} vs {(r['n_virtual'], r['n_ray'], r['n_endpoint'])}"
    assert seeds.shape == r["seeds"].shape
    assert np.array_equal(seeds, r["seeds"]), f"seeds differ at {np.argwhere(seeds != r['seeds'])[:5].tolist()}"
    assert np.array_equal(rows_info, r["rows_info"])


GRAPH_KEYS = ("node_labels", "node_cluster_indices", "node_label_counts", "node_label_clusters", "node_label_types",
              "edges", "edge_lengths", "edge_clearances")


def assert_graph_parity(got, ref):
    """GvdGraph.msg arrays of the library against the oracle: bit-exact, reference order."""
    assert got["n_merged_seeds"] == len(ref["merged_seeds"])
    assert got["n_voronoi_edges"] == ref["n_voro_edges"]
    assert got["n_boundary_points"] == ref["n_boundary_points_precrop"]
    assert got["nodes"].shape == ref["nodes"].shape, f"nodes {got['nodes'].shape} vs {ref['nodes'].shape}"
    assert np.array_equal(got["nodes"], ref["nodes"]), f"node coordinates differ at {np.argwhere(got['nodes'] != ref['nodes'])[:4].tolist()}"
    assert np.all(got["nodes_xyz"][:, 2] == 0.0)
    assert got["edges"].shape == ref["edges"].shape, f"edges {got['edges'].shape} vs {ref['edges'].shape}"
    if len(ref.get("corner_points", [])):
        assert np.array_equal(got["corner_points"], ref["corner_points"]), \
            f"corner points differ at rows {np.unique(np.argwhere(got['corner_points'] != ref['corner_points'])[:, 0])[:6].tolist()}"
    for k in GRAPH_KEYS:
        assert got[k].shape == ref[k].shape, f"{k}: {got[k].shape} vs {ref[k].shape}"
        assert np.array_equal(got[k], ref[k]), f"{k} differs at {np.argwhere(got[k] != ref[k])[:4].tolist()}"
    assert got["resolution"] == float(ref["resolution"]) and got["origin_x"] == ref["origin_x"] and got["origin_y"] == ref["origin_y"]
