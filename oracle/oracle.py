"""ctypes front-end of the C oracle (oracle/aos_oracle_{seed,gvd}.c).  TEST INFRASTRUCTURE ONLY.

`seed_stage` restates aos_seed_gen_node::processPointCloud (src/aos_seed_gen_node.cpp:452-579),
`gvd_stage` restates aos_gvd_node (src/aos_gvd_node.cpp:84-128,255-318) with the real
cv2.Subdiv2D standing in for cv::Subdiv2D (oracle/subdiv.py).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libaos_oracle.so")


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("aos_oracle_seed.c", "aos_oracle_gvd.c", "aos_oracle_fast.c", "aos_oracle_fast.h",
                                             "aos_oracle.h")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


class _SeedParams(C.Structure):
    _fields_ = [
        ("clipping_minz", C.c_float), ("clipping_maxz", C.c_float),
        ("clipping_minx", C.c_float), ("clipping_maxx", C.c_float),
        ("clipping_miny", C.c_float), ("clipping_maxy", C.c_float),
        ("grid_resolution", C.c_float), ("inflation_radius", C.c_float),
        ("cluster_min_length", C.c_double),
        ("n_poly", C.c_int), ("poly", C.POINTER(C.c_double)),
        ("n_excl", C.c_int), ("excl", C.POINTER(C.c_float)),
    ]


_P8 = C.POINTER(C.c_int8)
_P32 = C.POINTER(C.c_int32)
_P64 = C.POINTER(C.c_int64)
_PF = C.POINTER(C.c_float)
_PD = C.POINTER(C.c_double)


class _SeedResult(C.Structure):
    _fields_ = [
        ("w", C.c_int), ("h", C.c_int), ("origin_x", C.c_double), ("origin_y", C.c_double), ("res", C.c_float),
        ("occ_raw", _P8), ("occ_inflated", _P8), ("occ_border", _P8), ("opened", _P8), ("skel", _P8),
        ("skel_framed", _P8),
        ("n_clusters", C.c_int), ("labels", _P32), ("cl_first", _P32), ("cl_size", _P32),
        ("cl_sumx", _P64), ("cl_sumy", _P64), ("cl_cx", _PF), ("cl_cy", _PF), ("cl_maxd2", _P64),
        ("cl_len", _PF), ("cl_cell_off", _P32), ("cl_cells", _P32),
        ("n_rows", C.c_int), ("row_cluster", _P32), ("rows", _PD), ("rows_info", _PD),
        ("n_seeds", C.c_int), ("n_virtual", C.c_int), ("n_ray", C.c_int), ("n_endpoint", C.c_int),
        ("seeds", _PD),
    ]


class _Graph(C.Structure):
    _fields_ = [
        ("n_nodes", C.c_int), ("n_edges", C.c_int), ("nodes", _PD), ("node_labels", _P32),
        ("node_cluster_indices", _P32), ("node_label_counts", _P32), ("n_label_entries", C.c_int),
        ("node_label_clusters", _P32), ("node_label_types", _P32), ("edges", _P32),
        ("edge_lengths", _PF), ("edge_clearances", _PF),
        ("n_voro_edges", C.c_int), ("n_boundary_points_precrop", C.c_int), ("corner_points", _PD),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_seed_stage.argtypes = [C.POINTER(_SeedParams), _PF, C.c_size_t, C.c_size_t, C.POINTER(_SeedResult)]
        _lib.orc_seed_stage.restype = C.c_int
        _lib.orc_gvd_merge_seeds.argtypes = [_PD, C.c_int, _PD]
        _lib.orc_gvd_merge_seeds.restype = C.c_int
        _lib.orc_trim_path.argtypes = [_PD, C.c_int, _P8, C.c_int, C.c_int, C.c_double, C.c_double, C.c_float, C.c_double]
        _lib.orc_trim_path.restype = C.c_int
        _lib.orc_gvd_graph.argtypes = [_PF, _P32, C.c_int, _P8, C.c_int, C.c_int, C.c_double, C.c_double,
                                       C.c_float, _PD, C.c_int, C.POINTER(_Graph)]
        _lib.orc_gvd_graph.restype = C.c_int
        _lib.orc_point_in_polygon.argtypes = [C.c_double, C.c_double, _PD, C.c_int]
        _lib.orc_point_in_polygon.restype = C.c_int
        _lib.orc_thin_zhangsuen.argtypes = [_P8, C.c_int, C.c_int]
        _lib.orc_thin_zhangsuen.restype = C.c_int
        _lib.orc_set_fast.argtypes = [C.c_int, C.c_int, C.c_int]
        for name in ("orc_inflate",):
            getattr(_lib, name).argtypes = [_P8, C.c_int, C.c_int, C.c_int, _P8]
        for name in ("orc_mark_borders", "orc_open_cross"):
            getattr(_lib, name).argtypes = [_P8, C.c_int, C.c_int, _P8]
    return _lib


@dataclass
class SeedParams:
    """Mirror of the aos_seed_gen_node parameters that reach the path (config/aos_planner_params.yaml)."""
    clipping_minz: float = -0.4
    clipping_maxz: float = 0.5
    clipping_minx: float = -5.0
    clipping_maxx: float = 72.0
    clipping_miny: float = -10.0
    clipping_maxy: float = 20.0
    grid_resolution: float = 0.05
    inflation_radius: float = 0.8
    cluster_min_length: float = 2.0
    polygon: np.ndarray = field(default_factory=lambda: np.zeros((0, 2)))
    exclusion: np.ndarray = field(default_factory=lambda: np.zeros((0, 3), np.float32))

    def to_c(self):
        poly = np.ascontiguousarray(self.polygon, dtype=np.float64).reshape(-1, 2)
        excl = np.ascontiguousarray(self.exclusion, dtype=np.float32).reshape(-1, 3)
        p = _SeedParams(self.clipping_minz, self.clipping_maxz, self.clipping_minx, self.clipping_maxx,
                        self.clipping_miny, self.clipping_maxy, self.grid_resolution, self.inflation_radius,
                        self.cluster_min_length, len(poly), poly.ctypes.data_as(_PD), len(excl),
                        excl.ctypes.data_as(_PF))
        p._keep = (poly, excl)
        return p


def set_fast(on: bool, threads: int = 0, skip_labels: bool = False) -> None:
    """Switch the oracle to its indexed / multi-threaded loops (identical results; see aos_oracle_fast.h)."""
    if threads <= 0:
        threads = os.cpu_count() or 1
    lib().orc_set_fast(int(on), int(threads), int(skip_labels))


def _arr(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)


def seed_stage(params: SeedParams, points: np.ndarray) -> dict:
    """points: float32 [N, k>=3] (x, y, z, ...).  Returns every intermediate artefact (copied)."""
    pts = np.ascontiguousarray(points, dtype=np.float32)
    assert pts.ndim == 2 and pts.shape[1] >= 3
    cp = params.to_c()
    r = _SeedResult()
    rc = lib().orc_seed_stage(C.byref(cp), pts.ctypes.data_as(_PF), pts.shape[0], pts.shape[1], C.byref(r))
    assert rc == 0
    n = r.w * r.h
    out = dict(w=r.w, h=r.h, origin_x=r.origin_x, origin_y=r.origin_y, res=np.float32(r.res))
    for k in ("occ_raw", "occ_inflated", "occ_border", "opened", "skel", "skel_framed"):
        out[k] = _arr(getattr(r, k), n, np.int8).reshape(r.h, r.w)
    out["labels"] = _arr(r.labels, n, np.int32).reshape(r.h, r.w) if r.labels else None
    nc = r.n_clusters
    out["n_clusters"] = nc
    for k, dt in (("cl_first", np.int32), ("cl_size", np.int32), ("cl_sumx", np.int64), ("cl_sumy", np.int64),
                  ("cl_cx", np.float32), ("cl_cy", np.float32), ("cl_maxd2", np.int64), ("cl_len", np.float32)):
        out[k] = _arr(getattr(r, k), nc, dt)
    out["cl_cell_off"] = _arr(r.cl_cell_off, nc + 1, np.int32)
    out["cl_cells"] = _arr(r.cl_cells, int(out["cl_cell_off"][-1]) if nc else 0, np.int32)
    out["n_rows"] = r.n_rows
    out["row_cluster"] = _arr(r.row_cluster, r.n_rows, np.int32)
    out["rows"] = _arr(r.rows, 7 * r.n_rows, np.float64).reshape(-1, 7)
    out["rows_info"] = _arr(r.rows_info, 4 * r.n_rows, np.float64).reshape(-1, 4)
    out["seeds"] = _arr(r.seeds, 2 * r.n_seeds, np.float64).reshape(-1, 2)
    out["n_virtual"], out["n_ray"], out["n_endpoint"] = r.n_virtual, r.n_ray, r.n_endpoint
    lib().orc_seed_result_free(C.byref(r))
    return out


def merge_seeds(seeds: np.ndarray) -> np.ndarray:
    s = np.ascontiguousarray(seeds, dtype=np.float64).reshape(-1, 2)
    out = np.zeros_like(s)
    m = lib().orc_gvd_merge_seeds(s.ctypes.data_as(_PD), len(s), out.ctypes.data_as(_PD))
    return out[:m].copy()


def trim_path(path_xy: np.ndarray, grid: np.ndarray, origin_x: float, origin_y: float, res, safety_distance: float = 0.2) -> int:
    """trimPathNearOccupiedRegions (path_gen:1570-1630): number of poses that remain."""
    p = np.ascontiguousarray(path_xy, dtype=np.float64).reshape(-1, 2)
    g = np.ascontiguousarray(grid, dtype=np.int8)
    h, w = g.shape
    return int(lib().orc_trim_path(p.ctypes.data_as(_PD), len(p), g.ctypes.data_as(_P8), w, h, float(origin_x), float(origin_y),
                                   float(np.float32(res)), float(safety_distance)))


def gvd_stage(seeds: np.ndarray, skel_framed: np.ndarray, origin_x: float, origin_y: float, res,
              rows_info: np.ndarray) -> dict:
    """seeds: the /voronoi_seeds PoseArray positions (un-merged).  Returns GvdGraph.msg arrays."""
    from . import subdiv

    skel = np.ascontiguousarray(skel_framed, dtype=np.int8)
    h, w = skel.shape
    res = np.float32(res)
    merged = merge_seeds(seeds)
    merged = merged[np.isfinite(merged).all(axis=1)]  # gvd:266-270
    g = dict(merged_seeds=merged)
    if len(merged) == 0:
        return g
    minx, miny = float(origin_x), float(origin_y)
    maxx = minx + float(np.float32(np.float32(w) * res))  # gvd:279 uint32*float -> float
    maxy = miny + float(np.float32(np.float32(h) * res))
    facets_xy, facet_off, sub_pts = subdiv.voronoi_facets(merged, minx, maxx, miny, maxy)
    g["subdiv_points"] = sub_pts
    g["facets_xy"], g["facet_off"] = facets_xy, facet_off
    rows = np.ascontiguousarray(rows_info, dtype=np.float64).reshape(-1, 4)
    gr = _Graph()
    rc = lib().orc_gvd_graph(facets_xy.ctypes.data_as(_PF), facet_off.ctypes.data_as(_P32), len(facet_off) - 1,
                             skel.ctypes.data_as(_P8), w, h, minx, miny, C.c_float(float(res)),
                             rows.ctypes.data_as(_PD), len(rows), C.byref(gr))
    assert rc == 0
    g.update(
        resolution=float(res), origin_x=minx, origin_y=miny,
        nodes=_arr(gr.nodes, 2 * gr.n_nodes, np.float64).reshape(-1, 2),
        node_labels=_arr(gr.node_labels, gr.n_nodes, np.int32),
        node_cluster_indices=_arr(gr.node_cluster_indices, gr.n_nodes, np.int32),
        node_label_counts=_arr(gr.node_label_counts, gr.n_nodes, np.int32),
        node_label_clusters=_arr(gr.node_label_clusters, gr.n_label_entries, np.int32),
        node_label_types=_arr(gr.node_label_types, gr.n_label_entries, np.int32),
        edges=_arr(gr.edges, 2 * gr.n_edges, np.int32).reshape(-1, 2),
        edge_lengths=_arr(gr.edge_lengths, gr.n_edges, np.float32),
        edge_clearances=_arr(gr.edge_clearances, gr.n_edges, np.float32),
        corner_points=_arr(gr.corner_points, 8 * len(rows) if gr.n_nodes else 0, np.float64).reshape(-1, 4, 2),
        n_voro_edges=gr.n_voro_edges, n_boundary_points_precrop=gr.n_boundary_points_precrop,
    )
    lib().orc_graph_free(C.byref(gr))
    return g


def radius_outlier_removal(points: np.ndarray, radius: float = 0.2, min_neighbors: int = 2) -> np.ndarray:
    """pcl::RadiusOutlierRemoval as globalMapCallback uses it (src/aos_seed_gen_node.cpp:229-248), dense-cloud branch
    of PCL 1.12 restated (PCL is absent: PARITY UNPINNED): keep a point iff the (min_neighbors + 1)-th nearest
    neighbour -- the query point is its own nearest -- has squared distance <= radius^2, the distance being FLANN's
    L2_Simple<float>: ((dx*dx + dy*dy) + dz*dz) in float32, compared in double.  Brute force, O(N^2): small clouds
    only.  Returns the boolean keep mask; non-finite points are dropped."""
    p = np.ascontiguousarray(points[:, :3], dtype=np.float32)
    n = len(p)
    keep = np.zeros(n, bool)
    fin = np.isfinite(p).all(axis=1)
    r2 = float(np.float32(radius)) ** 2 if isinstance(radius, np.floating) else float(radius) * float(radius)
    idx = np.nonzero(fin)[0]
    q = p[idx]
    for a in range(0, len(q), 512):
        blk = q[a:a + 512]
        dx = blk[:, None, 0] - q[None, :, 0]
        dy = blk[:, None, 1] - q[None, :, 1]
        dz = blk[:, None, 2] - q[None, :, 2]
        d2 = (dx * dx + dy * dy) + dz * dz            # float32 throughout
        cnt = (d2.astype(np.float64) <= r2).sum(axis=1) - 1   # minus the query itself
        keep[idx[a:a + 512]] = cnt >= min_neighbors
    return keep


def _pack_bits(img: np.ndarray) -> np.ndarray:
    """{0,100} int8 [H, W] -> uint32 [H, pitch]: 32 cells per word, LSB = lowest x, pitch padded to 4 words (the
    bit-grid layout include/aos_gpu.h documents for AOS_FMT_BITS)."""
    h, w = img.shape
    pitch = (((w + 31) >> 5) + 3) & ~3
    out = np.zeros((h, pitch * 4), np.uint8)
    for r0 in range(0, h, 2048):   # in slabs: the unpacked pad of a 20000^2 grid would be 400 MB
        blk = img[r0:r0 + 2048] == 100
        out[r0:r0 + 2048, :(w + 7) // 8] = np.packbits(blk, axis=1, bitorder="little")
    return out.view(np.uint32).reshape(h, pitch)


def result_artefacts(r: dict, g: dict | None) -> dict:
    """The oracle's results under the names aos_gpu.lib.digest_of() hashes (seed_stage dict r, gvd_stage dict g)."""
    art = {"occupancy": _pack_bits(r["occ_border"]), "skeleton": _pack_bits(r["skel"]),
           "skeleton_framed": _pack_bits(r["skel_framed"]),
           "clusters.label": r["cl_first"], "clusters.size": r["cl_size"], "clusters.center_x": r["cl_cx"],
           "clusters.center_y": r["cl_cy"], "clusters.length": r["cl_len"], "clusters.sum_x": r["cl_sumx"],
           "clusters.sum_y": r["cl_sumy"], "clusters.max_d2": r["cl_maxd2"],
           "rows.cluster": r["row_cluster"], "seeds": r["seeds"], "rows_info": r["rows_info"]}
    for i, f in enumerate(("center_x", "center_y", "start_x", "start_y", "end_x", "end_y", "length")):
        art["rows." + f] = r["rows"][:, i] if len(r["rows"]) else np.zeros(0)
    have = g is not None and "nodes" in g
    if have:
        xyz = np.zeros((len(g["nodes"]), 3))
        xyz[:, :2] = g["nodes"]
    for k in ("nodes_xyz", "node_labels", "node_cluster_indices", "node_label_counts", "node_label_clusters", "node_label_types",
              "edges", "edge_lengths", "edge_clearances"):
        art["graph." + k] = None if not have else (xyz if k == "nodes_xyz" else g[k])
    return art
