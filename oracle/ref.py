"""oracle/_ref front-end: the reference's OWN node classes, compiled from /root/reference against the stand-in headers
of oracle/ref_shim/ (recipe: oracle/Makefile target `_ref`), driven in-process.  TEST INFRASTRUCTURE ONLY.

What is the reference's code and what is not:
  * everything in src/aos_seed_gen_node.cpp, src/aos_gvd_node.cpp, src/utils/voronoi_diagram.cpp runs as compiled from
    those files (processPointCloud, generateOccupancyGrid, applyInflation, markBoundariesAsOccupied,
    skeletonizeOccupancyGrid, clusterOccupiedCells, convertClustersToTreeRows, generateVirtualSeeds, ray casts,
    voronoiSeedsCallback, VoronoiDiagram::compute / extractBoundaryPoints, buildGraphFromBoundaryPoints, the edge test,
    the crop, the TL/TR/BL/BR search, publishGraph ...);
  * third-party calls go through hooks installed here: cv::morphologyEx / getStructuringElement and cv::Subdiv2D run
    the REAL OpenCV of this image (cv2 4.13); cv::ximgproc::thinning (opencv_contrib) and pcl::RadiusOutlierRemoval
    are absent from the image and run the oracle's restatements (parity unpinned for those two, as everywhere);
    pcl::PassThrough and Eigen::Vector2d are restated inside the stand-in headers.

`/root/reference` exists only in the authoring container: the library is built there and travels to the GPU box as a
prebuilt file (oracle/_ref/ is git-ignored, not gpurun-ignored).  available() is False where it is missing.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libaos_ref.so")
REFERENCE_ROOT = "/root/reference"

_P8 = C.POINTER(C.c_int8)
_PU8 = C.POINTER(C.c_uint8)
_P32 = C.POINTER(C.c_int32)
_PF = C.POINTER(C.c_float)
_PD = C.POINTER(C.c_double)
_PI = C.POINTER(C.c_int)

MORPH_HOOK = C.CFUNCTYPE(None, _PU8, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _PU8)
THIN_HOOK = C.CFUNCTYPE(None, _PU8, C.c_int, C.c_int, C.c_int, _PU8)
SUBDIV_HOOK = C.CFUNCTYPE(None, _PI, _PF, C.c_int, _PI, _PI, _PI, _PF, _PF)
ROR_HOOK = C.CFUNCTYPE(None, _PF, C.c_int, C.c_double, C.c_int, _PU8)


class _SeedParams(C.Structure):
    _fields_ = [("clipping_minz", C.c_float), ("clipping_maxz", C.c_float), ("grid_resolution", C.c_float),
                ("inflation_radius", C.c_float), ("cluster_min_length", C.c_double), ("n_poly", C.c_int), ("poly", _PD)]


class _SeedResult(C.Structure):
    _fields_ = [("w", C.c_int), ("h", C.c_int), ("origin_x", C.c_double), ("origin_y", C.c_double), ("res", C.c_float),
                ("occ_border", _P8), ("skel", _P8), ("skel_framed", _P8), ("occ_raw_nodisc", _P8),
                ("n_seeds", C.c_int), ("seeds", _PD), ("n_rows_info", C.c_int), ("rows_info", _PD),
                ("n_cluster_info", C.c_int), ("cluster_info", _PD),
                ("n_clusters", C.c_int), ("cl_size", _P32), ("cl_first", _P32), ("cl_cx", _PF), ("cl_cy", _PF),
                ("cl_len", _PF), ("cl_cell_off", _P32), ("cl_cells", _P32), ("n_after_ror", C.c_int)]


class _Graph(C.Structure):
    _fields_ = [("published", C.c_int), ("resolution", C.c_double), ("origin_x", C.c_double), ("origin_y", C.c_double),
                ("n_nodes", C.c_int), ("n_edges", C.c_int), ("nodes_xyz", _PD), ("node_labels", _P32),
                ("node_cluster_indices", _P32), ("node_label_counts", _P32), ("n_label_entries", C.c_int),
                ("node_label_clusters", _P32), ("node_label_types", _P32), ("edges", _P32), ("edge_lengths", _PF),
                ("edge_clearances", _PF), ("n_merged_seeds", C.c_int), ("merged_seeds", _PD), ("n_voro_edges", C.c_int),
                ("n_boundary_points_precrop", C.c_int)]


def build(force: bool = False) -> str | None:
    """Compile oracle/_ref/libaos_ref.so from the sources under /root/reference (only where that tree exists)."""
    if not os.path.isdir(REFERENCE_ROOT):
        return _SO if os.path.exists(_SO) else None
    subprocess.check_call(["make", "-C", _HERE, "-s", "_ref"] + (["-B"] if force else []))
    return _SO


def available() -> bool:
    return os.path.exists(_SO)


_lib = None
_keep = []


def _morph(src, rows, cols, op, shape, kw, kh, dst):
    import cv2
    a = np.ctypeslib.as_array(src, shape=(rows, cols))
    k = cv2.getStructuringElement(int(shape), (int(kw), int(kh)))
    out = cv2.morphologyEx(a, int(op), k)
    np.ctypeslib.as_array(dst, shape=(rows, cols))[:] = out


def _thin(src, rows, cols, typ, dst):
    from . import oracle as O
    assert typ == 0, "only THINNING_ZHANGSUEN is used by the reference"
    a = np.ctypeslib.as_array(src, shape=(rows, cols))
    g = np.where(a == 255, 100, 0).astype(np.int8)   # orc_thin_zhangsuen works on {0,100} grids
    O.lib().orc_thin_zhangsuen(g.ctypes.data_as(_P8), cols, rows)
    np.ctypeslib.as_array(dst, shape=(rows, cols))[:] = np.where(g == 100, 255, 0).astype(np.uint8)


_subdiv_cache = {}


def _subdiv(rect, pts, n, n_facets, n_vertices, sizes, xy, centers):
    import cv2
    key = (rect[0], rect[1], rect[2], rect[3], n, C.addressof(pts.contents) if n else 0)
    if not sizes:   # first call: run the real Subdiv2D, remember the result for the fill call
        p = np.ctypeslib.as_array(pts, shape=(n, 2)).copy() if n else np.zeros((0, 2), np.float32)
        sd = cv2.Subdiv2D((int(rect[0]), int(rect[1]), int(rect[2]), int(rect[3])))
        for x, y in p:
            try:
                sd.insert((float(x), float(y)))
            except cv2.error:
                continue   # vd:83-88
        facets, cen = sd.getVoronoiFacetList([])
        _subdiv_cache[key] = (facets, cen)
        n_facets[0] = len(facets)
        n_vertices[0] = int(sum(len(f) for f in facets))
        return
    facets, cen = _subdiv_cache.pop(key)
    k = 0
    for i, f in enumerate(facets):
        sizes[i] = len(f)
        f = np.asarray(f, np.float32).reshape(-1, 2)
        for j in range(len(f)):
            xy[2 * k] = f[j, 0]
            xy[2 * k + 1] = f[j, 1]
            k += 1
        centers[2 * i] = cen[i][0]
        centers[2 * i + 1] = cen[i][1]


def _ror(xyzp, n, radius, min_neighbors, keep):
    from . import oracle as O
    p = np.ctypeslib.as_array(xyzp, shape=(n, 4))
    k = O.radius_outlier_removal(p[:, :3], radius, min_neighbors)
    np.ctypeslib.as_array(keep, shape=(n,))[:] = k.astype(np.uint8)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libaos_ref.so is missing (built only where /root/reference exists)")
        L = C.CDLL(_SO)
        L.ref_seed_run.argtypes = [C.POINTER(_SeedParams), _PF, C.c_size_t, C.c_size_t, C.c_int, C.POINTER(_SeedResult)]
        L.ref_seed_run.restype = C.c_int
        L.ref_gvd_run.argtypes = [_PD, C.c_int, _P8, C.c_int, C.c_int, C.c_double, C.c_double, C.c_float, _PD, C.c_int,
                                  C.POINTER(_Graph)]
        L.ref_gvd_run.restype = C.c_int
        L.ref_step_inflate.argtypes = [_P8, C.c_int, C.c_int, C.c_float, C.c_float, _P8]
        L.ref_step_mark_borders.argtypes = [_P8, C.c_int, C.c_int, _P8]
        L.ref_step_skeletonize.argtypes = [_P8, C.c_int, C.c_int, _P8]
        L.ref_step_point_in_polygon.argtypes = [C.c_double, C.c_double, _PD, C.c_int]
        L.ref_step_point_in_polygon.restype = C.c_int
        if hasattr(L, "ref_trim_path"):
            L.ref_trim_path.argtypes = [_PD, C.c_int, _P8, C.c_int, C.c_int, C.c_double, C.c_double, C.c_float]
            L.ref_trim_path.restype = C.c_int
        hooks = (MORPH_HOOK(_morph), THIN_HOOK(_thin), SUBDIV_HOOK(_subdiv), ROR_HOOK(_ror))
        _keep.append(hooks)
        L.ref_set_hooks.argtypes = [MORPH_HOOK, THIN_HOOK, SUBDIV_HOOK, ROR_HOOK]
        L.ref_set_hooks(*hooks)
        _lib = L
    return _lib


def _arr(ptr, n, dt):
    if n == 0 or not ptr:
        return np.zeros(0, dt)
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dt, copy=True)


def seed_stage(params, points: np.ndarray, through_global_map_callback: bool = False) -> dict:
    """params: oracle.SeedParams.  NOTE the reference hard-codes its 11 exclusion discs (seed_gen:487-499): compare
    only with oracle / library runs that pass synth.REFERENCE_EXCLUSION_DISCS."""
    pts = np.ascontiguousarray(points, dtype=np.float32)
    poly = np.ascontiguousarray(params.polygon, np.float64).reshape(-1, 2)
    cp = _SeedParams(params.clipping_minz, params.clipping_maxz, params.grid_resolution, params.inflation_radius,
                     params.cluster_min_length, len(poly), poly.ctypes.data_as(_PD))
    r = _SeedResult()
    rc = lib().ref_seed_run(C.byref(cp), pts.ctypes.data_as(_PF), pts.shape[0], pts.shape[1], int(through_global_map_callback),
                            C.byref(r))
    assert rc == 0, rc
    n = r.w * r.h
    out = dict(w=r.w, h=r.h, origin_x=r.origin_x, origin_y=r.origin_y, res=np.float32(r.res))
    for k in ("occ_border", "skel", "skel_framed", "occ_raw_nodisc"):
        out[k] = _arr(getattr(r, k), n, np.int8).reshape(r.h, r.w)
    out["seeds"] = _arr(r.seeds, 2 * r.n_seeds, np.float64).reshape(-1, 2)
    out["rows_info"] = _arr(r.rows_info, 4 * r.n_rows_info, np.float64).reshape(-1, 4)
    out["cluster_info"] = _arr(r.cluster_info, 3 * r.n_cluster_info, np.float64).reshape(-1, 3)
    nc = r.n_clusters
    out["n_clusters"] = nc
    for k, dt in (("cl_size", np.int32), ("cl_first", np.int32), ("cl_cx", np.float32), ("cl_cy", np.float32),
                  ("cl_len", np.float32)):
        out[k] = _arr(getattr(r, k), nc, dt)
    out["cl_cell_off"] = _arr(r.cl_cell_off, nc + 1, np.int32)
    out["cl_cells"] = _arr(r.cl_cells, int(out["cl_cell_off"][-1]) if nc else 0, np.int32)
    out["n_after_ror"] = r.n_after_ror
    lib().ref_seed_result_free(C.byref(r))
    return out


def gvd_stage(seeds, skel_framed, origin_x, origin_y, res, rows_info) -> dict:
    sd = np.ascontiguousarray(seeds, np.float64).reshape(-1, 2)
    sk = np.ascontiguousarray(skel_framed, np.int8)
    rw = np.ascontiguousarray(rows_info, np.float64).reshape(-1, 4)
    h, w = sk.shape
    g = _Graph()
    rc = lib().ref_gvd_run(sd.ctypes.data_as(_PD), len(sd), sk.ctypes.data_as(_P8), w, h, float(origin_x), float(origin_y),
                           C.c_float(float(np.float32(res))), rw.ctypes.data_as(_PD), len(rw), C.byref(g))
    assert rc == 0, rc
    out = dict(published=g.published, merged_seeds=_arr(g.merged_seeds, 2 * g.n_merged_seeds, np.float64).reshape(-1, 2),
               n_voro_edges=g.n_voro_edges)
    if g.published:
        xyz = _arr(g.nodes_xyz, 3 * g.n_nodes, np.float64).reshape(-1, 3)
        out.update(resolution=g.resolution, origin_x=g.origin_x, origin_y=g.origin_y, n_nodes=g.n_nodes, n_edges=g.n_edges,
                   nodes_xyz=xyz, nodes=xyz[:, :2].copy(),
                   node_labels=_arr(g.node_labels, g.n_nodes, np.int32),
                   node_cluster_indices=_arr(g.node_cluster_indices, g.n_nodes, np.int32),
                   node_label_counts=_arr(g.node_label_counts, g.n_nodes, np.int32),
                   node_label_clusters=_arr(g.node_label_clusters, g.n_label_entries, np.int32),
                   node_label_types=_arr(g.node_label_types, g.n_label_entries, np.int32),
                   edges=_arr(g.edges, 2 * g.n_edges, np.int32).reshape(-1, 2),
                   edge_lengths=_arr(g.edge_lengths, g.n_edges, np.float32),
                   edge_clearances=_arr(g.edge_clearances, g.n_edges, np.float32))
    lib().ref_graph_free(C.byref(g))
    return out


def inflate(grid: np.ndarray, res: float, inflation_radius: float) -> np.ndarray:
    g = np.ascontiguousarray(grid, np.int8)
    out = np.empty_like(g)
    lib().ref_step_inflate(g.ctypes.data_as(_P8), g.shape[1], g.shape[0], C.c_float(res), C.c_float(inflation_radius),
                           out.ctypes.data_as(_P8))
    return out


def mark_borders(grid: np.ndarray) -> np.ndarray:
    g = np.ascontiguousarray(grid, np.int8)
    out = np.empty_like(g)
    lib().ref_step_mark_borders(g.ctypes.data_as(_P8), g.shape[1], g.shape[0], out.ctypes.data_as(_P8))
    return out


def skeletonize(grid: np.ndarray) -> np.ndarray:
    g = np.ascontiguousarray(grid, np.int8)
    out = np.empty_like(g)
    lib().ref_step_skeletonize(g.ctypes.data_as(_P8), g.shape[1], g.shape[0], out.ctypes.data_as(_P8))
    return out


def point_in_polygon(x: float, y: float, poly: np.ndarray) -> bool:
    p = np.ascontiguousarray(poly, np.float64).reshape(-1, 2)
    return bool(lib().ref_step_point_in_polygon(float(x), float(y), p.ctypes.data_as(_PD), len(p)))


def trim_path(path_xy, grid, origin_x, origin_y, res) -> int:
    p = np.ascontiguousarray(path_xy, np.float64).reshape(-1, 2)
    g = np.ascontiguousarray(grid, np.int8)
    return int(lib().ref_trim_path(p.ctypes.data_as(_PD), len(p), g.ctypes.data_as(_P8), g.shape[1], g.shape[0],
                                   float(origin_x), float(origin_y), C.c_float(float(np.float32(res)))))
