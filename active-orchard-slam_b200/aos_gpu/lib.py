"""ctypes binding of libaos_gpu.so (include/aos_gpu.h).

This is plumbing for the tests and bench.py; the product is the C-ABI library.  There is NO CPU
fallback: if the library is missing or no CUDA device is present, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libaos_gpu.so")

AOS_MEM_HOST, AOS_MEM_DEVICE = 0, 1
GRID_RAW, GRID_INFLATED, GRID_OCCUPANCY, GRID_OPENED, GRID_SKELETON, GRID_SKELETON_FRAMED = range(6)
FMT_INT8, FMT_BITS = 0, 1

EXPORTED_SYMBOLS = [
    "aos_set_profiling", "aos_get_stage_times",
    "aos_create", "aos_destroy", "aos_last_error", "aos_version", "aos_set_stream", "aos_synchronize",
    "aos_bits_pitch_words", "aos_grid_geometry", "aos_seed_stage", "aos_seed_summary_get", "aos_get_grid",
    "aos_grid_device_bits", "aos_get_labels", "aos_get_clusters", "aos_get_tree_rows", "aos_inflate_bits",
    "aos_open_bits", "aos_thin_bits", "aos_pack_int8", "aos_unpack_int8",
    "aos_select_seeds", "aos_get_seeds", "aos_get_rows_info", "aos_get_launch_count",
    "aos_gvd_stage", "aos_gvd_stage_bits", "aos_get_graph", "aos_map_to_graph", "aos_map_to_graph_batch", "aos_merge_seeds", "aos_voronoi_facets", "aos_voronoi_facets_device", "aos_merge_seeds_device", "aos_trim_path", "aos_set_subdiv_outer_factor", "aos_set_subdiv_literal_splices", "aos_set_device_gate",
    "aos_radius_outlier_removal", "aos_edt_bits", "aos_inflate_bits_edt", "aos_set_clearance", "aos_band_halo_rows", "aos_band_raster", "aos_band_thin_launch", "aos_band_ipc_export", "aos_band_ipc_import",
    "aos_band_thin_launch_p2p", "aos_band_grid_device", "aos_seed_stage_tail", "aos_band_ipc_release", "aos_get_device_gate", "aos_set_voronoi_mode",
    "aos_set_host_wait", "aos_set_subdiv_simd",
]


class AosError(RuntimeError):
    pass


class CSeedParams(C.Structure):
    _fields_ = [
        ("clipping_minz", C.c_float), ("clipping_maxz", C.c_float),
        ("clipping_minx", C.c_float), ("clipping_maxx", C.c_float),
        ("clipping_miny", C.c_float), ("clipping_maxy", C.c_float),
        ("grid_resolution", C.c_float), ("inflation_radius", C.c_float),
        ("cluster_min_length", C.c_double),
        ("n_polygon", C.c_int32), ("polygon", C.POINTER(C.c_double)),
        ("n_exclusion", C.c_int32), ("exclusion", C.POINTER(C.c_float)),
    ]


class CGridInfo(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("resolution", C.c_float),
                ("origin_x", C.c_double), ("origin_y", C.c_double)]


class CCluster(C.Structure):
    _fields_ = [("label", C.c_int32), ("size", C.c_int32), ("center_x", C.c_float), ("center_y", C.c_float),
                ("length", C.c_float), ("reserved", C.c_int32), ("sum_x", C.c_int64), ("sum_y", C.c_int64),
                ("max_d2", C.c_int64)]


class CTreeRow(C.Structure):
    _fields_ = [("center_x", C.c_double), ("center_y", C.c_double), ("start_x", C.c_double), ("start_y", C.c_double),
                ("end_x", C.c_double), ("end_y", C.c_double), ("length", C.c_double), ("cluster", C.c_int32),
                ("reserved", C.c_int32)]


class CStageTime(C.Structure):
    _fields_ = [("name", C.c_char * 32), ("ms", C.c_float)]


class CGvdGraph(C.Structure):
    _fields_ = [("resolution", C.c_float), ("origin_x", C.c_double), ("origin_y", C.c_double),
                ("n_nodes", C.c_int32), ("nodes_xyz", C.POINTER(C.c_double)),
                ("node_labels", C.POINTER(C.c_int32)), ("node_cluster_indices", C.POINTER(C.c_int32)),
                ("node_label_counts", C.POINTER(C.c_int32)), ("n_label_entries", C.c_int32),
                ("node_label_clusters", C.POINTER(C.c_int32)), ("node_label_types", C.POINTER(C.c_int32)),
                ("n_edges", C.c_int32), ("edges", C.POINTER(C.c_int32)), ("edge_lengths", C.POINTER(C.c_float)),
                ("edge_clearances", C.POINTER(C.c_float)), ("n_merged_seeds", C.c_int32),
                ("n_voronoi_edges", C.c_int32), ("n_boundary_points", C.c_int32),
                ("corner_points", C.POINTER(C.c_double)), ("n_rows", C.c_int32)]


class CBand(C.Structure):
    _fields_ = [("row0", C.c_int32), ("rows", C.c_int32), ("halo_lo", C.c_int32), ("halo_hi", C.c_int32)]


class CBatchItem(C.Structure):
    _fields_ = [("ctx", C.c_void_p), ("params", C.POINTER(CSeedParams)), ("points", C.c_void_p), ("n_points", C.c_size_t),
                ("point_step", C.c_uint32), ("off_x", C.c_uint32), ("off_y", C.c_uint32), ("off_z", C.c_uint32),
                ("points_mem", C.c_int), ("status", C.c_int)]


class CSeedSummary(C.Structure):
    _fields_ = [("info", CGridInfo), ("n_clusters", C.c_int32), ("n_rows", C.c_int32),
                ("thinning_launches", C.c_int32), ("thinning_subiters", C.c_int32), ("n_points_in", C.c_int64)]


CLUSTER_DTYPE = np.dtype([("label", "<i4"), ("size", "<i4"), ("center_x", "<f4"), ("center_y", "<f4"),
                          ("length", "<f4"), ("reserved", "<i4"), ("sum_x", "<i8"), ("sum_y", "<i8"),
                          ("max_d2", "<i8")])
ROW_DTYPE = np.dtype([("center_x", "<f8"), ("center_y", "<f8"), ("start_x", "<f8"), ("start_y", "<f8"),
                      ("end_x", "<f8"), ("end_y", "<f8"), ("length", "<f8"), ("cluster", "<i4"), ("reserved", "<i4")])

_lib = None


def load() -> C.CDLL:
    """Load libaos_gpu.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AosError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    L = C.CDLL(LIB_PATH)
    vp, i32, sz = C.c_void_p, C.c_int32, C.c_size_t
    L.aos_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.aos_destroy.argtypes = [vp]
    L.aos_destroy.restype = None
    L.aos_last_error.argtypes = [vp]
    L.aos_last_error.restype = C.c_char_p
    L.aos_version.restype = C.c_char_p
    L.aos_set_stream.argtypes = [vp, vp]
    L.aos_synchronize.argtypes = [vp]
    L.aos_bits_pitch_words.argtypes = [i32]
    L.aos_set_profiling.argtypes = [vp, C.c_int]
    L.aos_get_stage_times.argtypes = [vp, vp, i32, C.POINTER(i32)]
    L.aos_grid_geometry.argtypes = [C.POINTER(CSeedParams), C.POINTER(CGridInfo)]
    L.aos_seed_stage.argtypes = [vp, C.POINTER(CSeedParams), vp, sz, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]
    L.aos_seed_summary_get.argtypes = [vp, C.POINTER(CSeedSummary)]
    L.aos_get_grid.argtypes = [vp, C.c_int, C.c_int, vp, sz, C.c_int]
    L.aos_grid_device_bits.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(i32)]
    L.aos_get_labels.argtypes = [vp, vp, sz, C.c_int]
    L.aos_get_clusters.argtypes = [vp, vp, i32, C.POINTER(i32)]
    L.aos_get_tree_rows.argtypes = [vp, vp, i32, C.POINTER(i32)]
    L.aos_inflate_bits.argtypes = [vp, vp, vp, vp, i32, i32, i32]
    L.aos_open_bits.argtypes = [vp, vp, vp, i32, i32]
    L.aos_thin_bits.argtypes = [vp, vp, i32, i32, C.POINTER(i32), C.POINTER(i32)]
    L.aos_pack_int8.argtypes = [vp, vp, C.c_int, vp, i32, i32]
    L.aos_unpack_int8.argtypes = [vp, vp, vp, C.c_int, i32, i32]
    L.aos_select_seeds.argtypes = [vp, C.POINTER(i32), C.POINTER(i32)]
    L.aos_get_seeds.argtypes = [vp, vp, i32, C.POINTER(i32)]
    L.aos_get_rows_info.argtypes = [vp, vp, i32, C.POINTER(i32)]
    L.aos_get_launch_count.argtypes = [vp, C.POINTER(C.c_int64)]
    L.aos_gvd_stage.argtypes = [vp, vp, i32, vp, i32, vp, C.POINTER(CGridInfo)]
    L.aos_gvd_stage_bits.argtypes = [vp, vp, i32, vp, i32, vp, C.c_int, C.POINTER(CGridInfo)]
    L.aos_get_graph.argtypes = [vp, C.POINTER(CGvdGraph)]
    L.aos_map_to_graph.argtypes = [vp, C.POINTER(CSeedParams), vp, sz, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]
    L.aos_set_subdiv_outer_factor.argtypes = [C.c_float]
    L.aos_radius_outlier_removal.argtypes = [vp, vp, sz, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_float,
                                             i32, C.POINTER(vp), C.POINTER(sz)]
    L.aos_set_clearance.argtypes = [vp, C.c_int]
    L.aos_edt_bits.argtypes = [vp, vp, i32, i32, vp, vp]
    L.aos_inflate_bits_edt.argtypes = [vp, vp, vp, i32, i32, i32]
    L.aos_band_halo_rows.argtypes = [C.POINTER(CSeedParams)]
    L.aos_band_raster.argtypes = [vp, C.POINTER(CSeedParams), C.POINTER(CBand), vp, sz, C.c_uint32, C.c_uint32, C.c_uint32,
                                  C.c_uint32, C.c_int]
    L.aos_band_thin_launch.argtypes = [vp, C.POINTER(i32)]
    L.aos_band_ipc_export.argtypes = [vp, i32, vp]
    L.aos_band_ipc_import.argtypes = [vp, i32, i32, vp, i32]
    L.aos_band_thin_launch_p2p.argtypes = [vp, C.POINTER(i32)]
    L.aos_band_ipc_release.argtypes = [vp]
    L.aos_set_voronoi_mode.argtypes = [vp, i32]
    L.aos_get_device_gate.restype = i32
    L.aos_set_host_wait.argtypes = [C.c_int, i32]
    L.aos_set_subdiv_simd.argtypes = [i32]
    L.aos_band_grid_device.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(i32), C.POINTER(i32)]
    L.aos_seed_stage_tail.argtypes = [vp, C.POINTER(CSeedParams), vp, vp]
    L.aos_map_to_graph_batch.argtypes = [C.POINTER(CBatchItem), i32, i32]
    L.aos_merge_seeds.argtypes = [vp, i32, vp, C.POINTER(i32)]
    L.aos_trim_path.argtypes = [vp, vp, i32, C.c_double, vp, C.c_int, vp, C.POINTER(i32)]
    L.aos_merge_seeds_device.argtypes = [vp, vp, i32, vp, C.POINTER(i32)]
    L.aos_voronoi_facets_device.argtypes = [vp, vp, i32, C.c_double, C.c_double, C.c_double, C.c_double, vp, vp, i32,
                                            C.POINTER(i32)]
    L.aos_voronoi_facets.argtypes = [vp, i32, C.c_double, C.c_double, C.c_double, C.c_double, vp, i32, vp, i32,
                                     C.POINTER(i32), C.POINTER(i32)]
    _lib = L
    return L


@dataclass
class SeedParams:
    """aos_seed_gen_node parameters (config/aos_planner_params.yaml names and defaults)."""
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

    def to_c(self) -> CSeedParams:
        poly = np.ascontiguousarray(self.polygon, dtype=np.float64).reshape(-1, 2)
        excl = np.ascontiguousarray(self.exclusion, dtype=np.float32).reshape(-1, 3)
        p = CSeedParams(self.clipping_minz, self.clipping_maxz, self.clipping_minx, self.clipping_maxx,
                        self.clipping_miny, self.clipping_maxy, self.grid_resolution, self.inflation_radius,
                        self.cluster_min_length, len(poly), poly.ctypes.data_as(C.POINTER(C.c_double)), len(excl),
                        excl.ctypes.data_as(C.POINTER(C.c_float)))
        p._keep = (poly, excl)
        return p


def grid_geometry(params: SeedParams) -> CGridInfo:
    gi = CGridInfo()
    rc = load().aos_grid_geometry(C.byref(params.to_c()), C.byref(gi))
    if rc != 0:
        raise AosError(f"aos_grid_geometry -> {rc}")
    return gi


class DevicePoints:
    """A library-owned device cloud of 16-byte x,y,z,pad records (quacks like a [N,4] float32 CUDA tensor)."""

    def __init__(self, ptr: int, n: int):
        self.ptr, self.shape = ptr, (n, 4)

    def data_ptr(self):
        return self.ptr

    def stride(self, dim):
        return 4 if dim == 0 else 1

    def element_size(self):
        return 4


CLUSTER_DIGEST_FIELDS = ("label", "size", "center_x", "center_y", "length", "sum_x", "sum_y", "max_d2")
ROW_DIGEST_FIELDS = ("center_x", "center_y", "start_x", "start_y", "end_x", "end_y", "length", "cluster")
DIGEST_DTYPES = {"occupancy": np.uint32, "skeleton": np.uint32, "skeleton_framed": np.uint32, "clusters.label": np.int32,
                 "clusters.size": np.int32, "clusters.center_x": np.float32, "clusters.center_y": np.float32,
                 "clusters.length": np.float32, "clusters.sum_x": np.int64, "clusters.sum_y": np.int64, "clusters.max_d2": np.int64,
                 "rows.cluster": np.int32, "graph.nodes_xyz": np.float64, "graph.edge_lengths": np.float32,
                 "graph.edge_clearances": np.float32}


def digest_of(artefacts: dict, parts: bool = False):
    """sha256 per named array (fixed dtype per name, C order) and one digest over all of them."""
    import hashlib
    per = {}
    for k in sorted(artefacts):
        v = artefacts[k]
        if v is None:
            per[k] = hashlib.sha256(b"").hexdigest()
            continue
        dt = DIGEST_DTYPES.get(k, np.float64 if k.startswith("rows.") or k in ("seeds", "rows_info") else np.int32)
        per[k] = hashlib.sha256(np.ascontiguousarray(v, dtype=dt).tobytes()).hexdigest()
    total = hashlib.sha256("".join(f"{k}={per[k]};" for k in sorted(per)).encode()).hexdigest()
    return (total, per) if parts else total


GRAPH_DIGEST_KEYS = ("nodes_xyz", "node_labels", "node_cluster_indices", "node_label_counts", "node_label_clusters",
                     "node_label_types", "edges", "edge_lengths", "edge_clearances")


class Context:
    """One libaos_gpu context (device buffers are reused call after call)."""

    def __init__(self, device: int = 0):
        self.L = load()
        h = C.c_void_p()
        rc = self.L.aos_create(device, C.byref(h))
        if rc != 0:
            raise AosError(f"aos_create(device={device}) -> {rc} (no CUDA device? this library has no CPU path)")
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.L.aos_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise AosError(f"{what} -> {rc}: {self.L.aos_last_error(self.h).decode()}")

    def set_stream(self, cuda_stream_ptr: int):
        self._check(self.L.aos_set_stream(self.h, C.c_void_p(cuda_stream_ptr)), "aos_set_stream")

    def set_profiling(self, on: bool):
        self._check(self.L.aos_set_profiling(self.h, int(on)), "aos_set_profiling")

    def stage_times(self):
        n = C.c_int32()
        self._check(self.L.aos_get_stage_times(self.h, None, 0, C.byref(n)), "aos_get_stage_times")
        arr = (CStageTime * max(n.value, 1))()
        self._check(self.L.aos_get_stage_times(self.h, arr, n.value, C.byref(n)), "aos_get_stage_times")
        return [(arr[i].name.decode(), float(arr[i].ms)) for i in range(n.value)]

    def synchronize(self):
        self._check(self.L.aos_synchronize(self.h), "aos_synchronize")

    # ---- seed stage -------------------------------------------------------------------------
    def seed_stage(self, params: SeedParams, points, n_points=None, point_step=None, offsets=(0, 4, 8)):
        """points: numpy float32 [N, k] (host) or an object with data_ptr() (torch CUDA tensor, [N, k] float32)."""
        cp = params.to_c()
        if hasattr(points, "data_ptr"):
            ptr, mem = points.data_ptr(), AOS_MEM_DEVICE
            n = points.shape[0] if n_points is None else n_points
            step = points.stride(0) * points.element_size() if point_step is None else point_step
            self._keep_points = points
        else:
            pts = np.ascontiguousarray(points)
            ptr, mem = pts.ctypes.data, AOS_MEM_HOST
            n = pts.shape[0] if n_points is None else n_points
            step = pts.strides[0] if point_step is None else point_step
            self._keep_points = pts
        rc = self.L.aos_seed_stage(self.h, C.byref(cp), C.c_void_p(ptr), n, step, offsets[0], offsets[1], offsets[2], mem)
        self._check(rc, "aos_seed_stage")
        return self.seed_summary()

    def seed_summary(self) -> CSeedSummary:
        s = CSeedSummary()
        self._check(self.L.aos_seed_summary_get(self.h, C.byref(s)), "aos_seed_summary_get")
        return s

    def grid_int8(self, which: int) -> np.ndarray:
        s = self.seed_summary()
        out = np.empty((s.info.height, s.info.width), np.int8)
        self._check(self.L.aos_get_grid(self.h, which, FMT_INT8, out.ctypes.data_as(C.c_void_p), out.nbytes, AOS_MEM_HOST),
                    "aos_get_grid")
        return out

    def grid_bits(self, which: int) -> np.ndarray:
        s = self.seed_summary()
        pitch = self.L.aos_bits_pitch_words(s.info.width)
        out = np.empty((s.info.height, pitch), np.uint32)
        self._check(self.L.aos_get_grid(self.h, which, FMT_BITS, out.ctypes.data_as(C.c_void_p), out.nbytes, AOS_MEM_HOST),
                    "aos_get_grid")
        return out

    def grid_device_bits(self, which: int):
        p = C.c_void_p()
        pitch = C.c_int32()
        self._check(self.L.aos_grid_device_bits(self.h, which, C.byref(p), C.byref(pitch)), "aos_grid_device_bits")
        return p.value, pitch.value

    def labels(self) -> np.ndarray:
        s = self.seed_summary()
        out = np.empty((s.info.height, s.info.width), np.int32)
        self._check(self.L.aos_get_labels(self.h, out.ctypes.data_as(C.c_void_p), out.size, AOS_MEM_HOST), "aos_get_labels")
        return out

    def clusters(self) -> np.ndarray:
        n = C.c_int32()
        self._check(self.L.aos_get_clusters(self.h, None, 0, C.byref(n)), "aos_get_clusters")
        out = np.zeros(n.value, CLUSTER_DTYPE)
        if n.value:
            self._check(self.L.aos_get_clusters(self.h, out.ctypes.data_as(C.c_void_p), n.value, C.byref(n)), "aos_get_clusters")
        return out

    # ---- the whole path -----------------------------------------------------------------------
    def _points_args(self, points, n_points, point_step):
        if hasattr(points, "data_ptr"):
            self._keep_points = points
            return (points.data_ptr(), AOS_MEM_DEVICE, points.shape[0] if n_points is None else n_points,
                    points.stride(0) * points.element_size() if point_step is None else point_step)
        pts = np.ascontiguousarray(points)
        self._keep_points = pts
        return (pts.ctypes.data, AOS_MEM_HOST, pts.shape[0] if n_points is None else n_points,
                pts.strides[0] if point_step is None else point_step)

    def map_to_graph(self, params: SeedParams, points, fetch=None, n_points=None, point_step=None,
                     offsets=(0, 4, 8)) -> dict:
        """aos_map_to_graph: seed stage (GPU), seed selection (host), gvd stage (host Voronoi + GPU graph).
        With `fetch` (True or a dict of preallocated uint32 arrays "occ_bits"/"skel_bits") the published grids
        come back bit-packed together with clusters, rows, seeds and the graph arrays.  Returns a summary."""
        cp = params.to_c()
        ptr, mem, n, step = self._points_args(points, n_points, point_step)
        rc = self.L.aos_map_to_graph(self.h, C.byref(cp), C.c_void_p(ptr), n, step, offsets[0], offsets[1], offsets[2], mem)
        s = self.seed_summary() if rc in (0, -4) else None
        if rc == -4 and s is not None:   # AOS_ERR_STATE: no rows/seeds on this map, the reference publishes no graph
            graph = None
        else:
            self._check(rc, "aos_map_to_graph")
            graph = self.graph(copy=fetch is not None)
        info = {"pipeline": "aos_map_to_graph (seed_stage + select_seeds + gvd_stage)", "width": s.info.width,
                "height": s.info.height, "n_clusters": s.n_clusters, "n_rows": s.n_rows,
                "graph": None if graph is None else {"nodes": int(graph["n_nodes"]), "edges": int(graph["n_edges"]),
                                                     "merged_seeds": int(graph["n_merged_seeds"])}}
        if fetch is not None:
            if fetch is True:
                fetch = {}
            d2h = 0
            pitch = self.L.aos_bits_pitch_words(s.info.width)
            for key, gid in (("occ_bits", GRID_OCCUPANCY), ("skel_bits", GRID_SKELETON_FRAMED)):
                buf = fetch.get(key)
                if buf is None or buf.shape != (s.info.height, pitch):
                    buf = fetch[key] = np.empty((s.info.height, pitch), np.uint32)
                self._check(self.L.aos_get_grid(self.h, gid, FMT_BITS, buf.ctypes.data_as(C.c_void_p), buf.nbytes,
                                                AOS_MEM_HOST), "aos_get_grid")
                d2h += buf.nbytes
            cl, rows = self.clusters(), self.tree_rows()
            d2h += cl.nbytes + rows.nbytes
            if graph is not None:
                d2h += sum(v.nbytes for v in graph.values() if isinstance(v, np.ndarray))
            info["d2h_bytes"] = d2h
            info["fetched"] = fetch
            info["graph_arrays"] = graph
        return info

    # ---- row-band sharding (see bands.py) ------------------------------------------------------
    def band_halo_rows(self, params: SeedParams) -> int:
        return int(self.L.aos_band_halo_rows(C.byref(params.to_c())))

    def band_raster(self, params: SeedParams, band, points, n_points=None, point_step=None, offsets=(0, 4, 8)):
        cp = params.to_c()
        cb = CBand(band.row0, band.rows, band.halo_lo, band.halo_hi)
        ptr, mem, n, step = self._points_args(points, n_points, point_step)
        self._check(self.L.aos_band_raster(self.h, C.byref(cp), C.byref(cb), C.c_void_p(ptr), n, step, offsets[0],
                                           offsets[1], offsets[2], mem), "aos_band_raster")

    def band_thin_launch(self) -> bool:
        d = C.c_int32()
        self._check(self.L.aos_band_thin_launch(self.h, C.byref(d)), "aos_band_thin_launch")
        return bool(d.value)

    def band_ipc_export(self, buffer: int) -> bytes:
        h = (C.c_ubyte * 64)()
        self._check(self.L.aos_band_ipc_export(self.h, buffer, h), "aos_band_ipc_export")
        return bytes(h)

    def band_ipc_import(self, side: int, buffer: int, handle: bytes, peer_first_global_row: int):
        cache = self.__dict__.setdefault("_ipc_imported", {})
        if cache.get((side, buffer)) == (handle, peer_first_global_row):
            return          # same allocation as last map: the mapping is still open
        h = (C.c_ubyte * 64).from_buffer_copy(handle)
        cache.pop((side, buffer), None)
        self._check(self.L.aos_band_ipc_import(self.h, side, buffer, h, peer_first_global_row), "aos_band_ipc_import")
        cache[(side, buffer)] = (handle, peer_first_global_row)

    def set_voronoi_mode(self, device: bool):
        """False: Subdiv2D insertion replay (bit-exact, default).  True: parallel Voronoi cells on the device (opt-in)."""
        self._check(self.L.aos_set_voronoi_mode(self.h, 1 if device else 0), "aos_set_voronoi_mode")

    def band_ipc_release(self):
        self._check(self.L.aos_band_ipc_release(self.h), "aos_band_ipc_release")

    def band_thin_launch_p2p(self) -> bool:
        d = C.c_int32()
        self._check(self.L.aos_band_thin_launch_p2p(self.h, C.byref(d)), "aos_band_thin_launch_p2p")
        return bool(d.value)

    def trim_path(self, path_xy, safety_distance=0.2, skeleton_bits=None, info=None) -> int:
        """aos_trim_path (trimPathNearOccupiedRegions): poses that remain.  skeleton_bits: uint32 [H, pitch] host grid
        with info=(resolution, origin_x, origin_y, width), or None for this context's framed skeleton."""
        p = np.ascontiguousarray(path_xy, np.float64).reshape(-1, 2)
        n = C.c_int32()
        if skeleton_bits is not None:
            sk = np.ascontiguousarray(skeleton_bits, np.uint32)
            gi = CGridInfo(int(info[3]), sk.shape[0], float(info[0]), float(info[1]), float(info[2]))
            rc = self.L.aos_trim_path(self.h, p.ctypes.data_as(C.c_void_p), len(p), float(safety_distance),
                                      sk.ctypes.data_as(C.c_void_p), AOS_MEM_HOST, C.byref(gi), C.byref(n))
        else:
            rc = self.L.aos_trim_path(self.h, p.ctypes.data_as(C.c_void_p), len(p), float(safety_distance), None, AOS_MEM_HOST,
                                      None, C.byref(n))
        self._check(rc, "aos_trim_path")
        return n.value

    def merge_seeds_device(self, seeds):
        sd = np.ascontiguousarray(seeds, np.float64).reshape(-1, 2)
        out = np.zeros((max(len(sd), 1), 2), np.float64)
        n = C.c_int32()
        self._check(self.L.aos_merge_seeds_device(self.h, sd.ctypes.data_as(C.c_void_p), len(sd), out.ctypes.data_as(C.c_void_p),
                                                  C.byref(n)), "aos_merge_seeds_device")
        return out[:n.value].copy()

    def voronoi_facets_device(self, seeds, minx, maxx, miny, maxy):
        """aos_voronoi_facets_device: (slot xy float32 [K,2], next slot int32 [K])."""
        sd = np.ascontiguousarray(seeds, np.float64).reshape(-1, 2)
        n = C.c_int32()
        args = (self.h, sd.ctypes.data_as(C.c_void_p), len(sd), minx, maxx, miny, maxy)
        self._check(self.L.aos_voronoi_facets_device(*args, None, None, 0, C.byref(n)), "aos_voronoi_facets_device")
        xy = np.zeros((n.value, 2), np.float32)
        nxt = np.zeros(n.value, np.int32)
        self._check(self.L.aos_voronoi_facets_device(*args, xy.ctypes.data_as(C.c_void_p), nxt.ctypes.data_as(C.c_void_p),
                                                     n.value, C.byref(n)), "aos_voronoi_facets_device")
        return xy, nxt

    def band_grid_device(self, which: int):
        p, pitch, rows = C.c_void_p(), C.c_int32(), C.c_int32()
        self._check(self.L.aos_band_grid_device(self.h, which, C.byref(p), C.byref(pitch), C.byref(rows)), "aos_band_grid_device")
        return p.value, pitch.value, rows.value

    def seed_stage_tail(self, params: SeedParams, skeleton_bits, occupancy_bits=None):
        """skeleton_bits / occupancy_bits: full-size device bit grids (objects with data_ptr(), e.g. torch tensors)."""
        cp = params.to_c()
        self._keep_grids = (skeleton_bits, occupancy_bits)
        rc = self.L.aos_seed_stage_tail(self.h, C.byref(cp), C.c_void_p(skeleton_bits.data_ptr()),
                                        C.c_void_p(occupancy_bits.data_ptr()) if occupancy_bits is not None else None)
        self._check(rc, "aos_seed_stage_tail")
        return self.seed_summary()

    # ---- gvd stage ---------------------------------------------------------------------------
    def gvd_stage(self, seeds, rows_info, skeleton=None, info=None) -> dict:
        """seeds: /voronoi_seeds positions [S,2] (un-merged); rows_info [R,4]; skeleton: int8 [H,W] of
        /skeletonized_occupancy_grid with info=(resolution, origin_x, origin_y), or None to use this context's."""
        sd = np.ascontiguousarray(seeds, np.float64).reshape(-1, 2)
        rw = np.ascontiguousarray(rows_info, np.float64).reshape(-1, 4)
        if skeleton is not None and np.asarray(skeleton).dtype == np.uint32:   # AOS_FMT_BITS grid [H, pitch], host
            sk = np.ascontiguousarray(skeleton, np.uint32)
            gi = CGridInfo(int(info[3]), sk.shape[0], float(info[0]), float(info[1]), float(info[2]))
            rc = self.L.aos_gvd_stage_bits(self.h, sd.ctypes.data_as(C.c_void_p), len(sd), rw.ctypes.data_as(C.c_void_p), len(rw),
                                           sk.ctypes.data_as(C.c_void_p), AOS_MEM_HOST, C.byref(gi))
        elif skeleton is not None:
            sk = np.ascontiguousarray(skeleton, np.int8)
            gi = CGridInfo(sk.shape[1], sk.shape[0], float(info[0]), float(info[1]), float(info[2]))
            rc = self.L.aos_gvd_stage(self.h, sd.ctypes.data_as(C.c_void_p), len(sd), rw.ctypes.data_as(C.c_void_p), len(rw),
                                      sk.ctypes.data_as(C.c_void_p), C.byref(gi))
        else:
            rc = self.L.aos_gvd_stage(self.h, sd.ctypes.data_as(C.c_void_p), len(sd), rw.ctypes.data_as(C.c_void_p), len(rw),
                                      None, None)
        self._check(rc, "aos_gvd_stage")
        return self.graph()

    def graph(self, copy=True) -> dict:
        g = CGvdGraph()
        self._check(self.L.aos_get_graph(self.h, C.byref(g)), "aos_get_graph")

        def arr(ptr, n, dt):
            if n == 0:
                return np.zeros(0, dt)
            a = np.ctypeslib.as_array(ptr, shape=(n,))
            return a.astype(dt, copy=True) if copy else a

        xyz = arr(g.nodes_xyz, 3 * g.n_nodes, np.float64).reshape(-1, 3)
        return dict(resolution=float(g.resolution), origin_x=g.origin_x, origin_y=g.origin_y, n_nodes=g.n_nodes,
                    n_edges=g.n_edges, nodes_xyz=xyz, nodes=xyz[:, :2],
                    node_labels=arr(g.node_labels, g.n_nodes, np.int32),
                    node_cluster_indices=arr(g.node_cluster_indices, g.n_nodes, np.int32),
                    node_label_counts=arr(g.node_label_counts, g.n_nodes, np.int32),
                    node_label_clusters=arr(g.node_label_clusters, g.n_label_entries, np.int32),
                    node_label_types=arr(g.node_label_types, g.n_label_entries, np.int32),
                    edges=arr(g.edges, 2 * g.n_edges, np.int32).reshape(-1, 2),
                    edge_lengths=arr(g.edge_lengths, g.n_edges, np.float32),
                    edge_clearances=arr(g.edge_clearances, g.n_edges, np.float32),
                    corner_points=arr(g.corner_points, 8 * g.n_rows, np.float64).reshape(-1, 4, 2),
                    n_merged_seeds=g.n_merged_seeds, n_voronoi_edges=g.n_voronoi_edges,
                    n_boundary_points=g.n_boundary_points)

    def radius_outlier_removal(self, points, radius=0.2, min_neighbors=2, n_points=None, point_step=None, offsets=(0, 4, 8)):
        """aos_radius_outlier_removal.  Returns (device pointer of the kept x,y,z,1 records, count); pass them on
        as seed_stage(params, DevicePoints(ptr, count))."""
        ptr, mem, n, step = self._points_args(points, n_points, point_step)
        out, cnt = C.c_void_p(), C.c_size_t()
        self._check(self.L.aos_radius_outlier_removal(self.h, C.c_void_p(ptr), n, step, offsets[0], offsets[1], offsets[2], mem,
                                                      radius, min_neighbors, C.byref(out), C.byref(cnt)), "aos_radius_outlier_removal")
        return DevicePoints(out.value, cnt.value)

    def set_clearance(self, on: bool):
        self._check(self.L.aos_set_clearance(self.h, int(on)), "aos_set_clearance")

    def launch_count(self) -> int:
        n = C.c_int64()
        self._check(self.L.aos_get_launch_count(self.h, C.byref(n)), "aos_get_launch_count")
        return n.value

    # ---- host seed selection ----------------------------------------------------------------
    def select_seeds(self):
        """Returns (seeds [S,2] float64 in /voronoi_seeds publish order, (n_virtual, n_ray, n_endpoint),
        rows_info [R,4] float64 sorted by centre)."""
        n = C.c_int32()
        counts = (C.c_int32 * 3)()
        self._check(self.L.aos_select_seeds(self.h, C.byref(n), counts), "aos_select_seeds")
        seeds = np.zeros((n.value, 2), np.float64)
        if n.value:
            self._check(self.L.aos_get_seeds(self.h, seeds.ctypes.data_as(C.c_void_p), n.value, C.byref(n)), "aos_get_seeds")
        m = C.c_int32()
        self._check(self.L.aos_get_rows_info(self.h, None, 0, C.byref(m)), "aos_get_rows_info")
        rows = np.zeros((m.value, 4), np.float64)
        if m.value:
            self._check(self.L.aos_get_rows_info(self.h, rows.ctypes.data_as(C.c_void_p), m.value, C.byref(m)), "aos_get_rows_info")
        return seeds, tuple(counts), rows

    def result_artefacts(self) -> dict:
        """Everything the path publishes after aos_map_to_graph / the band tail, as plain arrays keyed like
        digest_of() expects: the three published bit grids, the cluster table, rows, seeds and every GvdGraph array."""
        art = {}
        for name, gid in (("occupancy", GRID_OCCUPANCY), ("skeleton", GRID_SKELETON), ("skeleton_framed", GRID_SKELETON_FRAMED)):
            art[name] = self.grid_bits(gid)
        cl, rows = self.clusters(), self.tree_rows()
        for f in CLUSTER_DIGEST_FIELDS:
            art["clusters." + f] = np.ascontiguousarray(cl[f])
        for f in ROW_DIGEST_FIELDS:
            art["rows." + f] = np.ascontiguousarray(rows[f])
        seeds, _counts, rows_info = self.select_seeds_cached()
        art["seeds"], art["rows_info"] = seeds, rows_info
        try:
            g = self.graph()
        except AosError:
            g = None
        for k in GRAPH_DIGEST_KEYS:
            art["graph." + k] = None if g is None else g[k]
        return art

    def result_digest(self, parts: bool = False):
        """sha256 over result_artefacts().  Used by bench.py and the tests to compare runs (maps in flight, GPUs, band
        splits, the CPU oracle) without moving the arrays around.  parts=True also returns the per-array digests."""
        return digest_of(self.result_artefacts(), parts)

    def select_seeds_cached(self):
        """Seeds / counts / rows_info of the last aos_select_seeds or aos_map_to_graph on this context (no recompute)."""
        n = C.c_int32()
        self._check(self.L.aos_get_seeds(self.h, None, 0, C.byref(n)), "aos_get_seeds")
        seeds = np.zeros((n.value, 2), np.float64)
        if n.value:
            self._check(self.L.aos_get_seeds(self.h, seeds.ctypes.data_as(C.c_void_p), n.value, C.byref(n)), "aos_get_seeds")
        counts = (0, 0, 0)   # the split into virtual / ray / endpoint seeds is only returned by aos_select_seeds
        m = C.c_int32()
        self._check(self.L.aos_get_rows_info(self.h, None, 0, C.byref(m)), "aos_get_rows_info")
        rows = np.zeros((m.value, 4), np.float64)
        if m.value:
            self._check(self.L.aos_get_rows_info(self.h, rows.ctypes.data_as(C.c_void_p), m.value, C.byref(m)), "aos_get_rows_info")
        return seeds, tuple(counts), rows

    def tree_rows(self) -> np.ndarray:
        n = C.c_int32()
        self._check(self.L.aos_get_tree_rows(self.h, None, 0, C.byref(n)), "aos_get_tree_rows")
        out = np.zeros(n.value, ROW_DTYPE)
        if n.value:
            self._check(self.L.aos_get_tree_rows(self.h, out.ctypes.data_as(C.c_void_p), n.value, C.byref(n)), "aos_get_tree_rows")
        return out


def map_to_graph_batch(contexts, params_list, points_list, max_threads: int = 0):
    """aos_map_to_graph_batch: map i on contexts[i] (one host thread each).  Returns the per-map status codes;
    results are read from each context with the usual getters."""
    n = len(contexts)
    items = (CBatchItem * n)()
    keep = []
    for i, (ctx, prm, pts) in enumerate(zip(contexts, params_list, points_list)):
        cp = prm.to_c()
        ptr, mem, cnt, step = ctx._points_args(pts, None, None)
        keep.append((cp, pts))
        items[i] = CBatchItem(ctx.h, C.pointer(cp), ptr, cnt, step, 0, 4, 8, mem, 0)
    rc = load().aos_map_to_graph_batch(items, n, max_threads)
    if rc != 0:
        raise AosError(f"aos_map_to_graph_batch -> {rc}")
    return [items[i].status for i in range(n)]


def merge_seeds(seeds) -> np.ndarray:
    """aos_merge_seeds: voronoiSeedsCallback's greedy 0.5 m merge (host, no device needed)."""
    sd = np.ascontiguousarray(seeds, np.float64).reshape(-1, 2)
    out = np.zeros_like(sd)
    n = C.c_int32()
    rc = load().aos_merge_seeds(sd.ctypes.data_as(C.c_void_p), len(sd), out.ctypes.data_as(C.c_void_p), C.byref(n))
    if rc != 0:
        raise AosError(f"aos_merge_seeds -> {rc}")
    return out[:n.value].copy()


def voronoi_facets(seeds, minx, maxx, miny, maxy):
    """aos_voronoi_facets: VoronoiDiagram::compute on the host.  Returns (xy float32 [K,2], off int32 [F+1])."""
    sd = np.ascontiguousarray(seeds, np.float64).reshape(-1, 2)
    L = load()
    nf, npts = C.c_int32(), C.c_int32()
    args = (sd.ctypes.data_as(C.c_void_p), len(sd), minx, maxx, miny, maxy)
    rc = L.aos_voronoi_facets(*args, None, 0, None, 0, C.byref(nf), C.byref(npts))
    if rc != 0:
        raise AosError(f"aos_voronoi_facets -> {rc}")
    xy = np.zeros((npts.value, 2), np.float32)
    off = np.zeros(nf.value + 1, np.int32)
    rc = L.aos_voronoi_facets(*args, xy.ctypes.data_as(C.c_void_p), npts.value, off.ctypes.data_as(C.c_void_p),
                              len(off), C.byref(nf), C.byref(npts))
    if rc != 0:
        raise AosError(f"aos_voronoi_facets -> {rc}")
    return xy, off


def unpack_bits(bits: np.ndarray, width: int) -> np.ndarray:
    """Host-side view of an AOS_FMT_BITS grid as bool [H, W] (test helper)."""
    b = np.unpackbits(bits.view(np.uint8), axis=1, bitorder="little")
    return b[:, :width].astype(bool)


def pack_bits(img: np.ndarray) -> np.ndarray:
    """bool/0-1 [H, W] -> uint32 [H, pitch] in the library's bit layout (test helper)."""
    h, w = img.shape
    pitch = (((w + 31) >> 5) + 3) & ~3
    pad = np.zeros((h, pitch * 32), np.uint8)
    pad[:, :w] = img != 0
    return np.packbits(pad, axis=1, bitorder="little").view(np.uint32).reshape(h, pitch)
