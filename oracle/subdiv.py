"""cv::Subdiv2D stand-in: the REAL OpenCV implementation through cv2 (TEST INFRASTRUCTURE ONLY).

Restates VoronoiDiagram::compute (src/utils/voronoi_diagram.cpp:16-94): bounding rect +-1 m, seeds
cast to float32 and clipped with a 0.1 m margin, inserted in order with per-seed try/catch, then
getVoronoiFacetList.  The reference pins OpenCV only through `libopencv-dev` (package.xml:48; ROS 2
Humble => 4.5.4); this container has cv2 4.13.0, so last-ulp circumcentres may differ from 4.5.4.
The rect is passed as the rounded int Rect that 4.5.4's Subdiv2D(Rect) constructor would receive.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import oracle as _o


def subdiv_inputs(seeds: np.ndarray, minx: float, maxx: float, miny: float, maxy: float):
    s = np.ascontiguousarray(seeds, dtype=np.float64).reshape(-1, 2)
    n = len(s)
    rect_i = (C.c_int * 4)()
    rect_f = (C.c_float * 4)()
    pts = np.zeros((n, 2), np.float32)
    keep = np.zeros(n, np.uint8)
    valid = C.c_int(0)
    L = _o.lib()
    L.orc_gvd_subdiv_inputs.argtypes = [C.POINTER(C.c_double), C.c_int, C.c_double, C.c_double, C.c_double,
                                        C.c_double, C.POINTER(C.c_int), C.POINTER(C.c_float),
                                        C.POINTER(C.c_float), C.POINTER(C.c_uint8), C.POINTER(C.c_int)]
    L.orc_gvd_subdiv_inputs(s.ctypes.data_as(C.POINTER(C.c_double)), n, minx, maxx, miny, maxy, rect_i, rect_f,
                            pts.ctypes.data_as(C.POINTER(C.c_float)), keep.ctypes.data_as(C.POINTER(C.c_uint8)),
                            C.byref(valid))
    return bool(valid.value), tuple(rect_i), tuple(rect_f), pts, keep.astype(bool)


def voronoi_facets(seeds: np.ndarray, minx: float, maxx: float, miny: float, maxy: float):
    """Returns (facets_xy float32 [K,2], facet_off int32 [F+1], inserted float32 points)."""
    import cv2

    valid, rect_i, _rect_f, pts, keep = subdiv_inputs(seeds, minx, maxx, miny, maxy)
    if not valid:
        return np.zeros((0, 2), np.float32), np.zeros(1, np.int32), np.zeros((0, 2), np.float32)
    sd = cv2.Subdiv2D(tuple(int(v) for v in rect_i))
    inserted = []
    for (x, y), k in zip(pts, keep):
        if not k:
            continue
        try:
            sd.insert((float(x), float(y)))
            inserted.append((x, y))
        except cv2.error:
            continue  # vd:83-88: individual insertion failures are ignored
    facets, _centers = sd.getVoronoiFacetList([])
    off = np.zeros(len(facets) + 1, np.int32)
    for i, f in enumerate(facets):
        off[i + 1] = off[i] + len(f)
    xy = np.concatenate([np.asarray(f, np.float32).reshape(-1, 2) for f in facets]) if facets else np.zeros((0, 2), np.float32)
    return np.ascontiguousarray(xy, np.float32), off, np.asarray(inserted, np.float32).reshape(-1, 2)


def outer_factor() -> float:
    """big_coord / max(side) of THIS image's cv::Subdiv2D::initDelaunay, probed from the real cv2 (3 up to OpenCV 4.5.x,
    6 in 4.13).  Tests, smoke() and bench.py pass it to aos_set_subdiv_outer_factor so that the library replays the
    Subdiv2D the oracle runs; the library's own default is the reference platform's 3."""
    import cv2
    sd = cv2.Subdiv2D((0, 0, 100, 100))
    return float(sd.getVertex(1)[0][0]) / 100.0
