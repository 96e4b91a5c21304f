"""Pins the CPU oracle (test infrastructure) where something independent exists to pin it against:
the real OpenCV of this image (cv2 4.13: morphologyEx, dilate), scipy (exact EDT, connected components),
an independent numpy restatement of the published Zhang-Suen iteration, and the committed golden vectors.
The reference itself has no tests or fixtures."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden  # noqa: E402

_P8 = C.POINTER(C.c_int8)


def _img(rng, h, w, p):
    return np.where(rng.random((h, w)) < p, 100, 0).astype(np.int8)


def _call(fn, img, *extra):
    out = np.zeros_like(img)
    fn(img.ctypes.data_as(_P8), img.shape[1], img.shape[0], *extra, out.ctypes.data_as(_P8))
    return out


def test_open_matches_cv2(oracle):
    """skeletonizeOccupancyGrid's cv::morphologyEx(MORPH_OPEN, 3x3 MORPH_ELLIPSE) (seed_gen:678-680)."""
    import cv2
    L = oracle.lib()
    k = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (3, 3))
    assert k.tolist() == [[0, 1, 0], [1, 1, 1], [0, 1, 0]]
    rng = np.random.default_rng(0)
    for h, w, p in [(40, 70, 0.7), (64, 64, 0.9), (5, 200, 0.8), (200, 3, 0.8), (1, 1, 1.0), (33, 97, 0.5)]:
        img = _img(rng, h, w, p)
        want = cv2.morphologyEx(np.where(img == 100, 255, 0).astype(np.uint8), cv2.MORPH_OPEN, k)
        got = _call(L.orc_open_cross, img)
        assert np.array_equal(got == 100, want == 255)


def test_inflation_is_disc_dilation_and_edt_threshold(oracle):
    """applyInflation (seed_gen:933-967): integer disc, clipped to the grid == cv2.dilate with that disc ==
    threshold of the exact squared EDT at R^2 (the legitimate use of an exact EDT on this path)."""
    import cv2
    from scipy import ndimage
    L = oracle.lib()
    rng = np.random.default_rng(1)
    for R in (0, 1, 3, 8, 16):
        img = _img(rng, 90, 130, 0.004)
        got = _call(L.orc_inflate, img, R) == 100
        yy, xx = np.mgrid[-R:R + 1, -R:R + 1]
        disc = (xx * xx + yy * yy <= R * R).astype(np.uint8)
        want = cv2.dilate((img == 100).astype(np.uint8), disc, borderType=cv2.BORDER_CONSTANT, borderValue=0) > 0
        assert np.array_equal(got, want)
        if (img == 100).any():
            d2 = np.rint(ndimage.distance_transform_edt(img != 100) ** 2).astype(np.int64)
            assert np.array_equal(got, d2 <= R * R)


def _zhang_suen_numpy(img):
    """Independent restatement of the published algorithm (Zhang & Suen 1984) as opencv_contrib iterates it:
    two parallel sub-iterations per pass, border pixels never removed, until a pass changes nothing."""
    im = (img != 0).astype(np.uint8)
    while True:
        before = im.copy()
        for it in (0, 1):
            p = np.pad(im, 1)
            p2, p3, p4, p5 = p[:-2, 1:-1], p[:-2, 2:], p[1:-1, 2:], p[2:, 2:]
            p6, p7, p8, p9 = p[2:, 1:-1], p[2:, :-2], p[1:-1, :-2], p[:-2, :-2]
            seq = [p2, p3, p4, p5, p6, p7, p8, p9, p2]
            A = sum(((seq[i] == 0) & (seq[i + 1] == 1)).astype(np.int32) for i in range(8))
            B = sum(s.astype(np.int32) for s in seq[:8])
            m1 = (p2 * p4 * p6) if it == 0 else (p2 * p4 * p8)
            m2 = (p4 * p6 * p8) if it == 0 else (p2 * p6 * p8)
            rem = (A == 1) & (B >= 2) & (B <= 6) & (m1 == 0) & (m2 == 0) & (im == 1)
            rem[0, :] = rem[-1, :] = False
            rem[:, 0] = rem[:, -1] = False
            im = im & ~rem.astype(np.uint8)
        if np.array_equal(im, before):
            return im


def test_thinning_against_independent_numpy_and_properties(oracle):
    from scipy import ndimage
    L = oracle.lib()
    rng = np.random.default_rng(2)
    for trial in range(6):
        h, w = int(rng.integers(20, 90)), int(rng.integers(20, 140))
        base = ndimage.binary_dilation(rng.random((h, w)) < 0.01, iterations=int(rng.integers(2, 7)))
        if trial == 0:
            base[:] = True    # fully occupied image: only the border ring constraint matters
        img = np.where(base, 100, 0).astype(np.int8)
        thin = img.copy()
        L.orc_thin_zhangsuen(thin.ctypes.data_as(_P8), w, h)
        assert np.array_equal(thin == 100, _zhang_suen_numpy(img) == 1)
        assert not ((thin == 100) & (img != 100)).any()           # subset of the input
        again = thin.copy()
        L.orc_thin_zhangsuen(again.ctypes.data_as(_P8), w, h)
        assert np.array_equal(again, thin)                          # idempotent
        s8 = np.ones((3, 3), int)
        assert ndimage.label(img == 100, s8)[1] == ndimage.label(thin == 100, s8)[1]   # topology: component count kept


def test_clusters_against_scipy_and_brute_force(oracle):
    from scipy import ndimage
    from aos_gpu import synth
    spec = synth.config("SMALL", seed=9)
    pts = synth.make_orchard(spec)
    p = oracle.SeedParams(grid_resolution=spec.grid_resolution, inflation_radius=spec.inflation_radius, polygon=spec.polygon,
                          cluster_min_length=0.0)
    r = oracle.seed_stage(p, pts)
    lab = r["labels"]
    # cells inside the polygon (seed_gen:998-1004), then 8-connected components, canonical label = min linear index
    ys, xs = np.nonzero(r["skel"] == 100)
    res = np.float32(r["res"])
    wx = (r["origin_x"] + (xs.astype(np.float32) * res).astype(np.float64)).astype(np.float32)
    wy = (r["origin_y"] + (ys.astype(np.float32) * res).astype(np.float64)).astype(np.float32)
    L = oracle.lib()
    poly = np.ascontiguousarray(spec.polygon, np.float64)
    inside = np.array([L.orc_point_in_polygon(float(a), float(b), poly.ctypes.data_as(C.POINTER(C.c_double)), len(poly))
                       for a, b in zip(wx, wy)], bool)
    mask = np.zeros(lab.shape, bool)
    mask[ys[inside], xs[inside]] = True
    cc, n = ndimage.label(mask, np.ones((3, 3), int))
    assert n == r["n_clusters"]
    lin = np.arange(lab.size).reshape(lab.shape)
    want = np.full(lab.shape, -1, np.int32)
    for c in range(1, n + 1):
        want[cc == c] = lin[cc == c].min()
    assert np.array_equal(lab, want)
    # per-cluster statistics by brute force
    for i in range(r["n_clusters"]):
        cy, cx = np.nonzero(lab == r["cl_first"][i])
        assert len(cx) == r["cl_size"][i] and cx.sum() == r["cl_sumx"][i] and cy.sum() == r["cl_sumy"][i]
        d2 = (cx[:, None] - cx[None, :]) ** 2 + (cy[:, None] - cy[None, :]) ** 2
        assert d2.max() == r["cl_maxd2"][i]
        assert r["cl_len"][i] == np.float32(np.sqrt(float(d2.max())) * float(res)) if d2.max() > 0 else r["cl_len"][i] == 0


def test_binning_against_numpy(oracle):
    """generateOccupancyGrid (seed_gen:581-622): float32 limits, double index arithmetic, truncation."""
    from aos_gpu import synth
    spec = synth.config("TINY", seed=3)
    pts = synth.make_orchard(spec)
    p = oracle.SeedParams(grid_resolution=spec.grid_resolution, inflation_radius=spec.inflation_radius, polygon=spec.polygon)
    r = oracle.seed_stage(p, pts)
    poly = spec.polygon
    minx, maxx = np.float32(poly[:, 0].min() - 2.5), np.float32(poly[:, 0].max() + 2.5)
    miny, maxy = np.float32(poly[:, 1].min() - 2.5), np.float32(poly[:, 1].max() + 2.5)
    x, y, z = pts[:, 0], pts[:, 1], pts[:, 2]
    keep = (z >= np.float32(-0.4)) & (z <= np.float32(0.5)) & (x >= minx) & (x <= maxx) & (y >= miny) & (y <= maxy)
    res = np.float64(np.float32(spec.grid_resolution))
    gx = np.trunc((x[keep].astype(np.float64) - np.float64(minx)) / res).astype(np.int64)
    gy = np.trunc((y[keep].astype(np.float64) - np.float64(miny)) / res).astype(np.int64)
    ok = (gx >= 0) & (gx < r["w"]) & (gy >= 0) & (gy < r["h"])
    want = np.zeros((r["h"], r["w"]), bool)
    want[gy[ok], gx[ok]] = True
    assert np.array_equal(r["occ_raw"] == 100, want)


@pytest.mark.parametrize("name", sorted(make_golden.CASES))
def test_oracle_reproduces_golden_vectors(oracle, name):
    import hashlib
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    spec, pts, pk = make_golden.case_inputs(name)
    assert np.array_equal(np.frombuffer(hashlib.sha256(pts.tobytes()).digest(), np.uint8), g["points_sha"]), \
        "synthetic generator drifted: regenerate tests/golden with make_golden.py"
    r = oracle.seed_stage(oracle.SeedParams(**pk), pts)
    assert (r["w"], r["h"]) == (int(g["w"]), int(g["h"]))
    for k in make_golden.GRIDS:
        assert np.array_equal(np.packbits(r[k] == 100, axis=1, bitorder="little"), g[k]), k
    for k in make_golden.SEED_KEYS:
        assert np.array_equal(r[k], g[k]), k
    gr = oracle.gvd_stage(r["seeds"], r["skel_framed"], r["origin_x"], r["origin_y"], r["res"], r["rows_info"])
    for k in make_golden.GRAPH_KEYS:
        assert np.array_equal(gr[k], g["g_" + k]), k


def test_trim_path_oracle_against_vectorised_restatement(oracle):
    """orc_trim_path (path_gen:1570-1630) vs an independent numpy statement of the same rule: the set of stencil
    offsets with sqrt(dx^2+dy^2)*res <= d, cell = trunc((p + offset*res - origin)/res), first pose i > 0 that hits."""
    rng = np.random.default_rng(5)
    for trial in range(40):
        h, w = int(rng.integers(8, 90)), int(rng.integers(8, 90))
        res = float(np.float32([0.05, 0.1, 0.025][trial % 3]))
        ox, oy = [(0.0, 0.0), (-4.5, -2.4)][trial % 2]
        g = np.zeros((h, w), np.int8)
        g[rng.random((h, w)) < 0.004] = 100
        n = int(rng.integers(1, 200))
        path = np.array([ox + w * res / 2, oy + h * res / 2]) + np.cumsum(rng.normal(0, res, (n, 2)), axis=0)
        d = 0.2
        rc = int(np.ceil(d / res))
        off = [(dx, dy) for dx in range(-rc, rc + 1) for dy in range(-rc, rc + 1) if np.sqrt(float(dx * dx + dy * dy)) * res <= d]
        want = n
        for i in range(1, n):
            hit = False
            for dx, dy in off:
                mx = int(np.trunc(((path[i, 0] + dx * res) - ox) / res))
                my = int(np.trunc(((path[i, 1] + dy * res) - oy) / res))
                if 0 <= mx < w and 0 <= my < h and g[my, mx] == 100:
                    hit = True
                    break
            if hit:
                want = i
                break
        assert oracle.trim_path(path, g, ox, oy, res, d) == want


def test_trim_path_oracle_reproduces_golden_counts(oracle):
    """Drift check: orc_trim_path on the golden skeletons gives the committed pose counts (tests/golden/trim_paths.npz)."""
    import os
    import sys
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sys.path.insert(0, here)
    import make_golden
    trim = np.load(os.path.join(here, "trim_paths.npz"))
    for name in sorted(make_golden.CASES):
        g = np.load(os.path.join(here, name + ".npz"))
        skel = np.unpackbits(g["skel_framed"], axis=1, bitorder="little")[:, :int(g["w"])].astype(np.int8) * 100
        got = [oracle.trim_path(p, skel, float(g["origin_x"]), float(g["origin_y"]), g["res"], 0.2) for p in make_golden.trim_paths(name, g)]
        assert got == list(trim[name])
        assert any(1 < k < len(p) for k, p in zip(got, make_golden.trim_paths(name, g)))   # some paths are really cut mid-way
