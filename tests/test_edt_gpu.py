"""Exact EDT kernels (k_edt.cu): squared distances against scipy's exact transform, nearest-site labels
consistent with them, and inflation as the EDT threshold d^2 <= R^2 identical to the stencil kernel and to the
oracle's applyInflation loop -- also at BASELINE's full 20000 x 20000 size (size-independent property)."""
import ctypes as C

import numpy as np
import pytest

from aos_gpu import lib

pytestmark = pytest.mark.gpu


def _edt(ctx, img):
    import torch
    h, w = img.shape
    bits = torch.from_numpy(lib.pack_bits(img).view(np.int32)).cuda()
    nearest = torch.empty((h, w), dtype=torch.int32, device="cuda")
    d2 = torch.empty((h, w), dtype=torch.int32, device="cuda")
    rc = ctx.L.aos_edt_bits(ctx.h, bits.data_ptr(), w, h, nearest.data_ptr(), d2.data_ptr())
    assert rc == 0, ctx.L.aos_last_error(ctx.h)
    return nearest.cpu().numpy().view(np.uint32), d2.cpu().numpy()


@pytest.mark.parametrize("h,w,p,seed", [(37, 61, 0.02, 0), (128, 300, 0.002, 1), (257, 95, 0.3, 2), (64, 64, 0.0, 3),
                                        (1, 200, 0.05, 4), (200, 1, 0.05, 5), (500, 700, 0.0005, 6)])
def test_edt_matches_scipy(gpu_ctx, h, w, p, seed):
    from scipy import ndimage
    rng = np.random.default_rng(seed)
    img = rng.random((h, w)) < p
    if p == 0.0:
        img[:] = False
        img[h // 3, w // 2] = True
    nearest, d2 = _edt(gpu_ctx, img)
    want = np.rint(ndimage.distance_transform_edt(~img) ** 2).astype(np.int64)
    assert np.array_equal(d2.astype(np.int64), want)
    nx, ny = (nearest & 0xffff).astype(np.int64), (nearest >> 16).astype(np.int64)
    assert img[ny, nx].all()                                     # every label is a site ...
    yy, xx = np.mgrid[0:h, 0:w]
    assert np.array_equal((xx - nx) ** 2 + (yy - ny) ** 2, want)  # ... at exactly the minimal distance


def test_edt_empty_grid(gpu_ctx):
    nearest, d2 = _edt(gpu_ctx, np.zeros((20, 50), bool))
    assert (nearest == 0xffffffff).all() and (d2 == 0x7fffffff).all()


def _inflate_both(ctx, bits, w, h, R):
    import torch
    a, b, border = torch.zeros_like(bits), torch.zeros_like(bits), torch.zeros_like(bits)
    L = ctx.L
    if R <= 64:
        assert L.aos_inflate_bits(ctx.h, bits.data_ptr(), a.data_ptr(), border.data_ptr(), w, h, R) == 0
        assert L.aos_synchronize(ctx.h) == 0
    assert L.aos_inflate_bits_edt(ctx.h, bits.data_ptr(), b.data_ptr(), w, h, R) == 0, L.aos_last_error(ctx.h)
    return (a if R <= 64 else None), b


@pytest.mark.parametrize("R", [0, 1, 5, 16, 32, 100])
def test_inflation_equals_edt_threshold(gpu_ctx, oracle, R):
    import torch
    rng = np.random.default_rng(R)
    h, w = 300, 417
    img = rng.random((h, w)) < 0.003
    bits = torch.from_numpy(lib.pack_bits(img).view(np.int32)).cuda()
    a, b = _inflate_both(gpu_ctx, bits, w, h, R)
    got = lib.unpack_bits(b.cpu().numpy().view(np.uint32), w)
    src = np.where(img, 100, 0).astype(np.int8)
    want = np.zeros_like(src)
    P8 = C.POINTER(C.c_int8)
    oracle.lib().orc_inflate(src.ctypes.data_as(P8), w, h, R, want.ctypes.data_as(P8))
    assert np.array_equal(got, want == 100)
    if a is not None:
        assert torch.equal(a, b)


def test_inflation_equals_edt_threshold_full_size(gpu_ctx):
    """BASELINE config 3 size (20000 x 20000 cells, R = 16): the two independent kernels agree on every word."""
    import torch
    from aos_gpu import synth
    spec = synth.config("C3", n_points=20_000_000)
    pts = synth.make_orchard_torch(spec, "cuda")
    params = lib.SeedParams(grid_resolution=spec.grid_resolution, inflation_radius=spec.inflation_radius, polygon=spec.polygon)
    gpu_ctx.seed_stage(params, pts)
    del pts
    s = gpu_ctx.seed_summary()
    w, h = s.info.width, s.info.height
    raw_ptr, pitch = gpu_ctx.grid_device_bits(lib.GRID_RAW)
    inf_ptr, _ = gpu_ctx.grid_device_bits(lib.GRID_INFLATED)
    out = torch.zeros((h, pitch), dtype=torch.int32, device="cuda")
    assert gpu_ctx.L.aos_inflate_bits_edt(gpu_ctx.h, raw_ptr, out.data_ptr(), w, h, 16) == 0, gpu_ctx.L.aos_last_error(gpu_ctx.h)
    from aos_gpu.bands import _CudaArray
    stencil = torch.as_tensor(_CudaArray(inf_ptr, (h, pitch)), device="cuda")
    assert torch.equal(out, stencil)
    assert int((out != 0).sum()) > 1000


def test_opt_in_edge_clearance(oracle):
    """aos_set_clearance: edge_clearances = min over the reference's edge samples of the exact distance to the
    framed skeleton; everything else stays bit-identical, and the default (off) stays 0.0f."""
    from scipy import ndimage
    from aos_gpu import synth
    spec = synth.config("SMALL", seed=2)
    pts = synth.make_orchard(spec)
    params = lib.SeedParams(grid_resolution=spec.grid_resolution, inflation_radius=spec.inflation_radius, polygon=spec.polygon)
    ctx = lib.Context(0)
    ctx.map_to_graph(params, pts)
    g0 = ctx.graph()
    assert (g0["edge_clearances"] == 0.0).all()
    ctx.set_clearance(True)
    ctx.map_to_graph(params, pts)
    g1 = ctx.graph()
    for k in ("nodes", "edges", "edge_lengths", "node_labels"):
        assert np.array_equal(g0[k], g1[k])
    skel = ctx.grid_int8(lib.GRID_SKELETON_FRAMED) == 100
    d2 = np.rint(ndimage.distance_transform_edt(~skel) ** 2).astype(np.int64)
    s = ctx.seed_summary()
    res, ox, oy = float(np.float32(s.info.resolution)), s.info.origin_x, s.info.origin_y
    h, w = skel.shape
    for e in range(0, len(g1["edges"]), 7):
        a, b = g1["nodes"][g1["edges"][e, 0]], g1["nodes"][g1["edges"][e, 1]]
        ex, ey = b[0] - a[0], b[1] - a[1]
        L = np.sqrt(ex * ex + ey * ey)
        n = int(L / (res * 0.5)) + 1
        best = None
        for i in range(n + 1):
            t = 1.0 if i == n else i / n
            px, py = a[0] + (t * (ex / L)) * L, a[1] + (t * (ey / L)) * L
            mx, my = int((px - ox) / res), int((py - oy) / res)
            if 0 <= mx < w and 0 <= my < h:
                best = d2[my, mx] if best is None else min(best, d2[my, mx])
        want = np.float32(np.sqrt(float(best)) * res)
        assert g1["edge_clearances"][e] == want, (e, g1["edge_clearances"][e], want)
    assert (g1["edge_clearances"] > 0).all()      # published edges never touch the skeleton (gvd:320-359)
    ctx.close()
