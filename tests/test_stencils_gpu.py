"""The stand-alone raster steps (aos_inflate_bits / aos_open_bits / aos_thin_bits / pack / unpack) against the
oracle's loops on random images that are nothing like an orchard: noise at several densities, widths that are not
multiples of 32 or of the 28-word tile, single rows / columns, every radius up to the stencil limit."""
import ctypes as C

import numpy as np
import pytest

from aos_gpu import lib

pytestmark = pytest.mark.gpu
P8 = C.POINTER(C.c_int8)


def _orc(oracle, name, img, *extra):
    src = np.where(img, 100, 0).astype(np.int8)
    out = np.zeros_like(src)
    getattr(oracle.lib(), name)(src.ctypes.data_as(P8), img.shape[1], img.shape[0], *extra, out.ctypes.data_as(P8))
    return out == 100


def _dev(img):
    import torch
    return torch.from_numpy(lib.pack_bits(img).view(np.int32)).cuda()


def _host(t, w):
    return lib.unpack_bits(t.cpu().numpy().view(np.uint32), w)


SHAPES = [(1, 1), (1, 70), (70, 1), (7, 31), (9, 33), (40, 897), (129, 896), (113, 500), (300, 1000), (257, 2049)]


@pytest.mark.parametrize("h,w", SHAPES)
def test_open_random(gpu_ctx, oracle, h, w):
    import torch
    rng = np.random.default_rng(h * 1000 + w)
    for p in (0.3, 0.7, 0.95):
        img = rng.random((h, w)) < p
        src = _dev(img)
        dst = torch.zeros_like(src)
        assert gpu_ctx.L.aos_open_bits(gpu_ctx.h, src.data_ptr(), dst.data_ptr(), w, h) == 0
        assert gpu_ctx.L.aos_synchronize(gpu_ctx.h) == 0
        assert np.array_equal(_host(dst, w), _orc(oracle, "orc_open_cross", img)), (h, w, p)


@pytest.mark.parametrize("h,w", SHAPES)
def test_inflate_random(gpu_ctx, oracle, h, w):
    import torch
    rng = np.random.default_rng(h * 7 + w)
    for R in (0, 1, 2, 7, 16, 31, 64):
        img = rng.random((h, w)) < 0.01
        img[rng.integers(0, h), rng.integers(0, w)] = True
        src = _dev(img)
        dst, border = torch.zeros_like(src), torch.zeros_like(src)
        assert gpu_ctx.L.aos_inflate_bits(gpu_ctx.h, src.data_ptr(), dst.data_ptr(), border.data_ptr(), w, h, R) == 0
        assert gpu_ctx.L.aos_synchronize(gpu_ctx.h) == 0
        want = _orc(oracle, "orc_inflate", img, R)
        assert np.array_equal(_host(dst, w), want), (h, w, R)
        src8 = np.where(want, 100, 0).astype(np.int8)
        wb = np.zeros_like(src8)
        oracle.lib().orc_mark_borders(src8.ctypes.data_as(P8), w, h, wb.ctypes.data_as(P8))
        assert np.array_equal(_host(border, w), wb == 100), (h, w, R, "5-cell frame")


@pytest.mark.parametrize("h,w", SHAPES)
def test_thin_random(gpu_ctx, oracle, h, w):
    from scipy import ndimage
    rng = np.random.default_rng(h + 31 * w)
    for kind in range(3):
        if kind == 0:
            img = rng.random((h, w)) < 0.6                     # noise
        elif kind == 1:
            img = ndimage.binary_dilation(rng.random((h, w)) < 0.02, iterations=4)   # blobs
        else:
            img = np.ones((h, w), bool)                          # solid: border rows / columns are never removed
        src = _dev(img)
        la, sb = C.c_int32(), C.c_int32()
        assert gpu_ctx.L.aos_thin_bits(gpu_ctx.h, src.data_ptr(), w, h, C.byref(la), C.byref(sb)) == 0
        g = np.where(img, 100, 0).astype(np.int8)
        oracle.lib().orc_thin_zhangsuen(g.ctypes.data_as(P8), w, h)
        assert np.array_equal(_host(src, w), g == 100), (h, w, kind)


def test_pack_unpack_roundtrip(gpu_ctx):
    import torch
    rng = np.random.default_rng(5)
    for h, w in [(3, 5), (17, 64), (33, 1001), (100, 2048)]:
        img8 = np.where(rng.random((h, w)) < 0.4, 100, 0).astype(np.int8)
        img8[rng.random((h, w)) < 0.05] = -1          # "unknown" cells of an OccupancyGrid are free (seed_gen:626-644)
        bits = torch.zeros((h, gpu_ctx.L.aos_bits_pitch_words(w)), dtype=torch.int32, device="cuda")
        assert gpu_ctx.L.aos_pack_int8(gpu_ctx.h, img8.ctypes.data_as(C.c_void_p), lib.AOS_MEM_HOST, bits.data_ptr(), w, h) == 0
        assert np.array_equal(_host(bits, w), img8 == 100)
        back = np.zeros((h, w), np.int8)
        assert gpu_ctx.L.aos_unpack_int8(gpu_ctx.h, bits.data_ptr(), back.ctypes.data_as(C.c_void_p), lib.AOS_MEM_HOST, w, h) == 0
        assert np.array_equal(back, np.where(img8 == 100, 100, 0))
