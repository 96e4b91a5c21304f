"""Row-band sharding logic on CPU (gloo, world_size 2 and 3): the halo exchange / convergence protocol of
aos_gpu/bands.py driven by a numpy backend whose "launch" is 8 Zhang-Suen sub-iterations on the local rows with
the library's band semantics (pixels outside the local grid read as 0, only GLOBAL border rows are protected).
The gathered result must equal the un-banded fixed point -- the oracle's thinning of the whole image."""
import ctypes as C
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from aos_gpu import bands


def _subiters(im, n_sub, y_off, gh):
    """n_sub parallel Zhang-Suen sub-iterations (0,1,0,1,...) on local rows; returns (image, deleted mask)."""
    im = im.copy()
    deleted = np.zeros(im.shape, bool)
    h, w = im.shape
    gy = np.arange(h)[:, None] + y_off
    protect = (gy <= 0) | (gy >= gh - 1) | (np.arange(w)[None, :] == 0) | (np.arange(w)[None, :] == w - 1)
    for s in range(n_sub):
        p = np.pad(im, 1)
        p2, p3, p4, p5 = p[:-2, 1:-1], p[:-2, 2:], p[1:-1, 2:], p[2:, 2:]
        p6, p7, p8, p9 = p[2:, 1:-1], p[2:, :-2], p[1:-1, :-2], p[:-2, :-2]
        seq = [p2, p3, p4, p5, p6, p7, p8, p9, p2]
        A = sum(((seq[i] == 0) & (seq[i + 1] == 1)).astype(np.int32) for i in range(8))
        B = sum(x.astype(np.int32) for x in seq[:8])
        m1 = (p2 * p4 * p6) if s % 2 == 0 else (p2 * p4 * p8)
        m2 = (p4 * p6 * p8) if s % 2 == 0 else (p2 * p6 * p8)
        rem = (A == 1) & (B >= 2) & (B <= 6) & (m1 == 0) & (m2 == 0) & (im == 1) & ~protect
        im = im & ~rem.astype(np.uint8)
        deleted |= rem
    return im, deleted


class NumpyBackend:
    def __init__(self, local, band, gh):
        self.img = torch.from_numpy(local)       # aliases `local`
        self.local, self.band, self.gh = local, band, gh

    def skeleton(self):
        return self.img

    def thin_launch(self):
        out, deleted = _subiters(self.local, bands.THIN_HALO, self.band.first_global_row, self.gh)
        self.local[:] = out
        b = self.band
        return bool(deleted[b.halo_lo:b.halo_lo + b.rows].any())


def _make_image(h, w, seed):
    from scipy import ndimage
    rng = np.random.default_rng(seed)
    img = ndimage.binary_dilation(rng.random((h, w)) < 0.012, iterations=6)
    img[h // 2 - 9:h // 2 + 9, 5:w - 5] = True      # a thick bar straddling the band boundary
    img[3:h - 3, w // 3:w // 3 + 15] = True         # and one crossing every band
    return img.astype(np.uint8)


def _worker(rank, world, port, h, w, seed, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = _make_image(h, w, seed)
    band = bands.band_for(h, world, rank, halo=12)
    a = band.first_global_row
    local = full[a:a + band.local_rows].copy()
    be = NumpyBackend(local, band, h)
    launches = bands.run_thinning(be, band, rank, world, dist)
    out = bands.gather_rows(be.skeleton(), band, h, rank, world, dist)
    if rank == 0:
        q.put((launches, out.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def _run(world, h, w, seed):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, h, w, seed, q)) for r in range(world)]
    for p in procs:
        p.start()
    launches, out = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    return launches, out


def _oracle_thin(oracle, img):
    g = np.where(img != 0, 100, 0).astype(np.int8)
    oracle.lib().orc_thin_zhangsuen(g.ctypes.data_as(C.POINTER(C.c_int8)), img.shape[1], img.shape[0])
    return (g == 100).astype(np.uint8)


def test_banded_thinning_equals_global_fixed_point(oracle):
    h, w, seed = 150, 120, 3
    want = _oracle_thin(oracle, _make_image(h, w, seed))
    # single "band": the protocol degenerates to plain launches
    one = _make_image(h, w, seed)
    b1 = bands.band_for(h, 1, 0, halo=12)
    assert (b1.row0, b1.rows, b1.halo_lo, b1.halo_hi) == (0, h, 0, 0)
    be = NumpyBackend(one, b1, h)
    bands.run_thinning(be, b1, 0, 1, dist)
    assert np.array_equal(one, want)
    for world in (2, 3):
        launches, out = _run(world, h, w, seed)
        assert launches >= 2
        assert np.array_equal(out, want), f"world {world}: {int((out != want).sum())} cells differ"


def test_band_geometry():
    assert bands.split_rows(10, 3) == [(0, 4), (4, 3), (7, 3)]
    b = bands.band_for(1000, 4, 0, 26)
    assert (b.row0, b.rows, b.halo_lo, b.halo_hi) == (0, 250, 0, 26)
    b = bands.band_for(1000, 4, 3, 26)
    assert (b.row0, b.rows, b.halo_lo, b.halo_hi) == (750, 250, 26, 0)
    b = bands.band_for(1000, 4, 2, 26)
    assert b.first_global_row == 474 and b.local_rows == 302
