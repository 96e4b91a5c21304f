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


class SharedMemBackend:
    """The p2p protocol of bands.run_thinning_p2p on CPU: the two ping-pong buffers are POSIX shared-memory blocks
    (the "IPC handle" is the block's name), a launch writes its band's edge rows into the neighbours' destination
    block and leaves its own halo rows next to the band to them -- what thin_kernel does through peer memory."""

    def __init__(self, local, band, gh):
        from multiprocessing import shared_memory
        self.band, self.gh, self.shape = band, gh, local.shape
        self.shm = [shared_memory.SharedMemory(create=True, size=local.size) for _ in range(2)]
        self.buf = [np.ndarray(local.shape, np.uint8, buffer=m.buf) for m in self.shm]
        self.buf[0][:] = local
        self.buf[1][:] = 0xEE                     # never-written rows must not matter
        self.cur = 0
        self.peer, self.peer_shm, self.peer_first = {}, [], {}

    def export_handle(self, buffer):
        return self.shm[buffer].name.encode().ljust(64, b"\0")

    def import_handle(self, side, buffer, handle, peer_first_global_row):
        from multiprocessing import shared_memory
        m = shared_memory.SharedMemory(name=handle.rstrip(b"\0").decode())
        self.peer_shm.append(m)
        rows = len(m.buf) // self.shape[1]
        self.peer[(side, buffer)] = np.ndarray((rows, self.shape[1]), np.uint8, buffer=m.buf[:rows * self.shape[1]])
        self.peer_first[side] = peer_first_global_row

    def skeleton(self):
        return torch.from_numpy(self.buf[self.cur])

    def thin_launch_p2p(self):
        b, K = self.band, bands.THIN_HALO
        nxt = self.cur ^ 1
        out, deleted = _subiters(self.buf[self.cur], K, b.first_global_row, self.gh)
        keep = np.ones(self.shape[0], bool)
        lo, hi = b.halo_lo, b.halo_lo + b.rows
        if b.halo_lo:
            keep[lo - K:lo] = False
            d = self.peer[(0, nxt)]
            r = b.first_global_row + lo - self.peer_first[0]
            d[r:r + K] = out[lo:lo + K]
        if b.halo_hi:
            keep[hi:hi + K] = False
            d = self.peer[(1, nxt)]
            r = b.first_global_row + hi - K - self.peer_first[1]
            d[r:r + K] = out[hi - K:hi]
        self.buf[nxt][keep] = out[keep]
        self.cur = nxt
        return bool(deleted[lo:hi].any())

    def close(self):
        self.peer.clear()
        self.buf = None
        for m in self.peer_shm:
            m.close()
        for m in self.shm:
            m.close()
            m.unlink()


def _make_image(h, w, seed):
    from scipy import ndimage
    rng = np.random.default_rng(seed)
    img = ndimage.binary_dilation(rng.random((h, w)) < 0.012, iterations=6)
    img[h // 2 - 9:h // 2 + 9, 5:w - 5] = True      # a thick bar straddling the band boundary
    img[3:h - 3, w // 3:w // 3 + 15] = True         # and one crossing every band
    return img.astype(np.uint8)


def _worker(rank, world, port, h, w, seed, q, mode="sendrecv"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = _make_image(h, w, seed)
    band = bands.band_for(h, world, rank, halo=12)
    a = band.first_global_row
    local = full[a:a + band.local_rows].copy()
    if mode == "p2p":
        be = SharedMemBackend(local, band, h)
        launches = bands.run_thinning_p2p(be, band, rank, world, dist)
    else:
        be = NumpyBackend(local, band, h)
        launches = bands.run_thinning(be, band, rank, world, dist)
    out = bands.gather_rows(be.skeleton().clone(), band, h, rank, world, dist)
    if rank == 0:
        q.put((launches, out.numpy().copy()))
    dist.barrier()
    if mode == "p2p":
        be.close()
    dist.destroy_process_group()


def _run(world, h, w, seed, mode="sendrecv"):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, h, w, seed, q, mode)) for r in range(world)]
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


def test_banded_thinning_p2p_protocol(oracle):
    """Halo rows stored by the neighbour into the destination buffer (ping-pong in lockstep), flags all-reduced."""
    h, w, seed = 150, 120, 5
    want = _oracle_thin(oracle, _make_image(h, w, seed))
    ref_launches = None
    for world in (2, 3):
        launches, out = _run(world, h, w, seed, mode="p2p")
        assert np.array_equal(out, want), f"world {world}: {int((out != want).sum())} cells differ"
        l2, out2 = _run(world, h, w, seed)
        assert launches == l2 and np.array_equal(out2, want)


def test_band_geometry():
    assert bands.split_rows(10, 3) == [(0, 4), (4, 3), (7, 3)]
    b = bands.band_for(1000, 4, 0, 26)
    assert (b.row0, b.rows, b.halo_lo, b.halo_hi) == (0, 250, 0, 26)
    b = bands.band_for(1000, 4, 3, 26)
    assert (b.row0, b.rows, b.halo_lo, b.halo_hi) == (750, 250, 26, 0)
    b = bands.band_for(1000, 4, 2, 26)
    assert b.first_global_row == 474 and b.local_rows == 302
