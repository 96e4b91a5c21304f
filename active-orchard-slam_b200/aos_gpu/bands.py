"""Row-band sharding of the raster stages over several GPUs (BASELINE.json config 4).

One process per GPU.  Rank r owns the image rows [row0, row0 + rows) of the global grid plus a halo.  Binning,
inflation and opening need no communication when the halo covers their stencil reach
(aos_band_halo_rows = R + 2 + 8 rows; every rank bins the points of its local rows).  Thinning is an iteration
with a global fixed point: before every launch (8 sub-iterations) each rank refreshes the 8 halo rows next to its
band from the neighbour that owns them (point-to-point send/recv: NCCL over NVLink on GPUs, gloo in the CPU test),
and an all-reduce(MAX) of the "deleted something" flags decides when to stop.  The bands are then gathered on
rank 0, which finishes the seed stage (aos_seed_stage_tail), selects the seeds and builds the graph: those
stages are O(skeleton cells) and do not shard.

The orchestration below only needs a backend with
    skeleton()      -> 2-D tensor [local_rows, pitch] ALIASING the backend's current thinning image
    thin_launch()   -> bool, True when the owned rows lost a pixel
so tests/test_bands_cpu.py drives it with a numpy backend under gloo, world_size 2.
"""
from __future__ import annotations

from dataclasses import dataclass

THIN_HALO = 8  # AOS_BAND_THIN_HALO


@dataclass
class Band:
    row0: int
    rows: int
    halo_lo: int
    halo_hi: int

    @property
    def local_rows(self) -> int:
        return self.halo_lo + self.rows + self.halo_hi

    @property
    def first_global_row(self) -> int:
        return self.row0 - self.halo_lo


def split_rows(height: int, world: int) -> list[tuple[int, int]]:
    """Contiguous bands, sizes differ by at most one row."""
    base, extra = divmod(height, world)
    out, r = [], 0
    for i in range(world):
        n = base + (1 if i < extra else 0)
        out.append((r, n))
        r += n
    return out


def band_for(height: int, world: int, rank: int, halo: int) -> Band:
    row0, rows = split_rows(height, world)[rank]
    if world > 1 and rows < THIN_HALO:
        raise ValueError("bands must be at least 8 rows high")
    return Band(row0, rows, min(halo, row0), min(halo, height - row0 - rows))


def exchange_thin_halo(skel, band: Band, rank: int, world: int, dist) -> None:
    """Refresh the THIN_HALO rows on either side of the band from the neighbours' band rows (in place)."""
    if world == 1:
        return
    ops, keep = [], []
    lo, hi = band.halo_lo, band.halo_lo + band.rows
    if rank > 0:            # neighbour below owns the rows under my band
        send = skel[lo:lo + THIN_HALO].contiguous()
        recv = skel[lo - THIN_HALO:lo]
        ops += [dist.P2POp(dist.isend, send, rank - 1), dist.P2POp(dist.irecv, recv, rank - 1)]
        keep.append(send)
    if rank < world - 1:    # neighbour above
        send = skel[hi - THIN_HALO:hi].contiguous()
        recv = skel[hi:hi + THIN_HALO]
        ops += [dist.P2POp(dist.isend, send, rank + 1), dist.P2POp(dist.irecv, recv, rank + 1)]
        keep.append(send)
    for req in dist.batch_isend_irecv(ops):
        req.wait()


def run_thinning(backend, band: Band, rank: int, world: int, dist, max_launches: int = 100000) -> int:
    """Thin to the global fixed point; returns the number of launches (identical on every rank)."""
    import torch
    launches = 0
    for _ in range(max_launches):
        skel = backend.skeleton()
        exchange_thin_halo(skel, band, rank, world, dist)
        if skel.is_cuda:
            torch.cuda.current_stream().synchronize()   # the library launches on its own stream
        deleted = bool(backend.thin_launch())
        launches += 1
        flag = torch.tensor([1 if deleted else 0], dtype=torch.int32, device=skel.device)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        if int(flag.item()) == 0:
            break
    return launches


def run_thinning_p2p(backend, band: Band, rank: int, world: int, dist, device=None, max_launches: int = 100000) -> int:
    """Thinning with the halo exchange FUSED into the kernel (aos_band_thin_launch_p2p): no send/recv.  Every rank
    publishes handles of its two ping-pong buffers, maps those of its neighbours (CUDA IPC: peer memory over
    NVLink), and each launch stores the band's edge rows straight into the neighbour's destination buffer.  The
    only collective left is the all-reduce(MAX) of the "deleted" flags, which also orders a launch's peer stores
    before the neighbour's next launch reads them.

    Backend protocol (the numpy/shared-memory backend of tests/test_bands_cpu.py implements the same):
        export_handle(buffer) -> bytes (64)     import_handle(side, buffer, handle, peer_first_global_row)
        thin_launch_p2p() -> bool
    """
    import torch
    if world > 1:
        mine = torch.tensor(list(backend.export_handle(0) + backend.export_handle(1)) +
                            list(int(band.first_global_row).to_bytes(8, "little", signed=True)),
                            dtype=torch.uint8, device=device)
        allh = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allh, mine)
        for side, peer in ((0, rank - 1), (1, rank + 1)):
            if 0 <= peer < world:
                raw = bytes(allh[peer].cpu().tolist())
                first = int.from_bytes(raw[128:136], "little", signed=True)
                backend.import_handle(side, 0, raw[:64], first)
                backend.import_handle(side, 1, raw[64:128], first)
        # the all_gather doubles as the barrier this needs: a rank leaves it only after EVERY rank has entered, i.e.
        # finished aos_band_raster, so nobody's buffers are still being written when the first peer stores arrive
    launches = 0
    flag = torch.zeros(1, dtype=torch.int32, device=device)
    for _ in range(max_launches):
        deleted = bool(backend.thin_launch_p2p())   # returns after the launch (and its peer stores) completed
        launches += 1
        flag.fill_(1 if deleted else 0)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        if int(flag.item()) == 0:
            break
    return launches


def gather_rows(local, band: Band, height: int, rank: int, world: int, dist):
    """Band rows of every rank -> the full [height, pitch] grid on rank 0 (None elsewhere)."""
    import torch
    mine = local[band.halo_lo:band.halo_lo + band.rows].contiguous()
    if world == 1:
        return mine
    if rank == 0:
        full = torch.empty((height, local.shape[1]), dtype=local.dtype, device=local.device)
        full[band.row0:band.row0 + band.rows] = mine
        reqs = []
        for r, (row0, rows) in enumerate(split_rows(height, world)):
            if r == 0:
                continue
            reqs.append(dist.irecv(full[row0:row0 + rows], r))
        for q in reqs:
            q.wait()
        return full
    dist.send(mine, 0)
    return None


# ---- GPU backend: libaos_gpu contexts ------------------------------------------------------------------
class _CudaArray:
    """Zero-copy view of library-owned device memory for torch.as_tensor (__cuda_array_interface__)."""

    def __init__(self, ptr: int, shape, typestr="<i4"):   # int32 view of the uint32 words (NCCL has no uint32)
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False), "version": 2}


class LibBackend:
    def __init__(self, ctx):
        self.ctx = ctx

    def grid(self, which):
        import torch
        ptr, pitch, rows = self.ctx.band_grid_device(which)
        return torch.as_tensor(_CudaArray(ptr, (rows, pitch)), device=f"cuda:{self.ctx.device}")

    def skeleton(self):
        from . import lib
        return self.grid(lib.GRID_SKELETON)

    def thin_launch(self):
        return self.ctx.band_thin_launch()

    def export_handle(self, buffer):
        return self.ctx.band_ipc_export(buffer)

    def import_handle(self, side, buffer, handle, peer_first_global_row):
        self.ctx.band_ipc_import(side, buffer, handle, peer_first_global_row)

    def thin_launch_p2p(self):
        return self.ctx.band_thin_launch_p2p()


def banded_map_to_graph(ctx, params, points, rank: int, world: int, dist, device_index: int, halo: str = "nccl"):
    """The whole path with the raster stages row-band sharded; results (as after aos_map_to_graph) on rank 0's
    context.  `points`: this rank's points (any superset of the points falling into its local rows).
    halo: "nccl" = send/recv of the halo rows between launches, "p2p" = stores into peer memory from the kernel."""
    import torch
    from . import lib
    gi = lib.grid_geometry(params)
    band = band_for(gi.height, world, rank, ctx.band_halo_rows(params))
    ctx.band_raster(params, band, points)
    be = LibBackend(ctx)
    if halo == "p2p":
        launches = run_thinning_p2p(be, band, rank, world, dist, device=torch.device("cuda", device_index))
    else:
        launches = run_thinning(be, band, rank, world, dist)
    skel = gather_rows(be.skeleton(), band, gi.height, rank, world, dist)
    occ = gather_rows(be.grid(lib.GRID_OCCUPANCY), band, gi.height, rank, world, dist)
    info = None
    if rank == 0:
        torch.cuda.synchronize(device_index)
        ctx.seed_stage_tail(params, skel, occ)
        seeds, counts, rows_info = ctx.select_seeds()
        graph = ctx.gvd_stage(seeds, rows_info) if len(seeds) else None
        info = {"thin_launches": launches, "band": band, "n_seeds": len(seeds), "graph": graph}
    return info
