"""Row-band sharding on real GPUs (needs >= 2; run with `gpurun --gpus 2`): the raster stages split over the
ranks with NCCL halo exchange ("nccl") or with the exchange fused into the thinning kernel as stores into the
neighbour's peer-mapped buffer ("p2p") must reproduce the single-GPU result bit for bit -- every local grid on its band
rows, and on rank 0 the gathered skeleton / occupancy, clusters, rows, seeds and the GvdGraph arrays."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, name, seed, npts, halo, q):
    import faulthandler
    import sys
    faulthandler.dump_traceback_later(100, exit=True, file=sys.stderr)   # a hang must not eat the GPU budget
    import torch
    import torch.distributed as dist
    from aos_gpu import bands, lib, synth
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
      try:
          spec = synth.config(name, seed=seed, n_points=npts)
          pts = synth.make_orchard(spec)
          params = lib.SeedParams(grid_resolution=spec.grid_resolution, inflation_radius=spec.inflation_radius,
                                  polygon=spec.polygon, exclusion=spec.exclusion)
          ref = lib.Context(rank)                      # single-GPU reference on this rank's own GPU
          ref.map_to_graph(params, pts)
          ctx = lib.Context(rank)
          info = bands.banded_map_to_graph(ctx, params, pts, rank, world, dist, rank, halo=halo)
          if halo == "p2p":   # a second map on the same contexts: cached peer mappings, buffers back to index 0
              info = bands.banded_map_to_graph(ctx, params, pts, rank, world, dist, rank, halo=halo)
          # local grids on the band rows
          gi = lib.grid_geometry(params)
          band = bands.band_for(gi.height, world, rank, ctx.band_halo_rows(params))
          errs = []
          if rank != 0:   # rank 0's context has moved on to the full grid (seed_stage_tail)
              be = bands.LibBackend(ctx)
              for gid, nm in ((lib.GRID_RAW, "raw"), (lib.GRID_INFLATED, "inflated"), (lib.GRID_OCCUPANCY, "occupancy"),
                              (lib.GRID_OPENED, "opened"), (lib.GRID_SKELETON, "skeleton")):
                  loc = be.grid(gid)[band.halo_lo:band.halo_lo + band.rows].cpu().numpy().view(np.uint32)
                  want = ref.grid_bits(gid)[band.row0:band.row0 + band.rows]
                  if not np.array_equal(loc, want):
                      errs.append(f"rank {rank} {nm}: {int((loc != want).sum())} words differ")
          else:
              for gid, nm in ((lib.GRID_OCCUPANCY, "occupancy"), (lib.GRID_SKELETON, "skeleton"), (lib.GRID_SKELETON_FRAMED, "framed")):
                  if not np.array_equal(ctx.grid_bits(gid), ref.grid_bits(gid)):
                      errs.append(f"gathered {nm} differs")
              a, b = ctx.clusters(), ref.clusters()
              if a.tobytes() != b.tobytes():
                  errs.append("clusters differ")
              if ctx.tree_rows().tobytes() != ref.tree_rows().tobytes():
                  errs.append("rows differ")
              s1, c1, r1 = ctx.select_seeds()
              s2, c2, r2 = ref.select_seeds()
              if not (np.array_equal(s1, s2) and c1 == c2 and np.array_equal(r1, r2)):
                  errs.append("seeds differ")
              g1, g2 = info["graph"], ref.graph()
              for k in ("nodes", "edges", "edge_lengths", "node_labels", "node_label_clusters", "node_label_types", "corner_points"):
                  if not np.array_equal(g1[k], g2[k]):
                      errs.append(f"graph {k} differs")
              errs.append(f"ok launches={info['thin_launches']} nodes={g1['n_nodes']}") if not errs else None
          q.put((rank, errs))
          dist.barrier()
      except Exception:
        import traceback
        q.put((rank, ["worker raised: " + traceback.format_exc()]))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,name,seed,npts,halo", [(1, "SMALL", 1, None, "nccl"), (2, "SMALL", 1, None, "nccl"),
                                                       (2, "C2", 0, 600_000, "nccl"), (4, "C2", 2, 600_000, "nccl"),
                                                       (1, "SMALL", 1, None, "p2p"), (2, "SMALL", 1, None, "p2p"),
                                                       (2, "C2", 0, 600_000, "p2p"), (4, "C2", 2, 600_000, "p2p")])
def test_banded_equals_single_gpu(world, name, seed, npts, halo):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, seed, npts, halo, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=150) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for r in range(world):
        bad = [e for e in res[r] if not e.startswith("ok")]
        assert not bad, bad
    assert any(e.startswith("ok") for e in res[0])
