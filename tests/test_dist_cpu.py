"""N > 1 host logic on CPU (gloo, world_size 2): map ownership covers every map exactly once, the timing
reduction is a MAX, the cell count a SUM -- the only cross-rank traffic of the path (one map per GPU)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from aos_gpu import dist as adist


def _worker(rank, world, port, n_maps, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = adist.map_assignment(n_maps, world, rank)
    cells = sum(1000 * (i + 1) for i in mine)           # stand-in for W*H of each owned map
    ms = 10.0 * (rank + 1) + len(mine)                   # ranks finish at different times
    g_ms, g_cells = adist.reduce_stats(ms, cells)
    q.put((rank, mine, ms, cells, g_ms, g_cells))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_gloo():
    world, n_maps = 2, 7
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_maps, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    owned = sorted(i for r in res for i in r[1])
    assert owned == list(range(n_maps))                                  # every map exactly once
    assert res[0][1] == [0, 2, 4, 6] and res[1][1] == [1, 3, 5]
    want_ms, want_cells = max(r[2] for r in res), sum(r[3] for r in res)
    for r in res:
        assert r[4] == want_ms and r[5] == want_cells                    # MAX of times, SUM of cells on every rank
    assert adist.throughput_mcells(want_cells, want_ms) == want_cells / (want_ms * 1e-3) / 1e6


def test_single_process_passthrough():
    assert adist.reduce_stats(12.5, 400) == (12.5, 400)
    assert adist.map_assignment(5, 1, 0) == [0, 1, 2, 3, 4]
    assert adist.map_assignment(3, 8, 5) == []
