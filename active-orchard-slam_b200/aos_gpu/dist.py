"""Multi-GPU plumbing of the path: it shards by INDEPENDENT MAPS (BASELINE.json config 5: a sweep of maps,
one map per GPU at a time), so there is no data-path collective -- only the bookkeeping below: which rank
owns which map, and max-over-ranks / sum-over-ranks of what each rank measured.  Works on NCCL (GPU tensors)
and on gloo (CPU tensors; tests/test_dist_cpu.py runs it with world_size 2)."""
from __future__ import annotations

import os


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def map_assignment(n_maps: int, world: int, rank: int) -> list[int]:
    """Round-robin: map i is processed by rank i mod world."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    return list(range(rank, n_maps, world))


def reduce_stats(local_ms: float, local_cells: int, device=None):
    """(max over ranks of the elapsed ms, sum over ranks of the processed cells)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(local_ms), int(local_cells)
    t = torch.tensor([float(local_ms)], dtype=torch.float64, device=device)
    c = torch.tensor([int(local_cells)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    return float(t.item()), int(c.item())


def throughput_mcells(total_cells: int, elapsed_ms: float) -> float:
    return total_cells / (elapsed_ms * 1e-3) / 1e6
