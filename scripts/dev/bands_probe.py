"""Row-band sharding probe on N GPUs (torchrun): ONE global strip-generated cloud; rank 0 also runs the single-GPU path
on the full cloud; the band results (both halo modes, twice each) must give the same digest.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
      scripts/dev/bands_probe.py [workload] [points]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "active-orchard-slam_b200")]
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from aos_gpu import bands, lib, synth  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "C4"
npts = int(sys.argv[2]) if len(sys.argv) > 2 else None
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
spec = synth.config(wl, seed=0, n_points=npts)
params = lib.SeedParams(grid_resolution=spec.grid_resolution, inflation_radius=spec.inflation_radius, polygon=spec.polygon)
gi = lib.grid_geometry(params)
ctx = lib.Context(local)
band = bands.band_for(gi.height, world, rank, ctx.band_halo_rows(params))
res = float(np.float32(spec.grid_resolution))
ylo = gi.origin_y + band.first_global_row * res - 2 * res
yhi = gi.origin_y + (band.first_global_row + band.local_rows) * res + 2 * res
pts = synth.make_orchard_strips_torch(spec, dev, y_range=(ylo, yhi) if world > 1 else None)
out = {"workload": wl, "world": world, "points_rank0": int(pts.shape[0])}
if rank == 0:
    full = synth.make_orchard_strips_torch(spec, dev) if world > 1 else pts
    ref = lib.Context(local)
    ref.map_to_graph(params, full)
    out["single"] = ref.result_digest()
    ref.map_to_graph(params, full)
    out["single_again"] = ref.result_digest()
    out["total_points"] = int(full.shape[0])
    g = ref.graph()
    out["single_graph"] = [int(g["n_nodes"]), int(g["n_edges"])]
    ref.close()
    del full
    torch.cuda.empty_cache()
for halo in ("nccl", "p2p", "nccl", "p2p"):
    info = bands.banded_map_to_graph(ctx, params, pts, rank, world, dist, local, halo=halo)
    if rank == 0:
        d, per = ctx.result_digest(parts=True)
        out.setdefault(halo, []).append(d)
        out.setdefault(halo + "_graph", []).append([int(info["graph"]["n_nodes"]), int(info["graph"]["n_edges"])] if info["graph"] else None)
        out.setdefault(halo + "_launches", []).append(info["thin_launches"])
        if d != out["single"]:
            out.setdefault(halo + "_diff_parts", []).append([k for k in per if per[k] != None][:0])
    if world > 1:
        dist.barrier()
if rank == 0:
    out["all_equal"] = all(d == out["single"] for h in ("nccl", "p2p") for d in out[h]) and out["single"] == out["single_again"]
    print(json.dumps(out), flush=True)
if world > 1:
    dist.destroy_process_group()
