#!/usr/bin/env python
"""bench.py -- map -> GvdGraph throughput of libaos_gpu on B200 (BASELINE.json metric: Mcells/s, ms/map).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C3|C2|...] [--impl ours|reference]

One "step" = one pass of the hot path over one synthetic orchard map (SURVEY.md section 8(d) generator):
point cloud -> occupancy grid -> inflation -> skeleton -> row clusters -> seeds -> Voronoi/GVD graph.
N > 1 (torchrun, one rank per GPU): every rank processes its own independent map (BASELINE.json config 5,
"one map per GPU"; no data-path collective), value = cells of all ranks / max-over-ranks time, weak scaling.

  value : inputs resident in HBM, CUDA events on the library's stream, max over ranks
  e2e   : the same call with HOST (pinned) points, H2D inside the timed region, and the results
          (published grids bit-packed, clusters, rows, seeds, graph) copied back to the host
  roofline / cpu_baseline / clocks : see DESIGN.md "Measurement"

--impl reference times the CPU restatement of the reference (oracle/, the reference itself cannot be built
here: ROS 2 / PCL / OpenCV-dev are absent) on the host cores, on a bounded crop of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "active-orchard-slam_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

METRIC = "map_to_gvdgraph_throughput"
UNIT = "Mcells/s"


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU while the timed region runs."""

    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None

    def _sample_nvml(self, h, nv):
        """One in-process NVML sample (microseconds of host time; `nvidia-smi` per sample costs a tenth of a second
        of one core, which the maps in flight need)."""
        R = nv.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
            else nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        def bit(*names):
            for n in names:
                v = getattr(nv, n, None)
                if v is not None:
                    return "Active" if (R & v) else "Not Active"
            return "Not Active"
        return [str(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), str(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)),
                bit("nvmlClocksEventReasonHwSlowdown", "nvmlClocksThrottleReasonHwSlowdown"),
                bit("nvmlClocksEventReasonHwThermalSlowdown", "nvmlClocksThrottleReasonHwThermalSlowdown"),
                bit("nvmlClocksEventReasonSwThermalSlowdown", "nvmlClocksThrottleReasonSwThermalSlowdown"),
                bit("nvmlClocksEventReasonSwPowerCap", "nvmlClocksThrottleReasonSwPowerCap")]

    def _run(self):
        h = nv = None
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.index
            if vis:   # NVML enumerates physical devices
                tok = [t for t in vis.split(",") if t.strip()]
                if idx < len(tok) and tok[idx].strip().isdigit():
                    idx = int(tok[idx])
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.index_physical = idx
        except Exception:
            h = None
        while not self._stop.is_set():
            try:
                if h is not None:
                    self.rows.append(self._sample_nvml(h, nv))
                else:
                    out = subprocess.run(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                    if out:
                        self.rows.append([c.strip() for c in out.splitlines()[0].split(",")])
            except Exception:
                h = None   # NVML call failed: fall back to nvidia-smi
            self._stop.wait(0.1 if h is not None else 0.5)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def make_params(lib, spec):
    return lib.SeedParams(grid_resolution=spec.grid_resolution, inflation_radius=spec.inflation_radius,
                          polygon=spec.polygon, exclusion=spec.exclusion)


# ---------------------------------------------------------------------------------------------------
def bench_ours(args):
    import concurrent.futures as cf

    import torch

    from aos_gpu import dist as adist
    from aos_gpu import lib, synth

    rank, world, local = adist.env_rank_world()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; libaos_gpu has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ncpu = os.cpu_count() or 1
    T = args.maps_in_flight if args.maps_in_flight > 0 else max(1, min(16, ncpu // max(world, 1)))
    if args.maps_in_flight <= 0:   # every map in flight holds its cloud twice (resident copy + e2e staging) plus grids
        spec_probe = synth.config(args.workload, n_points=args.points)
        per_map = 2 * 16 * spec_probe.n_points + 1.5e9
        free_b, _ = torch.cuda.mem_get_info(local)
        T = max(1, min(T, int(0.85 * free_b / per_map)))
    gate = (2 if T >= 4 else 0) if args.device_gate < 0 else args.device_gate
    lib.load().aos_set_device_gate(gate)   # maps admitted to the seed stage's kernel phase at a time (0 = no limit)
    spec0 = synth.config(args.workload, seed=rank * 64, n_points=args.points)
    params = make_params(lib, spec0)
    gi = lib.grid_geometry(params)
    cells = gi.width * gi.height
    t0 = time.time()
    # one independent map per stream in flight (different generator seeds), all resident in HBM
    maps = []
    for t in range(T):
        spec = synth.config(args.workload, seed=rank * 64 + t, n_points=args.points)
        maps.append(synth.make_orchard_torch(spec, dev))
    torch.cuda.synchronize()
    n_pts = maps[0].shape[0]
    try:
        host_pts = torch.empty(maps[0].shape, dtype=maps[0].dtype, pin_memory=True)   # e2e source (map 0, pinned)
        pinned = True
    except RuntimeError:
        host_pts = torch.empty(maps[0].shape, dtype=maps[0].dtype)
        pinned = False
    host_pts.copy_(maps[0])
    host_np = host_pts.numpy()
    gen_s = time.time() - t0

    ctxs = [lib.Context(local) for _ in range(T)]     # one context (own stream, own buffers) per map in flight
    pitch = lib.load().aos_bits_pitch_words(gi.width)

    def _result_grids():                               # e2e result buffers (published grids), pinned, reused every step
        out = {}
        for key in ("occ_bits", "skel_bits"):
            try:
                t = torch.empty((gi.height, pitch), dtype=torch.int32, pin_memory=True)
            except RuntimeError:
                t = torch.empty((gi.height, pitch), dtype=torch.int32)
            out[key] = t.numpy().view(np.uint32)
            out["_keep_" + key] = t
        return out
    fetch_bufs = [_result_grids() for _ in range(T)]
    pool = cf.ThreadPoolExecutor(max_workers=T)        # ctypes releases the GIL: host stages run on T cores

    def run_batch(steps, host=False):
        """Every stream processes `steps` maps back to back; returns the last summary of stream 0."""
        def work(t):
            torch.cuda.set_device(local)
            out = None
            for _ in range(steps):
                out = ctxs[t].map_to_graph(params, host_np if host else maps[t], fetch=fetch_bufs[t] if host else None)
            return out
        return list(pool.map(work, range(T)))[0]

    def timed(steps, host=False):
        """CUDA events on the current stream around the whole batch, device idle at both records."""
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        w0 = time.perf_counter()
        info = run_batch(steps, host)
        torch.cuda.synchronize()
        ev1.record()
        ev1.synchronize()
        wall_ms = (time.perf_counter() - w0) * 1e3
        barrier()
        return max(ev0.elapsed_time(ev1), wall_ms), info

    run_batch(args.warmup)
    l0 = sum(c.launch_count() for c in ctxs)
    with ClockSampler(local) as clk:
        dev_ms, info = timed(args.steps)
    launches = (sum(c.launch_count() for c in ctxs) - l0) / max(args.steps, 1)
    run_batch(min(args.warmup, 2), host=True)
    e2e_ms, info_h = timed(args.steps, host=True)

    # ---- one map at a time on one stream: latency per map, per-stage device times ------------------------
    c0 = ctxs[0]
    stream = torch.cuda.Stream(device=dev)
    c0.set_stream(stream.cuda_stream)
    with torch.cuda.stream(stream):
        c0.map_to_graph(params, maps[0])
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        ev0.record(stream)
        reps = 3
        for _ in range(reps):
            c0.map_to_graph(params, maps[0])
        ev1.record(stream)
        torch.cuda.synchronize()
        single_ms = ev0.elapsed_time(ev1) / reps
        c0.set_profiling(True)
        stage_acc = {}
        for _ in range(reps):
            c0.map_to_graph(params, maps[0])
            for name, ms in c0.stage_times():
                stage_acc[name] = stage_acc.get(name, 0.0) + ms / reps
        c0.set_profiling(False)

    dev_ms, total_cells = adist.reduce_stats(dev_ms, cells * args.steps * T, device=dev)   # MAX over ranks, SUM of cells
    e2e_ms, _ = adist.reduce_stats(e2e_ms, cells * args.steps * T, device=dev)
    ms_per_step = dev_ms / args.steps
    value = adist.throughput_mcells(total_cells, dev_ms)
    e2e_value = adist.throughput_mcells(total_cells, e2e_ms)

    line = None
    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        # dominant HBM kernel by algorithmic bytes: the point-binning pass reads 16 B per point once
        bin_ms = stage_acc.get("bin", float("nan"))
        bin_bytes = 16.0 * n_pts
        achieved = bin_bytes / (bin_ms * 1e-3) / 1e9 if bin_ms == bin_ms and bin_ms > 0 else None
        b_alg = 16.0 * n_pts + 9.125 * cells  # compulsory bytes of one map
        gpu_ms = sum(v for k, v in stage_acc.items() if k not in ("gvd_host_voronoi",))
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32 bit-planes / f32,f64 geometry", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {gi.width}x{gi.height} cells @ {spec0.grid_resolution} m, "
                                   f"{n_pts} points per map",
                       "maps_in_flight": T, "device_gate": gate, "step": f"{T} independent maps per GPU, one per stream/host thread "
                                                    "(the Subdiv2D replay of each map runs on its own host core)",
                       "host_cores": ncpu, "e2e_source": "pinned host memory" if pinned else "pageable host memory (pinning failed)",
                       "l2": "inputs (16 B x points per map) larger than L2; every map is re-read from HBM",
                       "pipeline": info.get("pipeline"), "graph": info.get("graph")},
            "e2e": {"value": round(e2e_value, 1), "unit": UNIT, "ms_per_step": round(e2e_ms / args.steps, 3),
                    "h2d_bytes_per_step": int(n_pts * 16) * T, "d2h_bytes_per_step": int(info_h.get("d2h_bytes", 0)) * T},
            "single_map": {"ms_per_map": round(single_ms, 3), "value": round(cells / (single_ms * 1e-3) / 1e6, 1), "unit": UNIT,
                           "device_stages_ms": round(gpu_ms, 3), "host_voronoi_ms": round(stage_acc.get("gvd_host_voronoi", 0.0), 3)},
            "gpu_launches": int(round(launches)),
            "roofline": {"bound": "hbm", "kernel": "bin_points_xyz16", "achieved": round(achieved, 1) if achieved else None,
                         "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4) if achieved else None,
                         "traffic": 3280462840 if args.workload == "C3" and args.points is None else None,
                         "traffic_source": "ncu --set full, profiles/r01_n_pipeline_c3.txt (dram__bytes_read.sum 3.241483 GB + dram__bytes_write.sum 38.98 MB per launch)",
                         "peak_source": peak_src,
                         "whole_map": {"algorithmic_bytes": int(b_alg),
                                       "device_stages_gbs": round(b_alg / (gpu_ms * 1e-3) / 1e9, 1) if gpu_ms > 0 else None,
                                       "device_stages_frac": round(b_alg / (gpu_ms * 1e-3) / 1e9 / peak, 4) if gpu_ms > 0 else None}},
            "stages_ms": {k: round(v, 4) for k, v in stage_acc.items()},
            "clocks": clk.summary(),
            "gen_s": round(gen_s, 2),
        }
        if not args.no_cpu_baseline and world == 1:   # reported beside the N = 1 line only
            line["cpu_baseline"] = cpu_baseline(args, cores=1, budget_s=args.cpu_budget)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    pool.shutdown()
    for c in ctxs:
        c.close()
    return line


def bench_sweep(args):
    """BASELINE config 5: a sweep of 256 independent C2-size maps (row pitch in [4, 8] m, inflation 0.6 / 0.8 / 1.0 m),
    map i on rank i mod N, each rank keeping several maps in flight.  One step = the whole sweep."""
    import concurrent.futures as cf
    import queue

    import torch

    from aos_gpu import dist as adist
    from aos_gpu import lib, synth

    rank, world, local = adist.env_rank_world()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    n_maps = 256
    mine = adist.map_assignment(n_maps, world, rank)
    ncpu = os.cpu_count() or 1
    T = args.maps_in_flight if args.maps_in_flight > 0 else max(1, min(16, ncpu // max(world, 1), len(mine)))
    specs, prms, clouds, hosts = [], [], [], []
    for i in mine:
        r = np.random.default_rng(5000 + i)
        spec = synth.OrchardSpec(row_pitch=float(r.uniform(4.0, 8.0)), inflation_radius=float(r.choice([0.6, 0.8, 1.0])),
                                 n_points=args.points or 2_000_000, seed=i)
        specs.append(spec)
        prms.append(make_params(lib, spec))
        clouds.append(synth.make_orchard_torch(spec, dev))
        h = torch.empty(clouds[-1].shape, dtype=torch.float32, pin_memory=True)
        h.copy_(clouds[-1])
        hosts.append(h.numpy())
    torch.cuda.synchronize()
    gi = lib.grid_geometry(prms[0])
    cells = gi.width * gi.height
    ctxs = [lib.Context(local) for _ in range(T)]
    fetch = [{} for _ in range(T)]
    pool = cf.ThreadPoolExecutor(max_workers=T)

    def sweep(host):
        todo = queue.SimpleQueue()
        for k in range(len(mine)):
            todo.put(k)
        stats = []

        def work(t):
            torch.cuda.set_device(local)
            n_nodes = 0
            while True:
                try:
                    k = todo.get_nowait()
                except queue.Empty:
                    return n_nodes
                info = ctxs[t].map_to_graph(prms[k], hosts[k] if host else clouds[k], fetch=fetch[t] if host else None)
                n_nodes += (info["graph"] or {}).get("nodes", 0)
        stats = list(pool.map(work, range(T)))
        return sum(stats)

    def timed(steps, host):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        w0 = time.perf_counter()
        nodes = 0
        for _ in range(steps):
            nodes = sweep(host)
        torch.cuda.synchronize()
        ev1.record()
        ev1.synchronize()
        ms = max(ev0.elapsed_time(ev1), (time.perf_counter() - w0) * 1e3)
        return ms, nodes

    for _ in range(args.warmup):
        sweep(False)
    l0 = sum(c.launch_count() for c in ctxs)
    with ClockSampler(local) as clk:
        dev_ms, nodes = timed(args.steps, False)
    launches = (sum(c.launch_count() for c in ctxs) - l0) / max(args.steps, 1)
    sweep(True)
    e2e_ms, _ = timed(args.steps, True)
    dev_ms, total_cells = adist.reduce_stats(dev_ms, cells * len(mine) * args.steps, device=dev)
    e2e_ms, _ = adist.reduce_stats(e2e_ms, 0, device=dev)
    if rank == 0:
        n_pts = sum(int(c.shape[0]) for c in clouds)
        line = {"metric": METRIC, "value": round(adist.throughput_mcells(total_cells, dev_ms), 1), "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dev_ms / args.steps, 3), "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "u32 bit-planes / f32,f64 geometry", "data": "synthetic",
                "config": {"workload": f"C5: sweep of {n_maps} independent {gi.width}x{gi.height} maps @ 0.05 m (row pitch 4-8 m, "
                                       "inflation 0.6/0.8/1.0 m, 2 M points each), map i on rank i mod N",
                           "maps_in_flight": T, "host_cores": ncpu, "maps_per_s": round(n_maps * args.steps / (dev_ms * 1e-3), 1),
                           "graph_nodes_rank0_sweep": int(nodes)},
                "e2e": {"value": round(adist.throughput_mcells(total_cells, e2e_ms), 1), "unit": UNIT,
                        "ms_per_step": round(e2e_ms / args.steps, 3), "h2d_bytes_per_step": int(n_pts * 16),
                        "d2h_bytes_per_step": None},
                "gpu_launches": int(round(launches)), "clocks": clk.summary()}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    pool.shutdown()
    for c in ctxs:
        c.close()


def bench_bands(args):
    """BASELINE config 4: ONE grid row-band sharded over the ranks (strong scaling of the raster stages; clusters,
    seeds and graph finish on rank 0).  Not the default: `--shard bands`, normally with --workload C4."""
    import torch
    import torch.distributed as dist

    from aos_gpu import bands
    from aos_gpu import dist as adist
    from aos_gpu import lib, synth

    rank, world, local = adist.env_rank_world()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    spec = synth.config(args.workload, seed=0, n_points=args.points)
    params = make_params(lib, spec)
    gi = lib.grid_geometry(params)
    ctx = lib.Context(local)
    band = bands.band_for(gi.height, world, rank, ctx.band_halo_rows(params))
    res = float(np.float32(spec.grid_resolution))
    ylo = gi.origin_y + band.first_global_row * res - 2 * res
    yhi = gi.origin_y + (band.first_global_row + band.local_rows) * res + 2 * res
    pts = synth.make_orchard_torch(spec, dev, y_range=(ylo, yhi)) if world > 1 else synth.make_orchard_torch(spec, dev)
    torch.cuda.synchronize()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def one():
        t = {}
        t0 = time.perf_counter()
        ctx.band_raster(params, band, pts)
        torch.cuda.synchronize()
        t["raster"] = time.perf_counter() - t0
        be = bands.LibBackend(ctx)
        t0 = time.perf_counter()
        if args.halo == "p2p":
            launches = bands.run_thinning_p2p(be, band, rank, world, dist, device=dev)
        else:
            launches = bands.run_thinning(be, band, rank, world, dist)
        torch.cuda.synchronize()
        t["thin+halo"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        skel = bands.gather_rows(be.skeleton(), band, gi.height, rank, world, dist)
        occ = bands.gather_rows(be.grid(lib.GRID_OCCUPANCY), band, gi.height, rank, world, dist)
        torch.cuda.synchronize()
        t["gather"] = time.perf_counter() - t0
        g = None
        if rank == 0:
            t0 = time.perf_counter()
            ctx.seed_stage_tail(params, skel, occ)
            seeds, counts, rows_info = ctx.select_seeds()
            g = ctx.gvd_stage(seeds, rows_info) if len(seeds) else None
            t["tail_rank0"] = time.perf_counter() - t0
        return t, launches, g

    for _ in range(args.warmup):
        one()
    sync_all()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    acc = {}
    with ClockSampler(local) as clk:
        ev0.record()
        w0 = time.perf_counter()
        for _ in range(args.steps):
            t, launches, g = one()
            for k, v in t.items():
                acc[k] = acc.get(k, 0.0) + v * 1e3 / args.steps
        sync_all()
        ev1.record()
        ev1.synchronize()
    ms = max(ev0.elapsed_time(ev1), (time.perf_counter() - w0) * 1e3)
    ms, _ = adist.reduce_stats(ms, 0, device=dev)
    n_local = torch.tensor([pts.shape[0]], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(n_local)
    cells = gi.width * gi.height
    if rank == 0:
        raster_ms = acc.get("raster", 0) + acc.get("thin+halo", 0) + acc.get("gather", 0)
        line = {"metric": METRIC, "value": round(cells * args.steps / (ms * 1e-3) / 1e6, 1), "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "u32 bit-planes / f32,f64 geometry", "data": "synthetic",
                "config": {"workload": f"{args.workload}: ONE {gi.width}x{gi.height} grid @ {spec.grid_resolution} m row-band "
                                       f"sharded over {world} GPU(s), {int(n_local.item())} points in total (halo overlap "
                                       f"included), halo {ctx.band_halo_rows(params)} rows, " +
                                       ("8 edge rows per neighbour stored into its peer-mapped buffer by the thinning kernel "
                                        "(NVLink P2P), flags all-reduced" if args.halo == "p2p" else
                                        "NCCL send/recv of 8 rows per neighbour and thinning launch"),
                           "shard": "bands", "halo": args.halo, "thin_launches": launches,
                           "graph": None if g is None else {"nodes": int(g["n_nodes"]), "edges": int(g["n_edges"])}},
                "stages_ms_rank0": {k: round(v, 3) for k, v in acc.items()},
                "raster_stages": {"ms": round(raster_ms, 3), "value": round(cells / (raster_ms * 1e-3) / 1e6, 1), "unit": UNIT},
                "clocks": clk.summary()}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


# ---------------------------------------------------------------------------------------------------
# CPU arm: the oracle (port of the reference's algorithms) on a bounded crop of the workload
# ---------------------------------------------------------------------------------------------------
def crop_spec(workload: str, seed: int):
    """A 200 m x 120 m window of the workload's orchard (same row pitch, tree model, point density): the largest
    crop on which the reference's quadratic loops (O(E*M) node search, O(n^2) cluster diameter) still finish in
    seconds per map; smaller crops flatter the CPU arm, the full 1 km^2 map would take hours."""
    from aos_gpu import synth
    full = synth.config(workload, seed=seed)
    ex, ey = min(full.extent_x, 200.0), min(full.extent_y, 120.0)
    density = full.n_points / (full.extent_x * full.extent_y)
    s = synth.OrchardSpec(extent_x=ex, extent_y=ey, row_pitch=full.row_pitch, tree_spacing=full.tree_spacing,
                          tree_radius=full.tree_radius, n_points=int(density * ex * ey), seed=seed,
                          grid_resolution=full.grid_resolution, inflation_radius=full.inflation_radius)
    return s


def _cpu_one(args_tuple):
    workload, seed, reps = args_tuple
    from aos_gpu import synth
    from oracle import oracle as O
    spec = crop_spec(workload, seed)
    pts = synth.make_orchard(spec)
    p = O.SeedParams(grid_resolution=spec.grid_resolution, inflation_radius=spec.inflation_radius, polygon=spec.polygon)
    times = []
    cells = 0
    for _ in range(reps):
        t = time.perf_counter()
        r = O.seed_stage(p, pts)
        O.gvd_stage(r["seeds"], r["skel_framed"], r["origin_x"], r["origin_y"], r["res"], r["rows_info"])
        times.append(time.perf_counter() - t)
        cells = r["w"] * r["h"]
    return cells, times, len(pts)


def _sample_text(args, cells, npts, what):
    spec = crop_spec(args.workload, 1000)
    return (f"{spec.extent_x:g}x{spec.extent_y:g} m crop of {args.workload} ({cells} cells, {npts} points), oracle port "
            f"(C, -O2) seed stage + gvd stage incl. cv2.Subdiv2D, {what}")


def cpu_baseline(args, cores: int, budget_s: float):
    """The oracle (port of the reference's algorithms) on ONE host core: maps of the bounded crop, back to back,
    until `budget_s` seconds of CPU work are spent (at least one map)."""
    from oracle import oracle as O
    O.build()
    t0 = time.perf_counter()
    total, n, cells, npts = 0.0, 0, 0, 0
    while n == 0 or (total + total / n) < budget_s:
        cells, times, npts = _cpu_one((args.workload, 1000 + n, 1))
        total += times[0]
        n += 1
    value = cells * n / total / 1e6
    return {"value": round(value, 3), "unit": UNIT, "cores": 1, "kind": "port",
            "sample": _sample_text(args, cells, npts, f"{n} map(s) on one core"),
            "ms_per_map": round(1e3 * total / n, 2), "wall_s": round(time.perf_counter() - t0, 1)}


def bench_reference(args):
    """CPU arm: every host core runs one crop map per step (independent maps, like the GPU arm's maps in flight)."""
    import multiprocessing as mp
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    O.build()
    cores = max(1, min(os.cpu_count() or 1, 64))
    vals, cells, npts, t_start = [], 0, 0, time.perf_counter()
    with mp.get_context("fork").Pool(cores) as pool:
        for i in range(args.warmup + args.steps):
            t1 = time.perf_counter()
            res = pool.map(_cpu_one, [(args.workload, 1000 + i * cores + k, 1) for k in range(cores)])
            wall = time.perf_counter() - t1
            cells, npts = res[0][0], res[0][2]
            if i >= args.warmup:
                vals.append(cells * cores / wall / 1e6)
            if time.perf_counter() - t_start > 240 and vals:   # keep the whole arm within minutes
                break
    value = float(np.mean(vals))
    cb = {"value": round(value, 3), "unit": UNIT, "cores": cores, "kind": "port",
          "sample": _sample_text(args, cells, npts, f"one map per core and step, {cores} cores"),
          "ms_per_map": round(1e3 * cells * cores / (value * 1e6), 2), "wall_s": round(time.perf_counter() - t_start, 1)}
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
            "steps": len(vals), "warmup": args.warmup, "ms_per_step": round(1e3 * cells * cores / (value * 1e6), 2),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "i8 grids / f32,f64 geometry",
            "data": "synthetic", "config": {"workload": f"{args.workload} (bounded crop, see cpu_baseline.sample)",
                                            "maps_in_flight": cores},
            "cpu_baseline": cb,
            "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def crop_cells(args):
    from aos_gpu import lib
    spec = crop_spec(args.workload, 1000)
    gi = lib.grid_geometry(make_params(lib, spec))
    return gi.width * gi.height


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C3")
    ap.add_argument("--points", type=int, default=None, help="override the workload's point count")
    ap.add_argument("--maps-in-flight", type=int, default=0,
                    help="independent maps processed concurrently per GPU (0 = min(16, host cores / ranks))")
    ap.add_argument("--halo", default="p2p", choices=["p2p", "nccl"],
                    help="--shard bands: p2p = halo rows stored into peer memory by the thinning kernel; nccl = send/recv")
    ap.add_argument("--shard", default="maps", choices=["maps", "bands"],
                    help="maps: independent maps per GPU (default, weak scaling); bands: one grid row-band sharded")
    ap.add_argument("--cpu-budget", type=float, default=20.0, help="seconds of CPU work for cpu_baseline")
    ap.add_argument("--device-gate", type=int, default=-1,
                    help="maps admitted to the seed stage's kernel phase at a time per GPU (aos_set_device_gate); 0 = no limit; "
                         "default 2 when at least 4 maps are in flight")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        bench_reference(args)
    elif args.shard == "bands":
        bench_bands(args)
    elif args.workload.upper() == "C5":
        bench_sweep(args)
    else:
        bench_ours(args)


if __name__ == "__main__":
    main()
