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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ncpu = os.cpu_count() or 1
    # One host thread per map in flight, as many maps as this rank has cores: a map keeps a core busy during its Subdiv2D
    # replay (0.2 s at C3).  More maps than cores (--maps-in-flight) with sleeping instead of spinning waits
    # (--host-wait blocking) fill the core a map leaves idle while it queues for the GPU, but measured on one box back to
    # back it is a wash: 16 maps spinning 28.9 / 29.7 k, 18 spinning 25.7 k, 18 blocking 29.2 / 28.9 k Mcells/s.
    cores = max(1, min(16, ncpu // max(world, 1)))
    T = args.maps_in_flight if args.maps_in_flight > 0 else cores
    if args.maps_in_flight <= 0:   # every map in flight holds its cloud twice (resident copy + e2e staging) plus grids
        spec_probe = synth.config(args.workload, n_points=args.points)
        per_map = 2 * 16 * spec_probe.n_points + 1.5e9
        free_b, _ = torch.cuda.mem_get_info(local)
        T = max(1, min(T, int(0.85 * free_b / per_map)))
    gate = (2 if T >= 4 else 0) if args.device_gate < 0 else args.device_gate
    if not args.no_cpu_baseline:   # the parity leg compares with the oracle, whose Subdiv2D is this image's cv2
        import ctypes
        from oracle import subdiv as _sd
        lib.load().aos_set_subdiv_outer_factor(ctypes.c_float(_sd.outer_factor()))
    lib.load().aos_set_device_gate(gate)   # maps admitted to the seed stage's kernel phase at a time (0 = no limit)
    host_wait = "spin"
    want_blocking = args.host_wait == "blocking" or (args.host_wait == "auto" and T > cores)
    if want_blocking and lib.load().aos_set_host_wait(local, 1) == 0:
        host_wait = "blocking during value/e2e (threads waiting for the GPU sleep), spinning for single_map"
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

    if host_wait != "spin":   # latency mode for what follows: a waiting thread spins (a wake-up costs ~50 us per wait)
        lib.load().aos_set_host_wait(local, 0)
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

    # digest of map 0's result (every published array) -- compared with the CPU oracle on the same cloud below
    gpu_digest, gpu_parts = c0.result_digest(parts=True)
    device_voronoi = None
    if not args.no_device_voronoi:
        # the opt-in parallel Voronoi (aos_set_voronoi_mode(AOS_VORONOI_DEVICE)): same maps, no host insertion replay; NOT
        # bit-identical to the reference, so it is reported beside the headline, never as it
        from aos_gpu.compare import compare_graphs
        g_replay = c0.graph()
        for c in ctxs:
            c.set_voronoi_mode(True)
        run_batch(2)
        dv_ms, _ = timed(args.steps)
        with torch.cuda.stream(stream):
            c0.set_profiling(True)
            c0.map_to_graph(params, maps[0])
            dv_stages = dict(c0.stage_times())
            c0.set_profiling(False)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(3):
                c0.map_to_graph(params, maps[0])
            e1.record(stream)
            torch.cuda.synchronize()
            dv_single = e0.elapsed_time(e1) / 3
        g_dev = c0.graph()
        for c in ctxs:
            c.set_voronoi_mode(False)
        dv_ms, dv_cells = adist.reduce_stats(dv_ms, cells * args.steps * T, device=dev)
        device_voronoi = {"value": round(adist.throughput_mcells(dv_cells, dv_ms), 1), "unit": UNIT,
                          "ms_per_step": round(dv_ms / args.steps, 3), "single_map_ms": round(dv_single, 3),
                          "device_stages_ms": round(sum(dv_stages.values()), 3),
                          "voronoi_stage_ms": round(dv_stages.get("gvd_device_voronoi", 0.0), 3),
                          "bit_exact": False, "agreement_with_replay_map0": compare_graphs(g_replay, g_dev),
                          "note": "opt-in aos_set_voronoi_mode(AOS_VORONOI_DEVICE); the headline value/e2e use the bit-exact replay"}
    dev_ms, total_cells = adist.reduce_stats(dev_ms, cells * args.steps * T, device=dev)   # MAX over ranks, SUM of cells
    e2e_ms, _ = adist.reduce_stats(e2e_ms, cells * args.steps * T, device=dev)
    ms_per_step = dev_ms / args.steps
    value = adist.throughput_mcells(total_cells, dev_ms)
    e2e_value = adist.throughput_mcells(total_cells, e2e_ms)

    line = None
    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        b_alg = 16.0 * n_pts + 9.125 * cells  # compulsory bytes of one map
        host_stages = ("gvd_host_voronoi",)
        gpu_ms = sum(v for k, v in stage_acc.items() if k not in host_stages)
        plane = cells / 8.0
        # algorithmic bytes per stage (each plane read / written once at 1 bit per cell)
        stage_bytes = {"bin": 16.0 * n_pts, "inflate": 3 * plane, "open": 2 * plane, "thin": 2 * plane, "frame": 2 * plane,
                       "cc_mask_scan": plane}
        kernels = []
        for k, ms in sorted(stage_acc.items(), key=lambda kv: -kv[1]):
            if k in host_stages or ms <= 0:
                continue
            e = {"stage": k, "ms": round(ms, 4), "share_of_device_time": round(ms / gpu_ms, 4) if gpu_ms > 0 else None}
            if k in stage_bytes:
                gbs = stage_bytes[k] / (ms * 1e-3) / 1e9
                e.update(algorithmic_bytes=int(stage_bytes[k]), gbs=round(gbs, 1), frac=round(gbs / peak, 4))
            kernels.append(e)
        whole_gbs = b_alg / (gpu_ms * 1e-3) / 1e9 if gpu_ms > 0 else None
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32 bit-planes / f32,f64 geometry", "data": "synthetic",
            "config": workload_config(args, gi, spec0, n_pts, T=T, gate=gate, ncpu=ncpu, pinned=pinned, host_wait=host_wait,
                                      pipeline=info.get("pipeline"), graph=info.get("graph")),
            "e2e": {"value": round(e2e_value, 1), "unit": UNIT, "ms_per_step": round(e2e_ms / args.steps, 3),
                    "h2d_bytes_per_step": int(n_pts * 16) * T, "d2h_bytes_per_step": int(info_h.get("d2h_bytes", 0)) * T},
            "single_map": {"ms_per_map": round(single_ms, 3), "value": round(cells / (single_ms * 1e-3) / 1e6, 1), "unit": UNIT,
                           "device_stages_ms": round(gpu_ms, 3), "host_voronoi_ms": round(stage_acc.get("gvd_host_voronoi", 0.0), 3)},
            "gpu_launches": int(round(launches)),
            # the map-level figure: SURVEY 8(d)'s algorithmic bytes of ONE map / the device time of ONE map (all stages,
            # CUDA events on the library's stream, one map in flight); per-stage figures beside it
            "roofline": {"bound": "hbm", "kernel": "whole map: all device stages of aos_map_to_graph",
                         "achieved": round(whole_gbs, 1) if whole_gbs else None, "peak": peak, "unit": "GB/s",
                         "frac": round(whole_gbs / peak, 4) if whole_gbs else None,
                         "algorithmic_bytes": int(b_alg), "device_ms_per_map": round(gpu_ms, 4),
                         **map_traffic(args.workload),
                         "peak_source": peak_src,
                         "time_dominant_stage": kernels[0] if kernels else None,
                         "largest_traffic_kernel": next((k for k in kernels if k["stage"] == "bin"), None),
                         "stages": kernels},
            "stages_ms": {k: round(v, 4) for k, v in stage_acc.items()},
            "clocks": clk.summary(),
            "gen_s": round(gen_s, 2), "points_per_map": int(n_pts),
        }
        if device_voronoi is not None:
            line["device_voronoi"] = device_voronoi
        if not args.no_cpu_baseline and world == 1:   # reported beside the N = 1 line only
            cb, parity = cpu_baseline_full(args, host_np, spec0, gpu_digest, gpu_parts)
            line["cpu_baseline"] = cb
            line["parity"] = parity
    pool.shutdown()
    for c in ctxs:
        c.close()
    del maps
    torch.cuda.empty_cache()
    return line


def workload_config(args, gi, spec, n_pts, **kw):
    """The `config` object; the reference arm prints the same workload string (same grid, same points per map)."""
    cfg = {"workload": f"{args.workload}: {gi.width}x{gi.height} cells @ {spec.grid_resolution} m, {spec.n_points} points per map "
                       "(nominal; seeded synthetic orchard, SURVEY.md 8(d))"}
    if "T" in kw:
        T = kw["T"]
        cfg.update({"maps_in_flight": T, "device_gate": kw["gate"], "host_wait": kw.get("host_wait", "spin"),
                    "step": f"{T} independent maps per GPU, one per stream/host thread (the Subdiv2D replay of a map occupies a "
                            "host core; maps queueing for the GPU leave theirs to the others)",
                    "host_cores": kw["ncpu"],
                    "e2e_source": "pinned host memory" if kw["pinned"] else "pageable host memory (pinning failed)",
                    "l2": "inputs (16 B x points per map) larger than L2; every map is re-read from HBM",
                    "pipeline": kw["pipeline"], "graph": kw["graph"]})
    return cfg


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
    pool.shutdown()
    for c in ctxs:
        c.close()
    return line if rank == 0 else None


def run_bands(args, workload, steps, warmup, halo):
    """BASELINE config 4: ONE grid row-band sharded over the ranks (strong scaling of the raster stages; clusters, seeds and
    graph finish on rank 0).  Every N works on the SAME global cloud (synth.make_orchard_strips_torch: fixed strips with
    their own generator seeds; a rank generates the strips that touch its rows), so the result digest must equal the
    single-GPU digest of that cloud, which rank 0 computes with aos_map_to_graph and compares.  Collective: call on every
    rank with the process group up.  Returns the record on rank 0."""
    import torch
    import torch.distributed as dist

    from aos_gpu import bands
    from aos_gpu import dist as adist
    from aos_gpu import lib, synth

    rank, world, local = adist.env_rank_world()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    spec = synth.config(workload, seed=0, n_points=args.points)
    params = make_params(lib, spec)
    gi = lib.grid_geometry(params)
    ctx = lib.Context(local)
    band = bands.band_for(gi.height, world, rank, ctx.band_halo_rows(params))
    res = float(np.float32(spec.grid_resolution))
    ylo = gi.origin_y + band.first_global_row * res - 2 * res
    yhi = gi.origin_y + (band.first_global_row + band.local_rows) * res + 2 * res
    single = None
    if rank == 0:   # the single-GPU answer on the whole cloud
        full = synth.make_orchard_strips_torch(spec, dev)
        ref = lib.Context(local)
        ref.map_to_graph(params, full)
        single = {"digest": ref.result_digest(), "points": int(full.shape[0])}
        ref.close()
        if world == 1:
            pts = full
        del full
    if world > 1:
        torch.cuda.empty_cache()
        pts = synth.make_orchard_strips_torch(spec, dev, y_range=(ylo, yhi))
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
        if halo == "p2p":
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
            t["tail_rank0_host_voronoi"] = dict(ctx.stage_times()).get("gvd_host_voronoi", 0.0) * 1e-3
        return t, launches, g

    if rank == 0:
        ctx.set_profiling(True)
    for _ in range(warmup):
        one()
    sync_all()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    acc = {}
    with ClockSampler(local) as clk:
        ev0.record()
        w0 = time.perf_counter()
        for _ in range(steps):
            t, launches, g = one()
            for k, v in t.items():
                acc[k] = acc.get(k, 0.0) + v * 1e3 / steps
        sync_all()
        ev1.record()
        ev1.synchronize()
    ms = max(ev0.elapsed_time(ev1), (time.perf_counter() - w0) * 1e3)
    ms, _ = adist.reduce_stats(ms, 0, device=dev)
    n_local = torch.tensor([pts.shape[0]], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(n_local)
    cells = gi.width * gi.height
    digest = ctx.result_digest() if rank == 0 else None
    # the same band run with the opt-in device Voronoi on rank 0 (no host insertion replay in the tail): timing only
    dv = None
    if not getattr(args, "no_device_voronoi", False):
        if rank == 0:
            ctx.set_voronoi_mode(True)
        one()
        sync_all()
        acc2 = {}
        w1 = time.perf_counter()
        for _ in range(steps):
            t2, _, g2 = one()
            for k, v in t2.items():
                acc2[k] = acc2.get(k, 0.0) + v * 1e3 / steps
        sync_all()
        ms2, _ = adist.reduce_stats((time.perf_counter() - w1) * 1e3, 0, device=dev)
        if rank == 0:
            ctx.set_voronoi_mode(False)
            dv = {"ms_per_map": round(ms2 / steps, 3), "value": round(cells * steps / (ms2 * 1e-3) / 1e6, 1), "unit": UNIT,
                  "stages_ms_rank0": {k: round(v, 3) for k, v in acc2.items() if k != "tail_rank0_host_voronoi"},
                  "graph": None if g2 is None else {"nodes": int(g2["n_nodes"]), "edges": int(g2["n_edges"])}, "bit_exact": False}
    rec = None
    if rank == 0:
        raster_ms = acc.get("raster", 0) + acc.get("thin+halo", 0) + acc.get("gather", 0)
        rec = {"workload": f"{workload}: ONE {gi.width}x{gi.height} grid @ {spec.grid_resolution} m row-band sharded over "
                           f"{world} GPU(s); one global cloud of {single['points']} points for every N "
                           f"({int(n_local.item())} generated incl. halo overlap)",
               "shard": "bands", "halo": halo, "halo_rows": ctx.band_halo_rows(params), "thin_launches": launches,
               "steps": steps, "warmup": warmup, "scaling": "strong",
               "ms_per_map": round(ms / steps, 3), "value": round(cells * steps / (ms * 1e-3) / 1e6, 1), "unit": UNIT,
               "stages_ms_rank0": {k: round(v, 3) for k, v in acc.items()},
               "raster_stages": {"ms": round(raster_ms, 3), "value": round(cells / (raster_ms * 1e-3) / 1e6, 1), "unit": UNIT},
               "graph": None if g is None else {"nodes": int(g["n_nodes"]), "edges": int(g["n_edges"])},
               "digest": digest, "single_gpu_digest": single["digest"], "equals_single_gpu": digest == single["digest"],
               "device_voronoi": dv, "clocks": clk.summary()}
    ctx.close()
    del pts
    torch.cuda.empty_cache()
    return rec


def bench_bands(args):
    """`--shard bands`: the band run as the whole bench line (normally with --workload C4)."""
    rec = run_bands(args, args.workload, args.steps, args.warmup, args.halo)
    if rec is None:
        return None
    return {"metric": METRIC, "value": rec["value"], "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": rec["ms_per_map"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u32 bit-planes / f32,f64 geometry", "data": "synthetic",
            "config": {"workload": rec["workload"], "shard": "bands", "halo": rec["halo"], "thin_launches": rec["thin_launches"],
                       "graph": rec["graph"]},
            "stages_ms_rank0": rec["stages_ms_rank0"], "raster_stages": rec["raster_stages"],
            "digest": rec["digest"], "single_gpu_digest": rec["single_gpu_digest"], "equals_single_gpu": rec["equals_single_gpu"],
            "clocks": rec["clocks"]}


# ---------------------------------------------------------------------------------------------------
# CPU arms: the oracle (port of the reference's algorithms, pinned to the compiled reference by tests/test_ref_cpu.py) on
# the SAME configuration as the GPU arm -- full-size maps.  The reference's literal loops are quadratic in the number of
# graph nodes / cluster cells (O(E*M) node search, O(M^2) proximity pairs, O(n^2) cluster diameter: 1e11..1e12 steps at
# config 3), so the full-size runs use the oracle's indexed loops (oracle/aos_oracle_fast.c: hash grids, convex hulls,
# rows in parallel threads), which tests/test_oracle_fast_cpu.py proves bit-identical to the literal ones.
# ---------------------------------------------------------------------------------------------------
def oracle_map(O, spec, pts):
    """One map through the oracle; returns (seconds seed stage, seconds gvd stage, seed dict, graph dict)."""
    p = O.SeedParams(grid_resolution=spec.grid_resolution, inflation_radius=spec.inflation_radius, polygon=spec.polygon,
                     exclusion=spec.exclusion)
    t0 = time.perf_counter()
    r = O.seed_stage(p, pts)
    t1 = time.perf_counter()
    g = O.gvd_stage(r["seeds"], r["skel_framed"], r["origin_x"], r["origin_y"], r["res"], r["rows_info"])
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1, r, g


def reference_small_configs(budget_s=12.0):
    """The reference's OWN compiled code (oracle/_ref, literal loops, one thread) timed end to end on BASELINE configs 1
    and 2 -- the sizes it was written for.  Empty when the prebuilt library did not travel."""
    out = {}
    try:
        from aos_gpu import synth
        from oracle import oracle as O
        from oracle import ref as R
        if not R.available():
            return {"unavailable": "oracle/_ref/libaos_ref.so not present"}
        for name in ("C1", "C2"):
            spec = synth.config(name, seed=0)
            pts = synth.make_orchard(spec)
            p = O.SeedParams(grid_resolution=spec.grid_resolution, inflation_radius=spec.inflation_radius, polygon=spec.polygon)
            t0 = time.perf_counter()
            n = 0
            while n == 0 or (time.perf_counter() - t0) < budget_s / 2:
                a = R.seed_stage(p, pts)
                R.gvd_stage(a["seeds"], a["skel_framed"], a["origin_x"], a["origin_y"], a["res"], a["rows_info"])
                n += 1
            dt = (time.perf_counter() - t0) / n
            out[name] = {"cells": a["w"] * a["h"], "points": int(len(pts)), "ms_per_map": round(dt * 1e3, 1),
                         "value": round(a["w"] * a["h"] / dt / 1e6, 2), "unit": UNIT, "cores": 1, "kind": "reference",
                         "note": "exclusion discs = the node's 11 hard-coded ones"}
    except Exception as e:   # the CPU arm must never take the bench line down
        out["error"] = repr(e)[:200]
    return out


def cpu_baseline_full(args, host_np, spec, gpu_digest, gpu_parts):
    """cpu_baseline of the N = 1 line: ONE full map of the bench workload (map 0's own cloud, as the GPU processed it)
    through the oracle with every host core, timed; its result digest is then compared with the GPU's (the parity check
    of this very run, outside the timed region)."""
    from aos_gpu import lib
    from oracle import oracle as O
    O.build()
    cores = os.cpu_count() or 1
    t_start = time.perf_counter()
    O.set_fast(True, threads=cores, skip_labels=True)
    try:
        ts, tg, r, g = oracle_map(O, spec, host_np)
    finally:
        O.set_fast(False)
    cells = r["w"] * r["h"]
    total, per = lib.digest_of(O.result_artefacts(r, g), parts=True)
    bad = sorted(k for k in per if per[k] != gpu_parts.get(k))
    cb = {"value": round(cells / (ts + tg) / 1e6, 2), "unit": UNIT, "cores": cores, "kind": "port",
          "sample": f"1 full {args.workload} map ({cells} cells, {len(host_np)} points: map 0 of this run), oracle port with its "
                    f"indexed loops on {cores} threads (seed stage {ts:.1f} s, gvd stage incl. cv2.Subdiv2D {tg:.1f} s); the "
                    "reference's literal loops do not finish this size (O(E*M), O(M^2), O(n^2) per cluster)",
          "ms_per_map": round(1e3 * (ts + tg), 1), "wall_s": None,
          "reference_small_configs": reference_small_configs()}
    cb["wall_s"] = round(time.perf_counter() - t_start, 1)
    parity = {"checked": True, "against": "oracle (CPU) on the same cloud, sha256 per published array",
              "equal": total == gpu_digest, "arrays": len(per), "differing": bad, "digest": gpu_digest}
    return cb, parity


def bench_reference(args):
    """CPU arm on the GPU arm's configuration: every step is ONE full map of the workload through the oracle with all host
    cores (kind "port": the reference's own loops are quadratic and do not finish this size; the port's results are
    bit-identical to them wherever both run).  A fresh seeded cloud per step, generated outside the timed part."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from aos_gpu import synth
    from oracle import oracle as O
    O.build()
    cores = os.cpu_count() or 1
    spec0 = synth.config(args.workload, seed=0, n_points=args.points)
    t_start = time.perf_counter()
    O.set_fast(True, threads=cores, skip_labels=True)
    times, cells, npts = [], 0, 0
    budget_s = args.reference_budget
    done = 0
    try:
        for i in range(args.warmup + args.steps):
            spec = synth.config(args.workload, seed=i % 4, n_points=args.points)
            pts = synth.make_orchard_strips(spec)
            ts, tg, r, g = oracle_map(O, spec, pts)
            cells, npts = r["w"] * r["h"], len(pts)
            del r, g
            if i >= args.warmup:
                times.append(ts + tg)
            done = i + 1
            # the whole arm has to end within minutes: stop early once the budget is spent (steps reports what ran)
            if time.perf_counter() - t_start > budget_s and times:
                break
    finally:
        O.set_fast(False)
    per_map = float(np.mean(times))
    value = cells / per_map / 1e6
    from types import SimpleNamespace
    gi = SimpleNamespace(width=int(round(spec0.extent_x / spec0.grid_resolution)), height=int(round(spec0.extent_y / spec0.grid_resolution)))
    try:
        from aos_gpu import lib
        g2 = lib.grid_geometry(make_params(lib, spec0))
        gi = SimpleNamespace(width=g2.width, height=g2.height)
    except Exception:
        pass
    n_nominal = npts
    cb = {"value": round(value, 3), "unit": UNIT, "cores": cores, "kind": "port",
          "sample": f"{len(times)} full {args.workload} map(s) ({cells} cells, {npts} points each), one per step, oracle port with "
                    f"its indexed loops on {cores} threads; {done - len(times)} warm-up map(s)",
          "ms_per_map": round(1e3 * per_map, 1), "wall_s": round(time.perf_counter() - t_start, 1)}
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": UNIT,
            "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": len(times), "warmup": min(args.warmup, done - len(times)),
            "ms_per_step": round(1e3 * per_map, 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "i8 grids / f32,f64 geometry", "data": "synthetic",
            "config": workload_config(args, gi, spec0, n_nominal),
            "cpu_baseline": cb, "points_per_map": int(npts),
            "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def map_traffic(workload):
    """roofline.traffic = dram__bytes_read + dram__bytes_write of ALL kernels of one map, from the committed ncu launch
    list of the same pipeline (scripts/profile_r02.sh -> scripts/ncu_traffic.py); not re-measured per run (ncu cannot run
    inside a timed bench), so the source file and its date travel with the number."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", f"r02_traffic_{workload.lower()}.json")
    try:
        t = json.load(open(path))
        return {"traffic": int(t["dram_bytes"]),
                "traffic_note": f"sum over the {t['launches']} launches of one map, ncu dram__bytes_read+write, captured "
                                f"{t['captured']} (profiles/{os.path.basename(path)}; cold-cache upper bound)"}
    except (OSError, KeyError, ValueError):
        return {"traffic": None, "traffic_note": "no committed ncu traffic capture for this workload under profiles/"}


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
                    help="band runs: p2p = halo rows stored into peer memory by the thinning kernel; nccl = send/recv")
    ap.add_argument("--shard", default="maps", choices=["maps", "bands"],
                    help="maps: independent maps per GPU (default, weak scaling) followed by the config-4 band run as the "
                         "`bands` sub-record; bands: only the band run, as the bench line")
    ap.add_argument("--no-bands", action="store_true", help="skip the config-4 band sub-record of the default run")
    ap.add_argument("--no-sweep", action="store_true", help="skip the config-5 sweep sub-record of the default run")
    ap.add_argument("--bands-workload", default="C4")
    ap.add_argument("--cpu-budget", type=float, default=20.0, help="(kept for compatibility; the CPU leg runs one full map)")
    ap.add_argument("--reference-budget", type=float, default=420.0,
                    help="--impl reference: seconds after which no further step is started")
    ap.add_argument("--device-gate", type=int, default=-1,
                    help="maps admitted to the seed stage's kernel phase at a time per GPU (aos_set_device_gate); 0 = no limit; "
                         "default 2 when at least 4 maps are in flight")
    ap.add_argument("--host-wait", default="auto", choices=["auto", "spin", "blocking"],
                    help="how threads wait for the GPU in the throughput legs (aos_set_host_wait); auto = blocking when more "
                         "maps are in flight than this rank has cores")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-device-voronoi", action="store_true", help="skip the opt-in device-Voronoi sub-record")
    args = ap.parse_args()
    if args.impl == "reference":
        bench_reference(args)
        return
    import torch
    from aos_gpu import dist as adist
    rank, world, local = adist.env_rank_world()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; libaos_gpu has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if args.shard == "bands":
        line = bench_bands(args)
    elif args.workload.upper() == "C5":
        line = bench_sweep(args)
    else:
        line = bench_ours(args)
        if not args.no_bands and args.workload.upper() == "C3" and args.points is None:
            # BASELINE config 4 (the row-band sharded grid) measured in the same driver-visible run: one 40000^2 grid over
            # the N ranks, digest-checked against the single-GPU result of the same cloud
            rec = run_bands(args, args.bands_workload, max(1, min(args.steps, 5)), 2, args.halo)
            if line is not None:
                line["bands"] = rec
        if not args.no_sweep and args.workload.upper() == "C3" and args.points is None:
            # BASELINE config 5 (256 independent small maps, map i on rank i mod N) in the same driver-visible run
            import copy
            a5 = copy.copy(args)
            a5.steps, a5.warmup, a5.maps_in_flight = max(1, min(args.steps, 3)), 2, 0   # contexts reach their buffer sizes in two sweeps
            rec5 = bench_sweep(a5)
            if line is not None and rec5 is not None:
                line["sweep"] = {"workload": rec5["config"]["workload"], "maps_in_flight": rec5["config"]["maps_in_flight"],
                                 "steps": rec5["steps"], "warmup": rec5["warmup"], "scaling": rec5["scaling"],
                                 "ms_per_sweep": rec5["ms_per_step"], "maps_per_s": rec5["config"]["maps_per_s"],
                                 "value": rec5["value"], "unit": rec5["unit"], "e2e": rec5["e2e"],
                                 "gpu_launches_per_sweep": rec5["gpu_launches"], "clocks": rec5["clocks"]}
    if rank == 0 and line is not None:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
