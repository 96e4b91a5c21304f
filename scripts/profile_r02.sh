#!/bin/bash
# ncu evidence for round 2, run on the GPU box (one GPU):  bash scripts/profile_r02.sh
# 1. the program exits 0 without ncu  2. launch list  3. --set full of the kernels VERDICT names.  Outputs under gpurun_out/r2p/.
set -u
O=gpurun_out/r2p
mkdir -p $O
B="python bench.py --steps 1 --warmup 1 --maps-in-flight 1 --no-bands --no-cpu-baseline --no-device-voronoi"
timeout 300 $B > $O/plain.json 2> $O/plain.err || { echo "plain run failed"; tail -5 $O/plain.err; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches.csv $B > $O/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"bfs_replay_kernel|cc_link_kernel|edge_test_kernel|vs_generate_kernel|ray_points_kernel|thin_kernel|bin_points|inflate_kernel|corner_kernel|mask_count_kernel" \
  -s 16 -c 16 -o $O/prof_r02 -f $B > $O/ncu_full.log 2>&1
# the opt-in device Voronoi kernel
cat > $O/vc.py <<'PY'
import sys, os
sys.path[:0] = [os.getcwd(), os.path.join(os.getcwd(), "active-orchard-slam_b200")]
import torch
from aos_gpu import lib, synth
spec = synth.config("C3", seed=0)
pts = synth.make_orchard_torch(spec, "cuda")
p = lib.SeedParams(grid_resolution=spec.grid_resolution, inflation_radius=spec.inflation_radius, polygon=spec.polygon)
ctx = lib.Context(0)
ctx.set_voronoi_mode(True)
for _ in range(2):
    ctx.map_to_graph(p, pts)
print("ok", ctx.graph()["n_nodes"])
PY
timeout 300 python $O/vc.py > $O/vc_plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:"vc_cell_kernel" -c 2 -o $O/prof_vc -f python $O/vc.py > $O/ncu_vc.log 2>&1
python scripts/ncu_summary.py $O/launches.csv $O/prof_r02.ncu-rep $O/prof_vc.ncu-rep > $O/summary.txt 2>&1
# gpurun copies back at most 64 MiB: keep the text exports, drop the reports
for r in prof_r02 prof_vc; do
  [ -f $O/$r.ncu-rep ] && ncu -i $O/$r.ncu-rep --page raw --csv 2>/dev/null | gzip > $O/$r.raw.csv.gz
  rm -f $O/$r.ncu-rep
done
rm -f gpurun_out/*.ncu-rep
du -sh gpurun_out
tail -5 $O/summary.txt
