#!/bin/bash
# ncu evidence for round 2, run on the GPU box (one GPU):  bash scripts/profile_r02.sh
# 1. the program exits 0 without ncu  2. launch list of the same command (time + DRAM bytes per launch)
# 3. --set full of the kernels VERDICT names and of the ones that lead the device time.  Outputs under gpurun_out/r2p/.
set -u
O=gpurun_out/r2p
mkdir -p $O
B="python bench.py --steps 1 --warmup 1 --maps-in-flight 1 --no-bands --no-cpu-baseline --no-device-voronoi"
timeout 300 $B > $O/plain.json 2> $O/plain.err || { echo "plain run failed"; tail -5 $O/plain.err; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches.csv $B > $O/ncu_launches.log 2>&1
# one map at a time through the default (bit-exact replay) path: time and DRAM bytes of every launch
M="python scripts/dev/time_map.py C3 replay"
timeout 300 $M > $O/map_plain.log 2>&1 || { echo "time_map failed"; tail -5 $O/map_plain.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2500 --csv \
  --log-file $O/launches3.csv $M > $O/ncu_launches3.log 2>&1
python scripts/ncu_traffic.py $O/launches3.csv $O/traffic_c3.json
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"bfs_chain_kernel|bfs_jump_kernel|bfs_prepare_kernel|cc_link_kernel|edge_test_kernel|vs_generate_kernel|ray_points_kernel|thin_kernel|bin_points|inflate_kernel|corner_kernel|mask_count_kernel|boundary_round_kernel|cluster_finalize_kernel" \
  -s 60 -c 60 -o $O/prof_r02 -f $M > $O/ncu_full.log 2>&1
# the opt-in device Voronoi kernel
D="python scripts/dev/time_map.py C3 device"
timeout 300 $D > $O/vc_plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:"vc_cell_kernel" -c 2 -o $O/prof_vc -f $D > $O/ncu_vc.log 2>&1
python scripts/ncu_summary.py $O/launches.csv $O/prof_r02.ncu-rep $O/prof_vc.ncu-rep > $O/summary.txt 2>&1
# gpurun copies back at most 64 MiB: keep the text exports, drop the reports
for r in prof_r02 prof_vc; do
  [ -f $O/$r.ncu-rep ] && ncu -i $O/$r.ncu-rep --page raw --csv 2>/dev/null | gzip > $O/$r.raw.csv.gz
  rm -f $O/$r.ncu-rep
done
rm -f gpurun_out/*.ncu-rep
du -sh gpurun_out
tail -5 $O/summary.txt
