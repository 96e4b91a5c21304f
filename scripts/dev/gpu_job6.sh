set -x
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --maps-in-flight 1"
$B > gpurun_out/plain_n.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^(?!.*at::).*' -c 1600 --csv --log-file gpurun_out/launches_n.csv $B > gpurun_out/ncu_launches_n.log 2>&1
$B > gpurun_out/plain_n2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'bin_points_xyz16|thin_kernel|vor_points_kernel|facet_fill_kernel|merge_round_kernel' -s 8 -c 8 -o gpurun_out/prof_n $B > gpurun_out/ncu_n.log 2>&1
tail -2 gpurun_out/ncu_n.log
ls -la gpurun_out/*.ncu-rep
