python bench.py --steps 1 --warmup 1 --no-cpu-baseline --maps-in-flight 1 > gpurun_out/plain4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'bin_points_xyz16' -s 1 -c 1 -o gpurun_out/prof_bin2 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --maps-in-flight 1 > gpurun_out/ncu4.log 2>&1
tail -2 gpurun_out/ncu4.log
