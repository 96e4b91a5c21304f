set -x
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --maps-in-flight 1 > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'bin_points_xyz16|thin_kernel' -s 2 -c 4 -o gpurun_out/prof_bin_thin python bench.py --steps 1 --warmup 1 --no-cpu-baseline --maps-in-flight 1 > gpurun_out/ncu3.log 2>&1
tail -2 gpurun_out/ncu3.log
