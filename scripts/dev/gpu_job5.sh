set -x
timeout 200 python bench.py --shard bands --workload C4 --steps 2 --warmup 1 2>&1 | grep "^{" > gpurun_out/bench_bands_c4_n1.json
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --maps-in-flight 1 > gpurun_out/plain5.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^(?!.*at::).*' -c 1400 --csv --log-file gpurun_out/launches5.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --maps-in-flight 1 > gpurun_out/ncu_launches5.log 2>&1
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --maps-in-flight 1 > gpurun_out/plain6.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'bin_points_xyz16|inflate_kernel|open_kernel' -s 3 -c 3 -o gpurun_out/prof_final python bench.py --steps 1 --warmup 1 --no-cpu-baseline --maps-in-flight 1 > gpurun_out/ncu6.log 2>&1
tail -2 gpurun_out/ncu6.log
