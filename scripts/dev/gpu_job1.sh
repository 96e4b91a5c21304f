set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -2 gpurun_out/bench_c3.err; cat gpurun_out/bench_c3.json
python bench.py --workload C2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; cat gpurun_out/bench_c2.json
KR='regex:bin_points|inflate_kernel|open_kernel|thin_kernel|frame_kernel|mask_count|scan_|cc_|acc_|cluster_finalize|bfs_replay|root_cellpos|labels_kernel|unpack|pack_kernel'
python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KR" -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:bin_points_xyz16 -s 1 -c 1 -o gpurun_out/prof_bin python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bin.log 2>&1
ls -la gpurun_out
