set -x
AOS_DEBUG=1 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --maps-in-flight 1 > gpurun_out/plain.log 2> gpurun_out/plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^(?!.*at::).*' -c 1200 --csv --log-file gpurun_out/launches2.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --maps-in-flight 1 > gpurun_out/ncu_launches2.log 2>&1
tail -3 gpurun_out/plain.err
tail -2 gpurun_out/ncu_launches2.log
