set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02_n_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02_n_pytest_gpu.log
python bench.py --impl reference > gpurun_out/r02_n_bench_reference_arm.json 2> gpurun_out/r02_n_bench_reference_arm.err; cut -c1-300 gpurun_out/r02_n_bench_reference_arm.json
python bench.py > gpurun_out/r02_n_bench_c3_n1.json 2> gpurun_out/r02_n_bench_c3_n1.err; tail -3 gpurun_out/r02_n_bench_c3_n1.err; cut -c1-1200 gpurun_out/r02_n_bench_c3_n1.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
