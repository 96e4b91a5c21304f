set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02_m_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02_m_pytest_gpu.log
python bench.py > gpurun_out/r02_m_bench_c3_n1.json 2> gpurun_out/r02_m_bench_c3_n1.err; tail -3 gpurun_out/r02_m_bench_c3_n1.err; cut -c1-1500 gpurun_out/r02_m_bench_c3_n1.json
