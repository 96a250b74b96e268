set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_jit.py tests/test_random_dags.py tests/test_edge_cases.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/k4_tests.log
cat gpurun_out/k4_tests.log
timeout 300 python tools/bench_kernels.py pure elementwise 2>&1 | cut -c1-330 > gpurun_out/k23.log
cat gpurun_out/k23.log
