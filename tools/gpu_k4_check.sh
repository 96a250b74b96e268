set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_edge_cases.py tests/test_jit.py tests/test_golden_gpu.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/k4_tests.log
cat gpurun_out/k4_tests.log
timeout 300 python tools/bench_kernels.py pure elementwise 2>&1 | cut -c1-330 > gpurun_out/k23.log
cat gpurun_out/k23.log
timeout 600 ncu --set full --clock-control none -k regex:frb_stage -s 4 -c 1 -f -o gpurun_out/prof_stage_r1o python tools/bench_kernels.py pure > gpurun_out/ncu_stage.log 2>&1
tail -2 gpurun_out/ncu_stage.log
