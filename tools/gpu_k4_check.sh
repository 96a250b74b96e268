set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_oscbank.py tests/test_recurrences.py tests/test_full_size.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/k4_tests.log
cat gpurun_out/k4_tests.log
timeout 300 python tools/k4_probe.py base > gpurun_out/k4_probe.log 2>&1
cat gpurun_out/k4_probe.log
