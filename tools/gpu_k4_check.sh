# GPU-box helper: the K4 / cfg3 checks used while tuning scan.cu (tests, then the delay-distribution probe)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_recurrences.py tests/test_oscbank.py tests/test_full_size.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/k4_tests.log
cat gpurun_out/k4_tests.log
timeout 300 python tools/k4_probe.py base base+ring ge256 > gpurun_out/k4_probe.log 2>&1
cat gpurun_out/k4_probe.log
