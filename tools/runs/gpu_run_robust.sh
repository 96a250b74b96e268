mkdir -p gpurun_out
(FRB_BLOCK_SAMPLES=4096 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3) > gpurun_out/r2s_pytest_block4096.log 2>&1; cat gpurun_out/r2s_pytest_block4096.log
(FRB_OSC_MIN_RANGES=8 FRB_BLOCK_SAMPLES=1048576 timeout 900 python -m pytest tests/test_oscbank.py tests/test_full_size.py tests/test_stream.py tests/test_edge_cases.py -x -q -m gpu 2>&1 | tail -3) > gpurun_out/r2s_pytest_ranges8.log 2>&1; cat gpurun_out/r2s_pytest_ranges8.log
