mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/r2v_pytest.log 2>&1
tail -5 gpurun_out/r2v_pytest.log
(FRB_OSC_GEMM=2 timeout 900 python -m pytest tests -m gpu -q -k "osc or full_size or bank or multi or shard or stream") > gpurun_out/r2v_pytest_forced.log 2>&1
tail -15 gpurun_out/r2v_pytest_forced.log
(time timeout 900 python bench.py --steps 3 --warmup 3) > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err
tail -4 gpurun_out/r2v_bench.err; cut -c1-1500 gpurun_out/r2v_bench.json
