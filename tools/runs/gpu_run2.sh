set -x
mkdir -p gpurun_out
(time timeout 1800 python -m pytest tests -m gpu -x -q) > gpurun_out/r2b_pytest.log 2>&1
tail -5 gpurun_out/r2b_pytest.log
timeout 300 python tools/bench_kernels.py refbank refbank256 cfg2 cfg2_64 elementwise pure > gpurun_out/r2b_kernels.jsonl 2>&1
cat gpurun_out/r2b_kernels.jsonl | cut -c1-420
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2b_cfg2_launches.csv python tools/bench_kernels.py cfg2 > gpurun_out/r2b_ncu_cfg2.log 2>&1
grep -c . gpurun_out/r2b_cfg2_launches.csv
