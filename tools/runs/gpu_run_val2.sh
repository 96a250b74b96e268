set -x
mkdir -p gpurun_out
(time timeout 1800 python -m pytest tests -m gpu -x -q) > gpurun_out/r2j_pytest.log 2>&1
tail -4 gpurun_out/r2j_pytest.log
timeout 300 python tools/bench_kernels.py pure elementwise refbank256 cfg1 > gpurun_out/r2j_kernels.jsonl 2>&1
cut -c1-260 gpurun_out/r2j_kernels.jsonl
timeout 300 build/bin/cfg1_latency 512 2000
(time timeout 900 python bench.py --steps 3 --warmup 3) > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err
tail -4 gpurun_out/r2j_bench.err; cut -c1-200 gpurun_out/r2j_bench.json
