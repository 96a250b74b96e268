set -x
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/r2n_pytest.log 2>&1
tail -3 gpurun_out/r2n_pytest.log
for th in 1 0 1; do
FRB_JIT_TOP_HOIST=$th timeout 300 python tools/bench_kernels.py pure elementwise refbank256 refbank 2>&1 | sed "s/^/tophoist=$th /" | cut -c1-330 >> gpurun_out/r2n_kernels.txt
done
cat gpurun_out/r2n_kernels.txt
timeout 120 build/bin/cfg1_latency 512 2000
