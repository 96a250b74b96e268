set -x
mkdir -p gpurun_out
(time timeout 1800 python -m pytest tests -m gpu -x -q) > gpurun_out/r2h_pytest.log 2>&1
tail -4 gpurun_out/r2h_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2h_smoke.log 2>&1; tail -3 gpurun_out/r2h_smoke.log
timeout 300 python tools/bench_kernels.py pure elementwise refbank refbank256 cfg1 cfg2 cfg3 cfg3_ring > gpurun_out/r2h_kernels.jsonl 2>&1
cut -c1-300 gpurun_out/r2h_kernels.jsonl
(time timeout 900 python bench.py --steps 3 --warmup 3) > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
tail -4 gpurun_out/r2h_bench.err; cut -c1-300 gpurun_out/r2h_bench.json
(time timeout 300 python bench.py --impl reference --steps 1 --warmup 0) > gpurun_out/r2h_bench_ref.json 2> gpurun_out/r2h_bench_ref.err
cut -c1-300 gpurun_out/r2h_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2h_launches.csv python bench.py --steps 2 --warmup 1 --no-parity --no-extra --no-cpu-baseline > gpurun_out/r2h_ncu_launches.log 2>&1
grep -c . gpurun_out/r2h_launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:osc_kernel -s 2 -c 1 -o gpurun_out/r2h_osc_full python bench.py --steps 1 --warmup 1 --no-parity --no-extra --no-cpu-baseline > gpurun_out/r2h_ncu_full.log 2>&1
ls -la gpurun_out/r2h_osc_full.ncu-rep
