set -x
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/r2t_pytest.log 2>&1
tail -3 gpurun_out/r2t_pytest.log
(time timeout 900 python bench.py --steps 3 --warmup 3) > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err
tail -4 gpurun_out/r2t_bench.err; cut -c1-200 gpurun_out/r2t_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2t_launches.csv python bench.py --steps 2 --warmup 1 --no-parity --no-extra --no-cpu-baseline > gpurun_out/r2t_ncu_launches.log 2>&1
grep -c . gpurun_out/r2t_launches.csv
(timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print(\"smoke ok\")") 2>&1 | tail -2
