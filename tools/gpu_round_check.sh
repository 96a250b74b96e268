set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r1i_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1i_smoke.log 2>&1
python bench.py > gpurun_out/bench_r1i.json 2> gpurun_out/bench_r1i.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_r1i.json 2>> gpurun_out/bench_r1i.err
python tools/bench_kernels.py pure elementwise cfg3 cfg3_unfused cfg2 cfg1 > gpurun_out/kernels_r1i.jsonl 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1i.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_i1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:osc_kernel -s 12 -c 1 -f -o gpurun_out/prof_osc_r1i python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_i2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dfcomb -s 3 -c 1 -f -o gpurun_out/prof_dfcomb_r1i python tools/bench_kernels.py cfg3 > gpurun_out/ncu_i3.log 2>&1
cat gpurun_out/r1i_tests.log gpurun_out/r1i_smoke.log; cut -c1-400 gpurun_out/bench_r1i.json; cat gpurun_out/kernels_r1i.jsonl | cut -c1-600
