set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dfcomb -s 3 -c 1 -f -o gpurun_out/prof_dfcomb_r1n python tools/bench_kernels.py cfg3 > gpurun_out/ncu_k4.log 2>&1
tail -2 gpurun_out/ncu_k4.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:osc_one -s 3 -c 1 -f -o gpurun_out/prof_oscone_r1n python tools/bench_kernels.py cfg3_ring > gpurun_out/ncu_osc1.log 2>&1
tail -2 gpurun_out/ncu_osc1.log
timeout 300 python tools/bench_kernels.py cfg3 cfg3_ring cfg3_unfused 2>/dev/null | cut -c1-900 > gpurun_out/kernels_cfg3.jsonl
cat gpurun_out/kernels_cfg3.jsonl
