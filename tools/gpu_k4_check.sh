set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none -k regex:frb_stage -s 4 -c 1 -f -o gpurun_out/prof_stage_r1o python tools/bench_kernels.py pure > gpurun_out/ncu_stage.log 2>&1
tail -2 gpurun_out/ncu_stage.log
