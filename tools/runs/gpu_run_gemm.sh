mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:osc_tc -c 1 -o gpurun_out/r3g_k1t python bench.py --steps 1 --warmup 1 --no-parity --no-extra --no-cpu-baseline > gpurun_out/r3g_ncu.log 2>&1
ls -la gpurun_out/r3g_k1t.ncu-rep
