mkdir -p gpurun_out
for c in 262144 1048576 4194304; do
  echo "block $c"; FRB_BLOCK_SAMPLES=$c timeout 200 python tools/bench_kernels.py pure elementwise 2>&1 | cut -c1-330
done > gpurun_out/k23_blocks.log 2>&1
cat gpurun_out/k23_blocks.log
