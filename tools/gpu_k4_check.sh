mkdir -p gpurun_out
for c in 4096 8192 16384 32768 65536 131072 262144; do
  echo "block $c"; FRB_BLOCK_SAMPLES=$c timeout 120 python tools/k4_probe.py base 2>&1 | tail -1
done > gpurun_out/k4_blocks.log 2>&1
cat gpurun_out/k4_blocks.log
