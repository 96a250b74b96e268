set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt
(time timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/r2a_pytest.log 2>&1
tail -5 gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -c 600 gpurun_out/r2a_bench.err
timeout 300 python tools/bench_kernels.py refbank refbank256 refbank_nojit cfg2 > gpurun_out/r2a_kernels.jsonl 2>&1
for mg in 8 4 2 1; do for tg in 1024 4096; do
  FRB_OSC_MIN_GROUPS=$mg FRB_OSC_SPLIT_TARGET=$tg timeout 120 python tools/bench_kernels.py cfg2 cfg2_64 2>&1 | sed "s/^/mg=$mg tg=$tg /" >> gpurun_out/r2a_cfg2_sweep.txt
done; done
FRB_OSC_ATTACK_SAME_L=1 timeout 120 python tools/bench_kernels.py cfg2 2>&1 | sed "s/^/sameL /" >> gpurun_out/r2a_cfg2_sweep.txt
cat gpurun_out/r2a_kernels.jsonl
cat gpurun_out/r2a_cfg2_sweep.txt | cut -c1-260
