mkdir -p gpurun_out
for q in 16 32 8; do
  FRB_JIT_QUADS=$q timeout 300 python tools/bench_kernels.py refbank refbank256 2>&1 | sed "s/^/quads=$q /" | cut -c1-330 >> gpurun_out/r2r_quads.txt
done
cat gpurun_out/r2r_quads.txt
timeout 300 python -m pytest tests/test_edge_cases.py tests/test_jit_loops_gpu.py -x -q -m gpu 2>&1 | tail -2
timeout 300 python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r2r_bench.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2r_bench.json')); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['parity']['ok'])"
