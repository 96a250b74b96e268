mkdir -p gpurun_out
for c in 10 12 16 20 24; do
  FRB_JIT_CTAS_PER_SM=$c timeout 120 python tools/bench_kernels.py pure elementwise 2>&1 | sed "s/^/ctas=$c /" | cut -c1-200 >> gpurun_out/r2i_k2.txt
done
FRB_JIT_CTAS_PER_SM=16 FRB_JIT_NO_BOUNDS=1 timeout 120 python tools/bench_kernels.py pure elementwise 2>&1 | sed "s/^/ctas=16 nobounds /" | cut -c1-200 >> gpurun_out/r2i_k2.txt
cat gpurun_out/r2i_k2.txt
# K1 cluster A/B: cfg2, 8-voice shard, 64 voices; parity tests under the cluster kernel
timeout 900 python -m pytest tests/test_oscbank.py tests/test_full_size.py tests/test_jit_loops_gpu.py -x -q -m gpu 2>&1 | tail -2
for cl in 1 0 1 0; do
  FRB_OSC_CLUSTER=$cl timeout 120 python tools/bench_kernels.py cfg2 2>&1 | sed "s/^/cluster=$cl /" | cut -c1-160
done
for cl in 1 0 1 0; do
  FRB_OSC_CLUSTER=$cl timeout 300 python bench.py --steps 5 --warmup 3 --voices 8 --no-parity --no-extra --no-cpu-baseline > gpurun_out/r2i_bench8v_cl$cl.json 2>/dev/null
  python -c "import json; d=json.load(open('gpurun_out/r2i_bench8v_cl$cl.json')); print('8 voices cluster=$cl', d['ms_per_step'], d['roofline']['k1_ms_per_step'], d['roofline']['k1_family_launches_per_step'])"
done
for cl in 1 0 1 0; do
  FRB_OSC_CLUSTER=$cl timeout 300 python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r2i_bench_cl$cl.json 2>/dev/null
  python -c "import json; d=json.load(open('gpurun_out/r2i_bench_cl$cl.json')); print('64 voices cluster=$cl', d['ms_per_step'], d['value'], d['roofline']['frac'], d['parity']['max_err_of_full_scale'])"
done
