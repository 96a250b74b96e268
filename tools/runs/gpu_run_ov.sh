set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_oscbank.py tests/test_full_size.py tests/test_stream.py -x -q -m gpu 2>&1 | tail -2
for ov in 1 0 1 0; do
  FRB_OSC_REDUCE_OVERLAP=$ov timeout 300 python bench.py --steps 3 --warmup 3 --no-parity --no-extra --no-cpu-baseline > gpurun_out/r2g_bench_ov$ov.json 2>/dev/null
  python -c "import json; d=json.load(open('gpurun_out/r2g_bench_ov$ov.json')); print('N=1 ov=$ov', d['ms_per_step'], d['value'], d['roofline']['frac'])"
done
# the 8-GPU shard on one GPU: 8 voices
for ov in 1 0 1 0; do
  FRB_OSC_REDUCE_OVERLAP=$ov timeout 300 python bench.py --steps 5 --warmup 3 --voices 8 --no-parity --no-extra --no-cpu-baseline > gpurun_out/r2g_bench8v_ov$ov.json 2>/dev/null
  python -c "import json; d=json.load(open('gpurun_out/r2g_bench8v_ov$ov.json')); print('8 voices ov=$ov', d['ms_per_step'], d['roofline']['k1_ms_per_step'])"
done
