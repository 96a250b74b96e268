set -x
mkdir -p gpurun_out
for pf in 0 1 0 1; do
  FRB_OSC_PREFETCH=$pf timeout 120 python tools/bench_kernels.py cfg2 2>&1 | sed "s/^/pf=$pf /" >> gpurun_out/r2e_pf_cfg2.txt
done
for pf in 0 1; do for mg in 8 4 2; do
  FRB_OSC_PREFETCH=$pf FRB_OSC_MIN_GROUPS=$mg timeout 120 python tools/bench_kernels.py cfg2 2>&1 | sed "s/^/pf=$pf mg=$mg /" >> gpurun_out/r2e_pf_cfg2.txt
done; done
cut -c1-200 gpurun_out/r2e_pf_cfg2.txt
for pf in 0 1; do
  FRB_OSC_PREFETCH=$pf timeout 300 python bench.py --steps 3 --warmup 3 --no-parity --no-extra --no-cpu-baseline > gpurun_out/r2e_bench_pf$pf.json 2>/dev/null
  python -c "import json; d=json.load(open('gpurun_out/r2e_bench_pf$pf.json')); print('pf=$pf', d['ms_per_step'], d['value'], d['roofline']['frac'])"
done
FRB_OSC_PREFETCH=1 timeout 600 python -m pytest tests/test_oscbank.py tests/test_full_size.py -x -q -m gpu 2>&1 | tail -2
