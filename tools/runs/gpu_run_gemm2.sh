mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/r2z_pytest.log 2>&1
tail -5 gpurun_out/r2z_pytest.log
(timeout 900 python bench.py --steps 3 --warmup 3) > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err
tail -4 gpurun_out/r2z_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2z_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','dtype')}, d['e2e']['value'], d['e2e']['ms_per_step'], d.get('parity',{}).get('max_err_of_full_scale'))
print(json.dumps({k:v for k,v in d['roofline'].items() if not k.endswith('note')}))
print(json.dumps(d.get('extra'))[:1500])
PY
