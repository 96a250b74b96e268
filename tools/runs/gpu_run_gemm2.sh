mkdir -p gpurun_out
rm -f gpurun_out/r3i.jsonl
FRB_OSC_GEMM=3 timeout 60 build/bin/osc_gemm_check 8 4096 40000 200 >> gpurun_out/r3i.jsonl 2>&1; echo "rc=$?" >> gpurun_out/r3i.jsonl
FRB_OSC_GEMM=3 timeout 60 build/bin/osc_gemm_check 5 1000 70000 300 7001 >> gpurun_out/r3i.jsonl 2>&1; echo "rc=$?" >> gpurun_out/r3i.jsonl
FRB_OSC_GEMM=3 timeout 60 build/bin/osc_gemm_check 4 40 30000 300 >> gpurun_out/r3i.jsonl 2>&1; echo "rc=$?" >> gpurun_out/r3i.jsonl
cut -c1-420 gpurun_out/r3i.jsonl
(timeout 900 python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline) > gpurun_out/r3i_bench.json 2> gpurun_out/r3i_bench.err
tail -3 gpurun_out/r3i_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3i_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['k1_ms_per_step'], d.get('parity',{}).get('max_err_of_full_scale'))
PY
