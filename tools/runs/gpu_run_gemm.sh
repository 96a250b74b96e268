mkdir -p gpurun_out
rm -f gpurun_out/r2y_tc.jsonl
FRB_OSC_GEMM=3 timeout 60 build/bin/osc_gemm_check 8 4096 40000 200 >> gpurun_out/r2y_tc.jsonl 2>&1; echo "rc=$?" >> gpurun_out/r2y_tc.jsonl
FRB_OSC_GEMM=3 timeout 60 build/bin/osc_gemm_check 8 4096 40000 200 5000 >> gpurun_out/r2y_tc.jsonl 2>&1; echo "rc=$?" >> gpurun_out/r2y_tc.jsonl
FRB_OSC_GEMM=3 timeout 60 build/bin/osc_gemm_check 5 1000 70000 300 >> gpurun_out/r2y_tc.jsonl 2>&1; echo "rc=$?" >> gpurun_out/r2y_tc.jsonl
FRB_OSC_GEMM=2 timeout 60 build/bin/osc_gemm_check 5 1000 70000 300 >> gpurun_out/r2y_tc.jsonl 2>&1; echo "rc=$?" >> gpurun_out/r2y_tc.jsonl
FRB_OSC_GEMM=3 timeout 100 build/bin/osc_gemm_check 64 65536 480000 40 >> gpurun_out/r2y_tc.jsonl 2>&1; echo "rc=$?" >> gpurun_out/r2y_tc.jsonl
cat gpurun_out/r2y_tc.jsonl
(FRB_OSC_GEMM=4 timeout 600 python bench.py --steps 3 --warmup 3 --no-extra) > gpurun_out/r2y_bench_tc.json 2> gpurun_out/r2y_bench_tc.err
tail -3 gpurun_out/r2y_bench_tc.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2y_bench_tc.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline'].get('k1_ms_per_step'), d.get('parity',{}).get('max_err_of_full_scale'))
PY
