mkdir -p gpurun_out
rm -f gpurun_out/r2u_gemm.jsonl
for m in 0 2; do
  FRB_OSC_GEMM=$m timeout 120 build/bin/osc_gemm_check 8 4096 40000 200 >> gpurun_out/r2u_gemm.jsonl 2>&1
  FRB_OSC_GEMM=$m timeout 120 build/bin/osc_gemm_check 8 4096 40000 200 5000 >> gpurun_out/r2u_gemm.jsonl 2>&1
done
FRB_OSC_GEMM=2 timeout 120 build/bin/osc_gemm_check 3 24 20000 100 >> gpurun_out/r2u_gemm.jsonl 2>&1
FRB_OSC_GEMM=0 timeout 200 build/bin/osc_gemm_check 64 65536 480000 6 >> gpurun_out/r2u_gemm.jsonl 2>&1
FRB_OSC_GEMM=2 timeout 200 build/bin/osc_gemm_check 64 65536 480000 6 >> gpurun_out/r2u_gemm.jsonl 2>&1
cat gpurun_out/r2u_gemm.jsonl
