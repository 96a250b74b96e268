set -x
mkdir -p gpurun_out
for b in 0 1 2; do
  (FRB_K4_BULK=$b timeout 900 python -m pytest tests/test_recurrences.py tests/test_full_size.py -x -q -m gpu) > gpurun_out/r2d_pytest_k4_bulk$b.log 2>&1
  tail -2 gpurun_out/r2d_pytest_k4_bulk$b.log
done
for rep in 1 2; do for b in 0 1 2; do
  FRB_K4_BULK=$b timeout 300 python tools/bench_kernels.py cfg3_ring 2>&1 | sed "s/^/bulk=$b /" >> gpurun_out/r2d_k4_ab.jsonl
done; done
for b in 0 2; do
  FRB_K4_BULK=$b timeout 300 python tools/bench_kernels.py cfg3 2>&1 | sed "s/^/bulk=$b /" >> gpurun_out/r2d_k4_ab.jsonl
done
cut -c1-330 gpurun_out/r2d_k4_ab.jsonl
