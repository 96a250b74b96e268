set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2s_k4occ.jsonl
for rep in 1 2; do
for o in 0 8 9 10; do
  FRB_K4_SPT16=$o timeout 300 python tools/k4_probe.py base+ring 2>&1 | sed "s/^/minb=$o /" >> gpurun_out/r2s_k4occ.jsonl
done
for o in 0 8 88 87; do
  FRB_K4_EXC_MINB=$o timeout 300 python tools/k4_probe.py base 2>&1 | sed "s/^/exc_minb=$o /" >> gpurun_out/r2s_k4occ.jsonl
done; done
cut -c1-130 gpurun_out/r2s_k4occ.jsonl
