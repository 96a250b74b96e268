#!/bin/bash
# tuning aid: K1 partial-range split target (voices x split) vs throughput, 64 and 8 voices per GPU
for t in 256 1024 2048 4096; do
  for v in 64 8; do
    echo -n "target=$t voices=$v: "
    FRB_OSC_SPLIT_TARGET=$t python bench.py --steps 3 --warmup 2 --no-cpu-baseline --voices $v 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.4e ms/step=%.2f exec=%.3f' % (d['value'], d['ms_per_step'], d['roofline']['executed_frac']))"
  done
done
