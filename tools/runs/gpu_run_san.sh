set -x
mkdir -p gpurun_out
timeout 120 build/bin/cfg4_multi 1 2 > gpurun_out/r2o_cfg4_c_first.json 2>&1; cat gpurun_out/r2o_cfg4_c_first.json
cat > /tmp/san_smoke.py <<'PY'
import sys
sys.path.insert(0, '.')
import numpy as np
import __graft_entry__ as g
g.smoke()
# the stage JIT with loops, the reference-vocabulary bank and a fused chain with exciters, small
from libfriendship_b200 import B200Renderer, FLAG_JIT_EAGER, KIND_OSCBANK
from workloads.banks import build_partial_sum_graph, harmonic_bank, partial_signals, detuned_bank
from workloads.filters import build_cfg3_graph
bank = harmonic_bank(64)
r = B200Renderer(flags=FLAG_JIT_EAGER)
build_partial_sum_graph(r, bank["amp"])
out = r.fill_buffer(1, 1000, 0, list(partial_signals(bank, 1000)))
print("jit loops ok", float(np.abs(out).max()), r.stats()["jit_launches"])
b1, _ = detuned_bank(8, 1, seed=5)
r = B200Renderer()
build_cfg3_graph(r, 8, excitation="osc", bank=b1, mix_to_one=True)
out = r.fill_buffer(1, 5000, 0)
out2 = r.fill_buffer(1, 3001, 5000)
print("fused chain ok", float(np.abs(out).max()), r.stats()["chain_launches"])
PY
(timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python /tmp/san_smoke.py) > gpurun_out/r2o_memcheck.log 2>&1; tail -8 gpurun_out/r2o_memcheck.log
(timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python /tmp/san_smoke.py) > gpurun_out/r2o_racecheck.log 2>&1; tail -6 gpurun_out/r2o_racecheck.log
