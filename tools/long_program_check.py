"""GPU-box helper (not yet run: DESIGN.md §9): a Sum2 chain of N nodes (default 70,000) is ONE stage of N + 2
instructions — above the 48 KB the interpreter kernel stages in shared memory and above the 65,536 instructions the
stage JIT accepts — rendered and compared bit for bit with the same f32 additions in numpy (same order, so the
roundings are the same)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from libfriendship_b200 import B200Renderer, KIND_F32CONSTANT, KIND_SUM2

n = int(sys.argv[1]) if len(sys.argv) > 1 else 70000
n_times = 4096
r = B200Renderer()
r.on_add_node(1, KIND_F32CONSTANT)
prev = 0
for i in range(n):
    h = 2 + i
    r.on_add_node(h, KIND_SUM2)
    r.on_add_edge((prev, h, 0, 0))
    r.on_add_edge((1, h, 0x3F800000, 1))          # + 1.0f
    prev = h
r.on_add_edge((prev, 0, 0, 0))
x = np.random.Generator(np.random.PCG64(3)).uniform(-1, 1, n_times).astype(np.float32)
t0 = time.perf_counter()
got = r.fill_buffer(1, n_times, 0, [x])
dt = time.perf_counter() - t0
want = x.copy()
one = np.float32(1.0)
for _ in range(n):
    want = want + one                              # float32 + float32: one rounding per node, like the chain
same = bool(np.array_equal(got[0].view(np.uint32), want.view(np.uint32)))
print(json.dumps({"case": "Sum2 chain in one stage", "nodes": n, "samples": n_times, "bit_exact": same,
                  "first_call_s": dt, "stats": {k: v for k, v in r.stats().items() if "launch" in k}}))
sys.exit(0 if same else 1)
