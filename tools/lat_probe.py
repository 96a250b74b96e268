import sys, time
sys.path.insert(0, '.')
import numpy as np
from libfriendship_b200 import B200Renderer
from workloads.graphs import build_cfg1_graph, cfg1_input
n = 48000
x = cfg1_input(n)
r = B200Renderer()
build_cfg1_graph(r)
out = np.zeros((2, n), dtype=np.float32)
for i in range(6):
    t0 = time.perf_counter(); r.fill_buffer(2, n, 0, [x], out=out); t1 = time.perf_counter()
    print("seek call", i, (t1 - t0) * 1e6, "us", r.stats()["schedule_builds"], r.stats()["jit_launches"])
r.set_profiling(True)
for i in range(3):
    t0 = time.perf_counter(); r.fill_buffer(2, n, 0, [x], out=out); t1 = time.perf_counter()
    print("profiled seek call", (t1 - t0) * 1e6, r.timing())
r.set_profiling(False)
idx = n
for i in range(4):
    t0 = time.perf_counter(); r.fill_buffer(2, n, idx, [x], out=out); t1 = time.perf_counter(); idx += n
    print("contiguous call", (t1 - t0) * 1e6, "us")
