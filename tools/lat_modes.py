"""cfg1 per-call latency through the Python binding in the three JIT modes (default: background compile, NO_JIT: interpreter,
JIT_EAGER: compile in the first call).  python tools/lat_modes.py on a GPU box; prints the median of the first 300 and last 1000 calls.
From plain C: build/bin/cfg1_latency with CFG1_FLAGS=<flags>.  FRB_TRACE_FILL=1 adds the host-clock phases of fill()."""
import sys, time
sys.path.insert(0, '.')
import numpy as np
from libfriendship_b200 import B200Renderer, FLAG_NO_JIT, FLAG_JIT_EAGER
from workloads.graphs import build_cfg1_graph, cfg1_input
x = cfg1_input(512 * 4000)
for name, flags in (("default", 0), ("no_jit", FLAG_NO_JIT), ("eager", FLAG_JIT_EAGER)):
    r = B200Renderer(flags=flags)
    build_cfg1_graph(r)
    out = np.zeros((2, 512), np.float32)
    ts = []
    switched = None
    for k in range(4000):
        blk = [x[k * 512:(k + 1) * 512]]
        t0 = time.perf_counter()
        r.fill_buffer(2, 512, k * 512, blk, out=out)
        ts.append(time.perf_counter() - t0)
        if switched is None and k % 50 == 0 and r.stats()["jit_launches"] > 0:
            switched = k
    ts = np.array(ts) * 1e6
    print(name, "switched_at<=", switched, "median first 300: %.1f us, last 1000: %.1f us" % (np.median(ts[:300]), np.median(ts[-1000:])), r.stats()["jit_launches"], r.stats()["interp_launches"], flush=True)
