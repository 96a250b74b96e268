"""GPU-box helper: cfg3's fused chain kernel under different comb-delay distributions (which lanes take the in-tile
phase path, and where they sit in the grid).  One JSON line per variant.  usage: k4_probe.py [base ge256 ge512 shuffled]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import numpy as np

from workloads import filters
import bench_kernels

_orig = filters.cfg3_filters


def variant(name):
    def f(n_voices, sr=48000.0):
        coef, delay, gain = _orig(n_voices, sr)
        v = np.arange(n_voices)
        if name == "ge256":
            delay = (300 + v % 700).astype(np.uint32)
        elif name == "ge128":
            delay = (128 + v % 872).astype(np.uint32)
        elif name == "ge512":
            delay = (600 + v % 400).astype(np.uint32)
        elif name == "shuffled":
            delay = np.random.Generator(np.random.PCG64(3)).permutation(delay).astype(np.uint32)
        return coef, delay, gain
    return f


if __name__ == "__main__":
    for name in sys.argv[1:] or ["base", "ge256", "shuffled"]:
        ring = name.endswith("+ring")                         # exciters on their own kernel and rings (FRB_FLAG_NO_EXCITER_FUSION)
        filters.cfg3_filters = variant(name.split("+")[0])
        res = bench_kernels.case_cfg3(flags=16 if ring else 0)
        res["delays"] = name
        print(json.dumps({k: res[k] for k in ("delays", "ms", "scan_ms", "osc_ms", "fold_ms", "K4_frac")}), flush=True)
