"""Times frb_define_oscbank for the cfg4 bank (pinned vs pageable host arrays); FRB_TRACE=1 prints the phases."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); import numpy as np
import torch
from workloads.banks import detuned_bank
from libfriendship_b200 import B200Renderer

nv, npart = int(sys.argv[1]) if len(sys.argv) > 1 else 64, int(sys.argv[2]) if len(sys.argv) > 2 else 65536
bank, ids = detuned_bank(nv, npart)
pinned = {k: (torch.from_numpy(np.ascontiguousarray(v)).pin_memory().numpy() if isinstance(v, np.ndarray) else v) for k, v in bank.items()}
r = B200Renderer()
for name, b in (("pageable", bank), ("pinned", pinned), ("pinned", pinned), ("pinned", pinned)):
    t0 = time.perf_counter()
    r.define_oscbank(7, **b)
    print(f"{name}: {1e3 * (time.perf_counter() - t0):.2f} ms", flush=True)
