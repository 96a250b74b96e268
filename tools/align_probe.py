"""GPU-box helper: a cfg3-shaped graph (1,024 voices) streamed in blocks whose starts are / are not multiples of 8:
the fused chain kernel's fast tiles need a 16-byte aligned block start (a multiple of 8 when it evaluates its exciters).
Prints ms per block."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

from libfriendship_b200 import B200Renderer
from workloads.banks import detuned_bank
from workloads.filters import build_cfg3_graph

# argv: [n_voices [block ...]]; FRB_NO_ALIGN_SPLIT=1 in the environment renders unaligned heads inside the block (the
# behaviour before run_range split them off)
N_VOICES = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
BLOCKS = tuple(int(a) for a in sys.argv[2:]) or (448, 441, 4096, 4097)
for flags, name in ((0, "exciters inside the chain kernel"), (16, "exciters on rings")):
    for block in BLOCKS:
        n_voices = N_VOICES
        r = B200Renderer(flags=flags)
        bank, _ = detuned_bank(n_voices, 1, seed=5)
        build_cfg3_graph(r, n_voices, excitation="osc", bank=bank, mix_to_one=True)
        out = torch.empty((1, block), dtype=torch.float32, device="cuda")
        idx = 0
        for _ in range(5):
            r.fill_buffer_device(out.data_ptr(), 1, block, idx, 0, None); idx += block
        r.sync()
        t0 = time.perf_counter()
        nb = 100 if block * n_voices < (1 << 24) else 20
        for _ in range(nb):
            r.fill_buffer_device(out.data_ptr(), 1, block, idx, 0, None); idx += block
        r.sync()
        print(json.dumps({"case": name, "voices": n_voices, "block": block, "align_split": os.environ.get("FRB_NO_ALIGN_SPLIT") != "1", "ms_per_block": (time.perf_counter() - t0) * 1e3 / nb}), flush=True)
