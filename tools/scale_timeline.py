"""GPU-box helper: where does a cfg4 step go on every rank at N GPUs?  (VERDICT r1 'weak' 6: 0.83 ms lost per step at N = 8.)

Under torchrun.  Per rank, CUDA events on the renderer's stream around one step: render (all of frb_fill_buffer_device) and
exchange (the NCCL reduce ordered on the same stream); then one step with frb_set_profiling: the K1 family, the mix stage
and the fill's total on the device; and the host time one step's calls take to ENQUEUE.  Rank 0 prints one JSON object:
the per-rank rows, the step time (max over ranks) and the ideal (1-GPU K1 time / N).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 tools/scale_timeline.py
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist

from libfriendship_b200.sharded import ShardedRenderer
from workloads.banks import build_voice_mix_graph, detuned_bank

NV, NP, NS = 64, 65536, 480000
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
sr = ShardedRenderer(rank=rank, world_size=world, device=local)
mine = sr.voices_of_rank(NV)
bank, ids = detuned_bank(NV, NP, voices=mine)
build_voice_mix_graph(sr.r, bank, ids)
st = sr.cuda_stream()
for _ in range(3):
    sr.fill_buffer_device(1, NS, 0)
sr.r.sync()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()

rows = []
for rep in range(3):
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    out = sr._block(1, NS)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    h0 = time.perf_counter()
    e0.record(st)
    sr.r.fill_buffer_device(out.data_ptr(), 1, NS, 0)
    e1.record(st)
    sr._reduce(out)
    e2.record(st)
    h1 = time.perf_counter()
    e2.synchronize()
    rows.append({"render_ms": e0.elapsed_time(e1), "exchange_ms": e1.elapsed_time(e2), "step_ms": e0.elapsed_time(e2),
                 "host_enqueue_ms": (h1 - h0) * 1e3})
best = min(rows, key=lambda r: r["step_ms"])
sr.r.set_profiling(True)
sr.r.fill_buffer_device(sr._block(1, NS).data_ptr(), 1, NS, 0)
sr.r.sync()
tim = sr.r.timing()
sr.r.set_profiling(False)
s0 = sr.r.stats()
sr.r.fill_buffer_device(sr._block(1, NS).data_ptr(), 1, NS, 0)
sr.r.sync()
s1 = sr.r.stats()
mine_row = dict(rank=rank, voices=len(mine), **best, k1_ms=tim["osc_ms"], stage_ms=tim["interp_ms"], fill_total_ms_profiled=tim["total_ms"],
                launches_per_step=int(s1["kernel_launches"] - s0["kernel_launches"]), k1_launches=int(s1["osc_launches"] - s0["osc_launches"]))
if world > 1:
    allrows = [None] * world
    dist.all_gather_object(allrows, mine_row)
else:
    allrows = [mine_row]
if rank == 0:
    print(json.dumps({"case": "cfg4 step timeline per rank", "n_gpus": world, "step_ms_max_over_ranks": max(r["step_ms"] for r in allrows),
                      "k1_ms_max": max(r["k1_ms"] for r in allrows), "ranks": allrows}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
