"""GPU: K5 — the cross-process mix without a collective: two ranks (two processes sharing cuda:0, gloo for the
barrier) store their shard's mix block straight into rank 0's slab through a CUDA-IPC mapping and rank 0 sums the
rows in rank order.  Must equal the single-process render of all voices up to the f32 order of the top-level sum."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_VOICES, N_PARTIALS, N = 6, 40, 5000


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from workloads.banks import build_voice_mix_graph, detuned_bank
    from libfriendship_b200.sharded import ShardedRenderer
    sr = ShardedRenderer(rank=rank, world_size=world, device=0, exchange="p2p")
    mine = sr.voices_of_rank(N_VOICES)
    bank, ids = detuned_bank(N_VOICES, N_PARTIALS, voices=mine)
    build_voice_mix_graph(sr.r, bank, ids, delay0=300.0)
    outs = []
    for k in range(3):                                   # three steps: both slabs are reused
        o = sr.fill_buffer(1, N, k * N)                   # (the returned host block is reused by the next call)
        outs.append(None if o is None else o.copy())
    if rank == 0:
        q.put(outs)
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_store_into_rank0_slab_and_sum_in_rank_order():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from workloads.banks import build_voice_mix_graph, detuned_bank, full_scale
    from libfriendship_b200 import B200Renderer
    bank, ids = detuned_bank(N_VOICES, N_PARTIALS)
    r = B200Renderer()
    build_voice_mix_graph(r, bank, ids, delay0=300.0)
    tol = 1e-6 * full_scale(bank) * N_VOICES
    for k in range(3):
        want = r.fill_buffer(1, N, k * N)
        assert np.abs(got[k] - want).max() <= tol, k
