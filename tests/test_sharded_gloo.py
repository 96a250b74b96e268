"""CPU, world_size 2, gloo: the N>1 path's host logic — round-robin voice shards, per-rank sub-graphs, one reduce
onto rank 0 — checked against the single-process render of the whole graph.  The per-rank renderer is the CPU
oracle injected into ShardedRenderer (no GPU here); on the GPU box the same class drives B200Renderer over NCCL."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_VOICES, N_PARTIALS, N = 5, 24, 700


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from workloads.banks import build_voice_mix_graph, detuned_bank
    from libfriendship_b200.sharded import ShardedRenderer
    from oracle.binding import OracleRenderer
    sr = ShardedRenderer(rank=rank, world_size=world, renderer=OracleRenderer())
    mine = sr.voices_of_rank(N_VOICES)
    bank, ids = detuned_bank(N_VOICES, N_PARTIALS, voices=mine)
    assert ids == mine
    build_voice_mix_graph(sr.r, bank, ids, delay0=100.0)
    out = sr.fill_buffer(1, N, 0)
    out2 = sr.fill_buffer(1, 300, N)            # a second, contiguous block
    blocks = []
    sr.render_stream(1, 0, N, 256, lambda blk, t: blocks.append((t, blk.copy())))   # N4: streamed in ragged blocks
    if rank == 0:
        assert [t for t, _ in blocks] == [0, 256, 512]
        assert np.array_equal(np.concatenate([b for _, b in blocks], axis=1), out)
        q.put((out, out2))
    else:
        assert out is None and not blocks
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_voice_shards_reduce_to_the_full_render():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, got2 = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from workloads.banks import build_voice_mix_graph, detuned_bank, full_scale
    from oracle.binding import OracleRenderer
    bank, ids = detuned_bank(N_VOICES, N_PARTIALS)
    r = OracleRenderer()
    build_voice_mix_graph(r, bank, ids, delay0=100.0)
    want = r.fill_buffer(1, N, 0)
    want2 = r.fill_buffer(1, 300, N)
    tol = 1e-6 * full_scale(bank) * N_VOICES     # only the f32 order of the top-level sum differs
    assert np.abs(got - want).max() <= tol
    assert np.abs(got2 - want2).max() <= tol


def test_round_robin_partition_covers_every_voice_once():
    from libfriendship_b200.sharded import voices_of_rank
    for n_voices in (1, 7, 64, 256):
        for world in (1, 2, 4, 8):
            seen = sorted(v for r in range(world) for v in voices_of_rank(n_voices, r, world))
            assert seen == list(range(n_voices))
            sizes = [len(voices_of_rank(n_voices, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
