"""GPU-box helper: is the N-GPU render the 1-GPU render?  (VERDICT r1 'missing' 2 / SURVEY.md §4 item 4.)

The whole cfg4 render (65,536 partials x 64 voices x 480,000 samples unless shrunk) on N GPUs against the same render on
one GPU, sample by sample: <= 4e-7 of full scale (voices are the same bits wherever they are rendered; only the order of
the mix additions differs).  Two ways to N GPUs, both checked:
  --mode ranks    one process per GPU under torchrun, libfriendship_b200/sharded.py, exchange = nccl and p2p
  --mode inproc   ONE process, frb_config::n_devices = N behind the C ABI (csrc/multi.cu)
One JSON line per case; exit status 1 if any case is out of tolerance.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/multi_gpu_check.py --mode ranks
  python tools/multi_gpu_check.py --mode inproc --gpus 2
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np

TOL = 4e-7


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", choices=["ranks", "inproc"], default="inproc")
    ap.add_argument("--gpus", type=int, default=2)
    ap.add_argument("--voices", type=int, default=64)
    ap.add_argument("--partials", type=int, default=65536)
    ap.add_argument("--samples", type=int, default=480000)
    args = ap.parse_args()
    import torch
    from libfriendship_b200 import B200Renderer
    from workloads.banks import build_voice_mix_graph, detuned_bank, full_scale

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    nv, npart, ns = args.voices, args.partials, args.samples
    bad = 0

    def report(case, n_gpus, got, want, fs, extra=None):
        nonlocal bad
        err = float(np.abs(got.astype(np.float64) - want.astype(np.float64)).max()) / fs
        ok = err <= TOL
        bad += not ok
        line = {"case": case, "n_gpus": n_gpus, "voices": nv, "partials": npart, "samples": ns,
                "max_abs_diff_of_full_scale": err, "tol": TOL, "ok": ok, "peak_of_full_scale": float(np.abs(want).max()) / fs}
        line.update(extra or {})
        print(json.dumps(line), flush=True)

    def one_gpu_render(device):
        bank, ids = detuned_bank(nv, npart)
        r = B200Renderer(device=device)
        build_voice_mix_graph(r, bank, ids)
        out = r.fill_buffer(1, ns, 0)
        return out, full_scale(bank) * nv * 1.3, bank, ids

    if args.mode == "inproc":
        want, fs, bank, ids = one_gpu_render(0)
        for n in sorted({2, args.gpus}):
            if n > torch.cuda.device_count():
                continue
            r = B200Renderer(n_devices=n)
            build_voice_mix_graph(r, bank, ids)
            r.fill_buffer(1, ns, 0)                                   # first call: schedules, JIT, buffers
            t0 = time.perf_counter()
            got = r.fill_buffer(1, ns, 0)
            dt = time.perf_counter() - t0
            report("one process, frb_config.n_devices = N (csrc/multi.cu) vs one device", n, got, want, fs,
                   {"host_ms_second_call": dt * 1e3, "partial_samples_per_s": nv * npart * ns / dt})
        sys.exit(1 if bad else 0)

    import torch.distributed as dist
    from libfriendship_b200.sharded import ShardedRenderer
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    want = fs = None
    if rank == 0:
        want, fs, _, _ = one_gpu_render(local)
    for exchange in ("nccl", "p2p"):
        sr = ShardedRenderer(rank=rank, world_size=world, device=local, exchange=exchange)
        mine = sr.voices_of_rank(nv)
        bank, ids = detuned_bank(nv, npart, voices=mine)
        build_voice_mix_graph(sr.r, bank, ids)
        got = sr.fill_buffer(1, ns, 0)
        got = sr.fill_buffer(1, ns, 0)
        if rank == 0:
            report(f"one process per GPU (sharded.py), exchange = {exchange}, vs one device", world, got.copy(), want, fs)
        dist.barrier()
        del sr
    flag = torch.tensor([bad], device="cuda")
    dist.broadcast(flag, src=0)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(1 if int(flag.item()) else 0)


if __name__ == "__main__":
    main()
