"""BASELINE.json configs[4] — offline large render: 1M partials x 256 voices, 192 kHz x 60 s, 8 x B200, output reduced
to rank 0.  Launch:  python -m torch.distributed.run --nproc-per-node 8 tools/render_cfg5.py
One render (no warm-up: an offline job runs once); prints one JSON line on rank 0.
The output leaves the device through the N4 streaming path (ShardedRenderer.render_stream): blocks of CFG5_BLOCK samples
(default 2^20) are reduced onto rank 0, copied to pinned host memory and appended to a float32 WAV file (CFG5_WAV, default
/tmp/cfg5.wav; empty = keep the blocks in memory only) while the next block renders."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist

from workloads.banks import build_voice_mix_graph, detuned_bank
from libfriendship_b200.sharded import ShardedRenderer

SR, N_VOICES, N_PARTIALS, N_SAMPLES = 192000.0, 256, 1 << 20, 11_520_000


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    n_voices = int(os.environ.get("CFG5_VOICES", N_VOICES))
    n_samples = int(os.environ.get("CFG5_SAMPLES", N_SAMPLES))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    t_setup = time.perf_counter()
    sr = ShardedRenderer(rank=rank, world_size=world, device=local)
    mine = sr.voices_of_rank(n_voices)
    bank, ids = detuned_bank(n_voices, N_PARTIALS, sr=SR, seed=2, voices=mine)
    build_voice_mix_graph(sr.r, bank, ids)
    del bank
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_setup = time.perf_counter() - t_setup
    block = int(os.environ.get("CFG5_BLOCK", 1 << 20))
    wav_path = os.environ.get("CFG5_WAV", "/tmp/cfg5.wav")
    wav = None
    if rank == 0 and wav_path:
        from libfriendship_b200.dispatch import WavClient
        os.makedirs(os.path.dirname(wav_path), exist_ok=True)
        wav = WavClient(wav_path, 1, int(SR))
    out = np.zeros((1, n_samples), dtype=np.float32) if rank == 0 else None

    def sink(blk, t):                              # rank 0, while the next block renders
        out[:, t:t + blk.shape[1]] = blk
        if wav:
            wav.audio_rendered(blk, t)

    t0 = time.perf_counter()
    sr.render_stream(1, 0, n_samples, block, sink)   # render + reduce to rank 0 + pipelined D2H + file
    if wav:
        wav.close()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    parity = None
    if rank == 0 and os.environ.get("CFG5_CHECK", "1") != "0":
        # the rendered file against the fp64 oracle evaluated on the FULL bank (every voice, every partial, both delay
        # taps) in three short windows: the start, past the last voice's delay tap, and the last samples at t ~ 11.5 M,
        # where an oscillator that accumulated its phase in f32 would be off by a radian
        from oracle.binding import OracleRenderer
        from workloads.banks import full_scale
        tc = time.perf_counter()
        full, fids = detuned_bank(n_voices, N_PARTIALS, sr=SR, seed=2)
        orc = OracleRenderer(ext_mode="fp64")
        build_voice_mix_graph(orc, full, fids)
        orc.fill_buffer(1, 0, 0)
        fs = full_scale(full) * n_voices * 1.3
        nthr = os.cpu_count() or 1
        width = int(os.environ.get("CFG5_CHECK_WIDTH", 4))
        wins = [0, 4800 + 37 * (n_voices - 1) + 1, n_samples - width]
        worst = 0.0
        for s0 in wins:
            want, _ = orc.fill_buffer_mt(1, width, s0, nthr)
            worst = max(worst, float(np.abs(out[0, s0:s0 + width].astype(np.float64) - want[0].astype(np.float64)).max()) / fs)
        parity = {"max_err_of_full_scale": worst, "tol": 1e-5, "ok": bool(worst <= 1e-5), "windows": [[int(w), width] for w in wins],
                  "full_scale": fs, "oracle": "fp64, full bank %d x %d, %d host threads, %.1f s" % (n_voices, N_PARTIALS, nthr, time.perf_counter() - tc)}
    if rank == 0:
        ps = n_voices * N_PARTIALS * n_samples
        print(json.dumps({"parity": parity,
            "workload": "cfg5: 2^20 partials x %d voices, 192 kHz x %.1f s (BASELINE.json configs[4])" % (n_voices, n_samples / SR),
            "n_gpus": world, "render_s": dt, "setup_s": t_setup, "partial_samples": ps, "partial_samples_per_s": ps / dt,
            "realtime_factor": (n_samples / SR) / dt, "out_bytes": int(out.nbytes), "block": block,
            "wav": (wav_path or None), "finite": bool(np.isfinite(out).all()),
            "peak_abs": float(np.abs(out).max()), "rms_last_second": float(np.sqrt(np.mean(out[0, -192000:] ** 2)))}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
