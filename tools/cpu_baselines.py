"""CPU baselines for every BASELINE.json configuration: the oracle (C++ restatement of the reference's per-sample
renderer; extension nodes evaluated the reference's way, f32 sinf/expf per partial-sample, sequential recurrences)
timed on THIS machine's host cores.  Large configurations are timed on a stated slice and extrapolated linearly.
Run on the GPU box so that the numbers sit beside the GPU numbers."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np

from workloads.banks import build_voice_mix_graph, detuned_bank, harmonic_bank
from workloads.filters import build_cfg3_graph
from workloads.graphs import build_cfg1_graph, cfg1_input
from oracle.binding import OracleRenderer

NT = os.cpu_count() or 1


def cfg1():
    n = 48000
    x = cfg1_input(n)
    r = OracleRenderer()
    build_cfg1_graph(r)
    t0 = time.perf_counter()
    r.fill_buffer(2, n, 0, [x])
    t1 = time.perf_counter() - t0
    r.fill_buffer(2, 0, n)                       # keep the input history, then evaluate in parallel over time
    _, tn = r.fill_buffer_mt(2, n, 0, NT)
    return {"config": "cfg1 48 kHz x 1 s, 2 output slots", "samples": n, "cpu_1thread_s": t1, "cpu_all_threads_s": tn, "threads": NT,
            "output_samples_per_s_1thread": 2 * n / t1, "output_samples_per_s_all": 2 * n / tn}


def cfg2():
    from libfriendship_b200 import KIND_OSCBANK
    n_slice = 24000                              # 0.5 s of the 10 s
    r = OracleRenderer(ext_mode="f32")
    r.define_oscbank(5, **harmonic_bank(1024))
    r.on_add_node(1, KIND_OSCBANK, 5)
    r.on_add_edge((1, 0, 0, 0))
    r.fill_buffer(1, 0, 0)
    _, t1 = r.fill_buffer_mt(1, n_slice // 8, 0, 1)
    _, tn = r.fill_buffer_mt(1, n_slice, 0, NT)
    ps1, psn = 1024 * (n_slice // 8) / t1, 1024 * n_slice / tn
    return {"config": "cfg2 1,024 partials x 1 voice x 10 s", "slice": f"{n_slice} samples (1 thread: {n_slice // 8})", "threads": NT,
            "partial_samples_per_s_1thread": ps1, "partial_samples_per_s_all": psn,
            "extrapolated_full_render_s_1thread": 1024 * 480000 / ps1, "extrapolated_full_render_s_all": 1024 * 480000 / psn}


def cfg3():
    n_voices, n = 16, 48000                      # slice: 16 of 4,096 voices, 1 of 10 s; recurrences are sequential in time
    bank, _ = detuned_bank(n_voices, 1, seed=5)
    r = OracleRenderer(ext_mode="f32")
    build_cfg3_graph(r, n_voices, excitation="osc", bank=bank, mix_to_one=False)
    t0 = time.perf_counter()
    r.fill_buffer(n_voices, n, 0)
    t1 = time.perf_counter() - t0
    vs = n_voices * n / t1
    return {"config": "cfg3 4,096 voices biquad + feedback delay x 10 s", "slice": f"{n_voices} voices x {n} samples, 1 thread (voices are independent: "
            f"x{NT} threads at best)", "voice_samples_per_s_1thread": vs, "extrapolated_full_render_s_1thread": 4096 * 480000 / vs,
            "extrapolated_full_render_s_all_threads_ideal": 4096 * 480000 / vs / NT, "threads": NT}


def cfg4():
    p, n = 4096, 12000
    bank, ids = detuned_bank(1, p)
    r = OracleRenderer(ext_mode="f32")
    build_voice_mix_graph(r, bank, ids)
    r.fill_buffer(1, 0, 0)
    _, t1 = r.fill_buffer_mt(1, n // 8, 4800, 1)
    _, tn = r.fill_buffer_mt(1, n, 4800, NT)
    ps1, psn = p * (n // 8) / t1, p * n / tn
    tot = 64 * 65536 * 480000
    return {"config": "cfg4 65,536 partials x 64 voices x 10 s", "slice": f"1 voice x {p} partials x {n} samples through the per-voice Delay/mix graph",
            "threads": NT, "partial_samples_per_s_1thread": ps1, "partial_samples_per_s_all": psn,
            "extrapolated_full_render_s_1thread": tot / ps1, "extrapolated_full_render_s_all": tot / psn}


if __name__ == "__main__":
    for fn in (cfg1, cfg2, cfg3, cfg4):
        print(json.dumps(fn()), flush=True)
