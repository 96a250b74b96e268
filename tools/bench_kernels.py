"""GPU-box helper: measures the HBM-bound kernels (K2/K3 fused elementwise + Delay, K4 recurrences) against the
measured HBM peak, device-resident, CUDA-event timed through frb_set_profiling.  Prints one JSON line per case."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from libfriendship_b200 import (B200Renderer, KIND_DELAY, KIND_F32CONSTANT, KIND_MULTIPLY, KIND_SUM2)
from workloads.graphs import GraphBuilder, f32_bits

HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6535.7) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6535.7


def timed_fill(r, d_out, n_slots, n, idx, d_in=0, offs=None, reps=5):
    for _ in range(2):
        r.fill_buffer_device(d_out.data_ptr(), n_slots, n, 0, d_in, offs)   # idx 0 every time = seek: full recompute
    r.sync()
    r.set_profiling(True)
    best = None
    for _ in range(reps):
        r.fill_buffer_device(d_out.data_ptr(), n_slots, n, 0, d_in, offs)
        r.sync()
        t = r.timing()
        if best is None or t["total_ms"] < best["total_ms"]:
            best = t
    r.set_profiling(False)
    return best


def case_pure_continuing(n_slots=64, n=1 << 22, reps=5):
    """The pure elementwise graph in consecutive calls (idx advances: no seek, so none of the per-slot history resets
    a seek costs): what the ingest kernel itself takes = total_ms - interp_ms of the best call."""
    from libfriendship_b200 import KIND_MINIMUM
    r = B200Renderer()
    g = GraphBuilder(r)
    for s in range(n_slots):
        x = g.input(s)
        a = g.node(KIND_MULTIPLY, x, g.const(0.5))
        b = g.node(KIND_SUM2, a, g.const(0.25))
        g.output(s, g.node(KIND_MINIMUM, b, x))
    x = torch.rand((n_slots, n), dtype=torch.float32, device="cuda")
    out = torch.empty((n_slots, n), dtype=torch.float32, device="cuda")
    offs = np.arange(n_slots + 1, dtype=np.uint64) * n
    idx = 0
    for _ in range(2):
        r.fill_buffer_device(out.data_ptr(), n_slots, n, idx, x.data_ptr(), offs)
        idx += n
    r.sync()
    r.set_profiling(True)
    calls = []
    for _ in range(reps):
        r.fill_buffer_device(out.data_ptr(), n_slots, n, idx, x.data_ptr(), offs)
        r.sync()
        idx += n
        calls.append(r.timing())
    r.set_profiling(False)
    t = min(calls, key=lambda c: c["total_ms"])      # a call in which no history buffer had to grow
    ingest = t["total_ms"] - t["interp_ms"]
    return {"case": "K2 pure elementwise, consecutive calls (no seek)", "slots": n_slots, "samples": n, "ms": t["total_ms"],
            "interp_ms": t["interp_ms"], "ingest_ms": ingest, "ingest_GBs": 8.0 * n_slots * n / ingest / 1e6, "peak_GBs": HBM,
            "all_calls_ms": [c["total_ms"] for c in calls]}


def case_elementwise(n_slots=64, n=1 << 22):
    """out_s = 0.5*in_s + 0.35*Delay(0.5*in_s, 12000)  for s in slots: reads one input plane, writes one output plane
    (+ the materialised delay source: one ring write + one ring read).  Algorithmic: 8 B/sample (in + out)."""
    r = B200Renderer()
    g = GraphBuilder(r)
    for s in range(n_slots):
        x = g.input(s)
        gain = g.node(KIND_MULTIPLY, x, g.const(0.5))
        d = g.node(KIND_DELAY, gain, g.const(12000.0))
        e = g.node(KIND_MULTIPLY, d, g.const(0.35))
        g.output(s, g.node(KIND_SUM2, gain, e))
    x = torch.rand((n_slots, n), dtype=torch.float32, device="cuda")
    out = torch.empty((n_slots, n), dtype=torch.float32, device="cuda")
    offs = np.arange(n_slots + 1, dtype=np.uint64) * n
    t = timed_fill(r, out, n_slots, n, 0, x.data_ptr(), offs)
    alg = 8.0 * n_slots * n
    return {"case": "K2/K3 fused elementwise+Delay", "slots": n_slots, "samples": n, "ms": t["total_ms"], "interp_ms": t["interp_ms"],
            "algorithmic_GBs": alg / t["interp_ms"] / 1e6, "peak_GBs": HBM, "frac": alg / t["interp_ms"] / 1e6 / HBM,
            "note": "total_ms includes the D2D ingest of the input rows into the history buffers"}


def case_pure_elementwise(n_slots=64, n=1 << 22):
    """out_s = (in_s * 0.5 + 0.25) min in_s : no Delay, single stage; 8 B/sample."""
    from libfriendship_b200 import KIND_MINIMUM
    r = B200Renderer()
    g = GraphBuilder(r)
    for s in range(n_slots):
        x = g.input(s)
        a = g.node(KIND_MULTIPLY, x, g.const(0.5))
        b = g.node(KIND_SUM2, a, g.const(0.25))
        g.output(s, g.node(KIND_MINIMUM, b, x))
    x = torch.rand((n_slots, n), dtype=torch.float32, device="cuda")
    out = torch.empty((n_slots, n), dtype=torch.float32, device="cuda")
    offs = np.arange(n_slots + 1, dtype=np.uint64) * n
    t = timed_fill(r, out, n_slots, n, 0, x.data_ptr(), offs)
    alg = 8.0 * n_slots * n
    return {"case": "K2 pure elementwise (3 nodes/slot)", "slots": n_slots, "samples": n, "ms": t["total_ms"], "interp_ms": t["interp_ms"],
            "algorithmic_GBs": alg / t["interp_ms"] / 1e6, "peak_GBs": HBM, "frac": alg / t["interp_ms"] / 1e6 / HBM}


def case_refbank(n_partials=1024, n=1 << 19, flags=0):
    """An additive bank in the reference's own vocabulary: n_partials external inputs, Multiply(in_p, C(amp_p)), left Sum2
    chain — ONE stage of 3 x n_partials instructions, which the stage JIT compiles as one loop (csrc/jit.cc).
    Algorithmic: 4 B x (n_partials reads + 1 write) per sample."""
    from workloads.banks import build_partial_sum_graph
    r = B200Renderer(flags=flags)
    build_partial_sum_graph(r, [1.0 / (p + 1) for p in range(n_partials)])
    x = torch.rand((n_partials, n), dtype=torch.float32, device="cuda")
    out = torch.empty((1, n), dtype=torch.float32, device="cuda")
    offs = np.arange(n_partials + 1, dtype=np.uint64) * n
    t0 = time.perf_counter()
    r.fill_buffer_device(out.data_ptr(), 1, n, 0, x.data_ptr(), offs)
    r.sync()
    first = time.perf_counter() - t0
    t = timed_fill(r, out, 1, n, 0, x.data_ptr(), offs, reps=3)
    alg = 4.0 * (n_partials + 1) * n
    st = r.stats()
    return {"case": "reference-vocabulary additive bank (Multiply + Sum2 chain over external inputs)", "partials": n_partials, "samples": n,
            "first_call_s": first, "ms": t["total_ms"], "stage_ms": t["interp_ms"], "algorithmic_GBs": alg / t["interp_ms"] / 1e6, "peak_GBs": HBM,
            "frac": alg / t["interp_ms"] / 1e6 / HBM, "jit_launches": st["jit_launches"], "interp_launches": st["interp_launches"],
            "code_instructions": r.jit_code_instructions(1, 0)}


def case_cfg3(n_voices=4096, n=480000, flags=0, osc_anchor=0):
    """cfg3: per voice 1-partial oscillator -> biquad -> feedback delay, mixed to one slot.  K4 algorithmic traffic
    (SURVEY.md §8d): 8 B per voice-sample for the fused biquad -> comb chain (read x, write y); the two separate kernels
    (FLAG_NO_CHAIN_FUSION) move 16 B."""
    from workloads.banks import detuned_bank
    from workloads.filters import build_cfg3_graph
    r = B200Renderer(flags=flags, osc_anchor=osc_anchor)
    bank, _ = detuned_bank(n_voices, 1, seed=5)
    build_cfg3_graph(r, n_voices, excitation="osc", bank=bank, mix_to_one=True)
    out = torch.empty((1, n), dtype=torch.float32, device="cuda")
    t = timed_fill(r, out, 1, n, 0, reps=3)
    alg = 8.0 * n_voices * n
    fused = r.stats()["chain_launches"] > 0
    exciter_fused = fused and r.stats()["osc_launches"] == 0     # the chain kernel evaluates the oscillators itself: writes only
    moved = alg / 2 if exciter_fused else alg if fused else 2 * alg
    return {"case": "cfg3 osc->DirectForm->FbDelay->mix" + (" (exciters inside the chain kernel)" if exciter_fused else "" if fused else " (chain fusion off)"), "voices": n_voices, "samples": n,
            "ms": t["total_ms"], "scan_ms": t["scan_ms"], "osc_ms": t["osc_ms"], "fold_ms": t["interp_ms"],
            "K4_algorithmic_GBs": alg / t["scan_ms"] / 1e6, "peak_GBs": HBM, "K4_frac": alg / t["scan_ms"] / 1e6 / HBM,
            "K4_kernel_bytes_GBs": moved / t["scan_ms"] / 1e6, "osc_write_GBs": (4.0 * n_voices * n / t["osc_ms"] / 1e6) if t["osc_ms"] > 0 else None,
            "fold_read_GBs": 4.0 * n_voices * n / t["interp_ms"] / 1e6, "voice_samples_per_s": n_voices * n / (t["total_ms"] * 1e-3)}


def case_cfg2(anchor=0):
    """cfg2: 1,024 harmonic partials x 1 voice, 48 kHz x 10 s (one voice: parallelism only along time)."""
    from workloads.banks import harmonic_bank
    from libfriendship_b200 import KIND_OSCBANK
    n = 480000
    r = B200Renderer(osc_anchor=anchor)
    r.define_oscbank(5, **harmonic_bank(1024))
    r.on_add_node(1, KIND_OSCBANK, 5)
    r.on_add_edge((1, 0, 0, 0))
    out = torch.empty((1, n), dtype=torch.float32, device="cuda")
    t = timed_fill(r, out, 1, n, 0, reps=5)
    ps = 1024 * n
    return {"case": "cfg2 1,024 partials x 1 voice x 10 s", "anchor": anchor or 128, "ms": t["total_ms"], "osc_ms": t["osc_ms"],
            "partial_samples_per_s": ps / (t["total_ms"] * 1e-3), "realtime_factor": 10.0 / (t["total_ms"] * 1e-3)}


def case_cfg1():
    """cfg1: 440 Hz sine through Multiply/Sum/Delay (+ Min/Mod/Div side chain), 48 kHz x 1 s, host in/out through
    frb_fill_buffer: one call, and 94 x 512-sample streaming calls.  Latency-bound: reported as us per block."""
    from workloads.graphs import build_cfg1_graph, cfg1_input
    n = 48000
    x = cfg1_input(n)
    r = B200Renderer()
    build_cfg1_graph(r)
    for _ in range(3):
        r.fill_buffer(2, n, 0, [x])
    t0 = time.perf_counter()
    for _ in range(10):
        r.fill_buffer(2, n, 0, [x])
    whole_us = (time.perf_counter() - t0) / 10 * 1e6
    out = np.zeros((2, 512), dtype=np.float32)
    def stream():
        for s0 in range(0, n, 512):
            m = min(512, n - s0)
            r.fill_buffer(2, m, s0, [x[s0:s0 + m]], out=out[:, :m] if m == 512 else None)
    stream()
    t0 = time.perf_counter()
    for _ in range(3):
        stream()
    blk_us = (time.perf_counter() - t0) / 3 / 94 * 1e6
    return {"case": "cfg1 render_prim-style graph, 48 kHz x 1 s, host buffers", "one_call_us": whole_us,
            "us_per_512_sample_block": blk_us, "realtime_factor_streaming": (512 / 48000 * 1e6) / blk_us,
            "jit_launches": r.stats()["jit_launches"]}


if __name__ == "__main__":
    which = sys.argv[1:] or ["pure", "elementwise", "cfg3"]
    for w in which:
        fn = {"pure": case_pure_elementwise, "pure_continuing": case_pure_continuing, "elementwise": case_elementwise, "cfg3": case_cfg3, "cfg3_unfused": lambda: case_cfg3(flags=8), "cfg3_ring": lambda: case_cfg3(flags=16), "cfg3_L64": lambda: case_cfg3(osc_anchor=64), "cfg3_L32": lambda: case_cfg3(osc_anchor=32),
          "cfg3_L128": lambda: case_cfg3(osc_anchor=128), "cfg3_L256": lambda: case_cfg3(osc_anchor=256), "cfg1": case_cfg1, "cfg2": case_cfg2, "refbank": case_refbank, "refbank_nojit": lambda: case_refbank(flags=2), "refbank256": lambda: case_refbank(256, 1 << 21), "cfg2_64": lambda: case_cfg2(64), "cfg2_32": lambda: case_cfg2(32)}[w]
        t0 = time.time()
        try:
            res = fn()
        except Exception as e:
            res = {"case": w, "error": repr(e)}
        res["wall_s"] = time.time() - t0
        print(json.dumps(res), flush=True)
