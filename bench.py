#!/usr/bin/env python
"""bench.py — rendered partial-samples/s on BASELINE.json configs[3]:
65,536 detuned partials x 64 voices with envelopes and a per-voice delay effect, 48 kHz x 10 s,
sharded by voice over N GPUs with one NCCL reduce of the mixed output.

One "step" = one full 10 s render of the whole graph (2.013e12 partial-samples over all voices).
  value : device-resident throughput (output block stays in HBM; reduce included), CUDA-timed, max over ranks
  e2e   : the same through the host-facing calls, every step: frb_define_oscbank from PINNED HOST parameter arrays (the
          synthesis graph's only input: host->device copy + device-side regroup/setup) + frb_fill_buffer into a host
          buffer (device->host copy of the mixed block on rank 0), all inside the timed region
  roofline : dominant kernel: K1T osc_tc_kernel (the bank as a matrix product, tcgen05) against the measured dense tensor peak;
             the resonator kernel K1 against the FP32 FMA pipe when FRB_OSC_GEMM=0 keeps the bank on it
  cpu_baseline : the CPU oracle (restatement of the reference's per-sample renderer) on the box's host cores
  parity   : after the timed regions (outside them), rank 0's rendered block — device path and host path — against the
             fp64 oracle evaluated on the FULL 64 x 65,536 bank in five windows; the run fails (rc 3) above 1e-5 of full scale
  extra    : (N = 1) the other BASELINE.json configurations measured in the same process: cfg1 latency, cfg2, cfg3
`--impl reference` times that CPU implementation alone on the same metric/config (bounded sample).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np

SR = 48000.0
N_VOICES = 64
N_PARTIALS = 65536
N_SAMPLES = 480000
FMA_SLOTS_PER_PARTIAL_SAMPLE = 6      # rotation 4 + accumulate 1 + envelope 1 (algorithmic)
EXECUTED_OPS_PER_PARTIAL_SAMPLE = 4   # what K1 issues: 3 FFMA + 1 FADD
BANK_KEY = 7


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)
    except Exception:
        return {}


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, device):
        super().__init__(daemon=True)
        self.device, self.stop_flag, self.samples, self.reasons, self.max_mhz = device, False, [], set(), None

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.device)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            names = {
                getattr(pynvml, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
                getattr(pynvml, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                getattr(pynvml, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                getattr(pynvml, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            }
            while not self.stop_flag:
                self.samples.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                try:
                    bits = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    bits = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if bits & bit:
                        self.reasons.add(name)
                time.sleep(0.05)
        except Exception as e:      # pragma: no cover
            self.reasons.add(f"sampler_error:{type(e).__name__}")

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def cpu_baseline(n_threads=None, budget_s=12.0):
    """The CPU oracle (reference-style f32 per-sample evaluation, as RefRenderer would do it) on a bounded sample of
    the same workload: voice 0, its first `p` partials, `n` samples, through the same per-voice delay/mix graph."""
    from workloads.banks import build_voice_mix_graph, detuned_bank
    from oracle.binding import OracleRenderer
    n_threads = n_threads or os.cpu_count() or 1
    p, n = 4096, 2400
    bank, ids = detuned_bank(1, p)
    r = OracleRenderer(ext_mode="f32")
    build_voice_mix_graph(r, bank, ids)
    r.fill_buffer(1, 0, 0)
    # calibrate on a short run, then size the sample for ~budget_s of CPU work
    _, sec = r.fill_buffer_mt(1, n, 0, n_threads)
    per = sec / (p * n * 2)      # the Delay re-evaluates the voice at t-d: 2 evaluations per output sample
    n2 = int(min(N_SAMPLES, max(n, budget_s / max(per * p * 2, 1e-12))))   # up to the configuration's full 10 s
    _, sec = r.fill_buffer_mt(1, n2, 4800, n_threads)     # start after the delay so both taps are live
    value = p * n2 / sec
    _, sec1 = r.fill_buffer_mt(1, max(n2 // n_threads, 64), 4800, 1)
    value1 = p * max(n2 // n_threads, 64) / sec1
    return {"value": value, "unit": "partial-samples/s", "cores": n_threads, "kind": "port",
            "sample": f"1 voice x {p} partials x {n2} samples of the cfg4 graph (per-voice delay+mix), oracle f32 per-sample "
                      f"evaluation (sinf/expf per partial-sample, as RefRenderer would), time axis split over {n_threads} threads",
            "value_1thread": value1}


PARITY_TOL = 1e-5          # of full scale (north_star: max abs error <= 1e-5 of full scale against an fp64 oracle)


def parity_windows(n_samples, n_voices, delay0=4800, delay_step=37, width=16):
    """Where the rendered block is compared with the oracle: the first samples, either end of the per-voice delay taps
    (voice v's wet path opens at t = delay0 + delay_step*v), the middle of the render, its last samples."""
    last_tap = delay0 + delay_step * (n_voices - 1)
    starts = [0, delay0 - width // 2, last_tap - width // 2, n_samples // 2, n_samples - 2 * width]
    wins = []
    for s0 in starts:
        s0 = max(0, min(int(s0), n_samples - 1))
        n = min(2 * width if s0 == n_samples - 2 * width else width, n_samples - s0)
        if n > 0 and (s0, n) not in wins:
            wins.append((s0, n))
    return wins


def parity_check(blocks, n_voices, n_partials, n_samples, n_threads=None):
    """blocks: {name: host array [1 x n_samples]} rendered by the GPU path(s) at the bench's own size.  The oracle
    (fp64 closed form per partial, oracle/ref_renderer.hpp osc_value; the graph around it evaluated the reference's way)
    renders the same windows from the full bank definition — every voice, every partial, both delay taps."""
    from oracle.binding import OracleRenderer
    from workloads.banks import build_voice_mix_graph, detuned_bank, full_scale
    n_threads = n_threads or os.cpu_count() or 1
    t0 = time.perf_counter()
    bank, ids = detuned_bank(n_voices, n_partials)
    orc = OracleRenderer(ext_mode="fp64")
    build_voice_mix_graph(orc, bank, ids, key=BANK_KEY)
    orc.fill_buffer(1, 0, 0)
    fs = full_scale(bank) * n_voices * 1.3          # every voice at full scale, dry + 0.3 wet
    wins = parity_windows(n_samples, n_voices)
    worst = {k: 0.0 for k in blocks}
    for s0, n in wins:
        want, _ = orc.fill_buffer_mt(1, n, s0, n_threads)
        for k, blk in blocks.items():
            err = float(np.abs(blk[0, s0:s0 + n].astype(np.float64) - want[0].astype(np.float64)).max())
            worst[k] = max(worst[k], err / fs)
    mx = max(worst.values())
    return {"max_err_of_full_scale": mx, "per_path": worst, "tol": PARITY_TOL, "ok": bool(mx <= PARITY_TOL),
            "windows": [[int(a), int(b)] for a, b in wins], "full_scale": fs,
            "oracle": f"fp64, full bank {n_voices} x {n_partials}, {n_threads} host threads, {time.perf_counter() - t0:.1f} s",
            "checked": "rank 0's block after the exchange: device-resident path and host (e2e) path"}


def extra_configs(peak_fma, hbm_gbs):
    """The other BASELINE.json configurations, measured after the main regions in this process (N = 1): what the
    driver would otherwise only know from builder-side files under profiles/."""
    from tools.bench_kernels import case_cfg1, case_cfg2, case_cfg3
    ex = {}
    try:
        c = case_cfg1()
        ex["cfg1"] = {"workload": "render_prim-style graph, 440 Hz sine input, 2 slots, 48 kHz x 1 s, host buffers through frb_fill_buffer (Python ctypes binding)",
                      "us_per_512_sample_call": c["us_per_512_sample_block"], "us_one_call_48000_samples": c["one_call_us"],
                      "realtime_factor_streaming": c["realtime_factor_streaming"], "bound": "latency"}
    except Exception as e:      # pragma: no cover
        ex["cfg1"] = {"error": repr(e)}
    try:
        # the same graph and call pattern from plain C (tools/microbench/cfg1_latency.c, built by __graft_entry__.build()):
        # what one frb_fill_buffer call costs without a Python binding in the way
        import subprocess
        exe = os.path.join(ROOT, "build", "bin", "cfg1_latency")
        r = subprocess.run([exe, "512", "2000"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
        c = json.loads([l for l in r.stdout.strip().splitlines() if '"us_per_call_median"' in l][-1])
        ex["cfg1"]["c_abi"] = {"us_per_512_sample_call_median": c["us_per_call_median"], "us_p10": c["us_p10"], "us_p90": c["us_p90"],
                               "realtime_factor": c["realtime_factor"], "calls": c["calls"], "program": "tools/microbench/cfg1_latency.c"}
    except Exception as e:      # pragma: no cover
        ex.setdefault("cfg1", {})["c_abi"] = {"error": repr(e)}
    try:
        c = case_cfg2()
        ps = 1024 * 480000
        ex["cfg2"] = {"workload": "1,024 harmonic partials x 1 voice, 48 kHz x 10 s, device resident", "ms": c["ms"], "osc_ms": c["osc_ms"],
                      "partial_samples_per_s": c["partial_samples_per_s"], "bound": "fp32_fma",
                      "frac_issued_ops": ps * EXECUTED_OPS_PER_PARTIAL_SAMPLE / (c["osc_ms"] * 1e-3) / peak_fma}
    except Exception as e:      # pragma: no cover
        ex["cfg2"] = {"error": repr(e)}
    try:
        c = case_cfg3()
        ex["cfg3"] = {"workload": "4,096 voices x (one-partial oscillator -> biquad -> feedback delay) -> mix, 48 kHz x 10 s, device resident",
                      "ms": c["ms"], "chain_ms": c["scan_ms"], "fold_ms": c["fold_ms"], "voice_samples_per_s": c["voice_samples_per_s"],
                      "bound": "hbm", "frac_8B_definition": c["K4_frac"], "frac_real_bytes": c["K4_kernel_bytes_GBs"] / hbm_gbs,
                      "note": "frac_8B_definition = SURVEY.md §8d's 8 B per voice-sample / chain kernel time / measured HBM peak; with the exciters "
                              "evaluated inside the chain kernel it moves 4 B (frac_real_bytes) and is issue-bound, not HBM-bound"}
    except Exception as e:      # pragma: no cover
        ex["cfg3"] = {"error": repr(e)}
    return ex


def run_reference(args, rank, world, out):
    if rank != 0:
        return
    vals = []
    cb = None
    for i in range(args.warmup + args.steps):
        cb = cpu_baseline(budget_s=6.0)
        if i >= args.warmup:
            vals.append(cb["value"])
    v = float(np.mean(vals))
    cb["value"] = v
    line = {
        "impl": "reference", "metric": "rendered partial-samples/sec", "value": v, "unit": "partial-samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * (N_VOICES * N_PARTIALS * N_SAMPLES) / v, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": cb,
        "e2e": {"value": v, "unit": "partial-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference (Rust 2017-nightly + LLVM 3.8) cannot be built here; this is the C++ restatement of its "
                "per-sample renderer (oracle/), all host threads, bounded sample extrapolated linearly",
    }
    print(json.dumps(line), file=out, flush=True)


def workload_config(n_gpus, exchange="NCCL reduce"):
    return {"workload": "cfg4: 65,536 detuned partials x 64 voices, envelopes + per-voice Delay/mix, 48 kHz x 10 s "
                        "(BASELINE.json configs[3])",
            "voices": N_VOICES, "partials_per_voice": N_PARTIALS, "samples": N_SAMPLES, "sample_rate": SR,
            "partial_samples_per_step": N_VOICES * N_PARTIALS * N_SAMPLES,
            "sharding": f"voices round-robin over {n_gpus} GPU(s), one {exchange} exchange of the [1 x 480000] mix per step",
            "cache": "compute-bound; per-step parameter stream 201 MB/GPU-shard-of-64 > 126 MB L2, re-read by every 16,384-sample tile of a voice",
            "inputs": "synthesis: the graph's only input is the bank's parameter arrays (24 B per partial, host). "
                      "`value`: uploaded once, resident in HBM when the timed region starts. `e2e`: re-uploaded from pinned "
                      "host memory through frb_define_oscbank every step, inside the timed region"}


def _claim_stdout():
    """The driver reads ONE JSON line from stdout.  Libraries write there too (NCCL prints its version line to stdout
    when NCCL_DEBUG is set in the environment), so file descriptor 1 is pointed at stderr for the whole run and the
    JSON line goes to a private duplicate of the original stdout."""
    sys.stdout.flush()
    out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return out


def main():
    out = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=["cfg4", "cfg5"],
                    help="cfg4 = the bench (BASELINE.json configs[3]); cfg5 = the offline large render (configs[4]: 2^20 partials x "
                         "256 voices, 192 kHz x 60 s, meant for 8 GPUs under torchrun), one render through the streaming path with "
                         "oracle windows on the full bank — tools/render_cfg5.py, its own JSON line")
    ap.add_argument("--voices", type=int, default=N_VOICES, help="debug: shrink the workload (invalidates the number)")
    ap.add_argument("--partials", type=int, default=N_PARTIALS)
    ap.add_argument("--samples", type=int, default=N_SAMPLES)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle windows (debug only)")
    ap.add_argument("--no-extra", action="store_true", help="skip the cfg1/cfg2/cfg3 measurements (N = 1)")
    ap.add_argument("--osc-anchor", type=int, default=0, help="debug: K1 segment length / re-anchor interval")
    ap.add_argument("--exchange", default="nccl", choices=["nccl", "p2p"],
                    help="N > 1: nccl = one NCCL reduce of the mix blocks (default); p2p = K5, the stage kernel stores its "
                         "block into rank 0's slab over NVLink (CUDA IPC) and rank 0 sums the rows in rank order")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world, out)
        return
    if args.workload == "cfg5":
        os.dup2(out.fileno(), 1)            # that tool prints its line itself
        sys.stdout = out
        from tools import render_cfg5
        render_cfg5.main()
        return

    import torch
    import torch.distributed as dist
    from workloads.banks import build_voice_mix_graph, detuned_bank
    from libfriendship_b200.sharded import ShardedRenderer

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    n_voices, n_partials, n_samples = args.voices, args.partials, args.samples
    sr = ShardedRenderer(rank=rank, world_size=world, device=local_rank, osc_anchor=args.osc_anchor, exchange=args.exchange)
    my_voices = sr.voices_of_rank(n_voices)
    bank, ids = detuned_bank(n_voices, n_partials, voices=my_voices)
    build_voice_mix_graph(sr.r, bank, ids, key=BANK_KEY)
    total_ps = n_voices * n_partials * n_samples

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    rstream = sr.cuda_stream()

    def timed(fn, steps):
        # CUDA events on the stream the kernels are launched on (the renderer's own stream); every step ends with a
        # device sync (renderer stream, and the collective's stream for N > 1), so the bracket covers all the work.
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record(rstream)
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        ev1.record(rstream)
        ev1.synchronize()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        barrier()
        ms = max(ev0.elapsed_time(ev1), wall * 1e3)      # wall clock of the synchronized loop as a cross-check
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # e2e inputs: this rank's parameter arrays in pinned host memory
    pinned = {k: (torch.from_numpy(np.ascontiguousarray(v)).pin_memory().numpy() if isinstance(v, np.ndarray) else v)
              for k, v in bank.items()}
    h2d_per_step = sum(v.nbytes for v in pinned.values() if isinstance(v, np.ndarray))

    step_dev = lambda: sr.fill_buffer_device(1, n_samples, 0)

    def step_e2e():
        sr.r.define_oscbank(BANK_KEY, **pinned)     # host -> device: the step's inputs
        return sr.fill_buffer(1, n_samples, 0)      # render + reduce + device -> host
    for _ in range(args.warmup):
        step_dev()
    s0 = sr.r.stats()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_dev = timed(step_dev, args.steps)
    s1 = sr.r.stats()
    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    s2 = sr.r.stats()
    ms_e2e = timed(step_e2e, args.steps)
    s3 = sr.r.stats()
    sampler.stop_flag = True
    h2d = torch.tensor([h2d_per_step], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(h2d, op=dist.ReduceOp.SUM)      # all ranks' uploads

    # dominant kernel: K1, timed live with CUDA events on the renderer's stream (frb_set_profiling)
    sr.r.set_profiling(True)
    step_dev()
    tim = sr.r.timing()
    sr.r.set_profiling(False)
    # what was rendered, for the parity check below (outside every timed region)
    dev_block = step_dev()
    sr.r.sync()                                  # the block is complete on the renderer's stream, which torch's .cpu() does not know
    torch.cuda.synchronize()
    host_dev = dev_block.cpu().numpy().copy() if rank == 0 else None
    host_e2e = step_e2e()
    host_e2e = host_e2e.copy() if rank == 0 else None
    osc_ms = torch.tensor([tim["osc_ms"]], dtype=torch.float64, device="cuda")
    tot_ms = torch.tensor([tim["total_ms"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(osc_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot_ms, op=dist.ReduceOp.MAX)

    parity_ok = True
    if rank == 0:
        peaks = measured_peaks()
        sm_max = peaks.get("sm_max_mhz", 1965.0)
        n_sm = torch.cuda.get_device_properties(local_rank).multi_processor_count
        peak_fma = n_sm * 128 * sm_max * 1e6                  # FP32 FMA lanes/s (derived; microbench measured 97% of it)
        ps_per_gpu = len(my_voices) * n_partials * n_samples  # rank 0's share (the largest shard)
        osc_s = float(osc_ms.item()) * 1e-3
        n_osc_launches = max(1, (s1["osc_launches"] - s0["osc_launches"]) // args.steps)
        achieved = ps_per_gpu * EXECUTED_OPS_PER_PARTIAL_SAMPLE * 2 / osc_s / 1e12     # flop K1 issues
        achieved_survey = ps_per_gpu * FMA_SLOTS_PER_PARTIAL_SAMPLE * 2 / osc_s / 1e12
        peak_tf = peak_fma * 2 / 1e12
        value = total_ps * args.steps / (ms_dev * 1e-3)
        e2e_v = total_ps * args.steps / (ms_e2e * 1e-3)
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "k1_traffic.json")) as f:
                tj = json.load(f)
            traffic = tj["dram_bytes_read_per_launch"] + tj["dram_bytes_write_per_launch"]
        except Exception:
            pass
        tensor_launches = int(s1.get("osc_tensor_launches", 0) - s0.get("osc_tensor_launches", 0))
        k1_ms = float(osc_ms.item())
        share = k1_ms / float(tot_ms.item())
        if tensor_launches > 0:
            # K1T / K1G (csrc/osc_tc.cuh, osc_gemm.cuh): the bank past its attack ramps is a matrix product on the tensor cores
            tile = 128 * 128
            ramp_end = 384                                          # longest attack ramp of the bench bank (48 * 7 = 336), in 128-sample blocks
            n_tiles = -(-n_samples // tile) - ramp_end // tile
            stages = -(-n_partials // 16)
            issued = len(my_voices) * n_tiles * stages * 6 * 2.0 * 128 * 128 * 16      # six M128 N128 K16 MMAs per 16 partials and tile
            algorithmic = ps_per_gpu * FMA_SLOTS_PER_PARTIAL_SAMPLE * 2                  # 6 slots = 12 flop per partial-sample
            peak_tensor = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1369.6))
            tensor_kernel = os.environ.get("FRB_OSC_GEMM", "1") not in ("2", "5")
            roofline = {
                "kernel": "osc_tc_kernel (K1T, tcgen05)" if tensor_kernel else "osc_gemm_kernel (K1G, mma.sync)",
                "bound": "tensor", "achieved": algorithmic / osc_s / 1e12, "peak": peak_tensor, "unit": "TFLOP/s",
                "frac": algorithmic / osc_s / 1e12 / peak_tensor,
                "frac_note": "SURVEY.md 8d's algorithmic 12 flop (6 FMA slots) per partial-sample over the K1 family's time, against the "
                             "measured dense 16-bit tensor peak.  The kernel ISSUES 12 flop per partial-sample of its tiles too "
                             "(fp16 hi/lo split: three M128 N128 K16 MMAs per 8 partials, 2 K flop per sample) — issued_frac counts the "
                             "tiles' padding as well.  What bounds the kernel is generating the operands (FMA / conversion pipes, "
                             "shared-memory stores), not the tensor pipe: DESIGN.md",
                "issued_tflops": issued / osc_s / 1e12, "issued_frac": issued / osc_s / 1e12 / peak_tensor,
                "traffic": None,
                "traffic_note": "DRAM traffic is the parameter records (48 B per partial and tile, L2 hits after the first tile of a voice) "
                                "and the output: under 1% of the HBM peak; no ncu capture of this kernel yet",
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (cuBLAS bf16 8192^3 back to back; fp16 runs at the same rate); "
                               f"burst {peaks.get('bf16_tflops')}",
                "algorithmic": "12 flop per partial-sample x partial-samples per step (SURVEY.md 8d)",
                "fp32_fma_equivalent_frac": ps_per_gpu * EXECUTED_OPS_PER_PARTIAL_SAMPLE / osc_s / peak_fma,
                "k1_family_launches_per_step": int(n_osc_launches), "tensor_launches_per_step": tensor_launches // max(1, args.steps),
                "k1_ms_per_step": k1_ms,
                "launches_note": "K1 family = the matrix-product kernel (one launch) + the resonator kernels of the attack-ramp region "
                                 "(main, ramp instance, plane reduce); time = CUDA events on the renderer's stream",
                "kernel_share_of_step": share,
            }
        else:
            roofline = {
                "kernel": "osc_kernel<16,false> (K1)", "bound": "fp32_fma", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved / peak_tf,
                "frac_note": "issued-op utilisation of the FP32 pipe: K1 issues 4 ops per partial-sample (3 FFMA + 1 FADD: the decay is "
                             "folded into the rotation's eigenvalues, the amplitude into the anchor).  SURVEY.md §8d's algorithmic "
                             "definition counts 6 slots per partial-sample; by it the same time gives frac_survey_definition",
                "frac_survey_definition": achieved_survey / peak_tf, "achieved_survey_definition": achieved_survey,
                "traffic": traffic,
                "traffic_note": "DRAM bytes of one K1 launch (ncu --set full, profiles/k1_traffic.json): parameter stream + "
                                "partial-range planes, per main-kernel launch (one 131,072-sample sub-block); irrelevant to the bound (0.2% of HBM peak)",
                "peak_source": f"derived {n_sm} SM x 128 lanes x 2 x sm_max_mhz {sm_max} (MEASURED_PEAKS.json); "
                               "tools/microbench/fma_peak.cu measured 3.60e13 FMA/s = 97% of it on this pool",
                "algorithmic": "issued: 4 FP32 ops (8 flop) per partial-sample x partial-samples per step; SURVEY.md §8d: 6 slots (12 flop)",
                "executed_frac": ps_per_gpu * EXECUTED_OPS_PER_PARTIAL_SAMPLE / osc_s / peak_fma,
                "k1_family_launches_per_step": int(n_osc_launches), "k1_ms_per_step": k1_ms,
                "launches_note": "K1 family = main kernel per sub-block + its plane reduce + one attack-ramp kernel; "
                                 "achieved = algorithmic flop of the step / summed K1 time (CUDA events on the renderer's stream)",
                "kernel_share_of_step": share,
            }
        line = {
            "metric": "rendered partial-samples/sec", "value": value, "unit": "partial-samples/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(world, "NCCL reduce" if args.exchange == "nccl" else "P2P-store + rank-ordered sum (K5)"),
            "e2e": {"value": e2e_v, "unit": "partial-samples/s", "h2d_bytes_per_step": int(h2d.item()),
                    "d2h_bytes_per_step": 4 * n_samples, "ms_per_step": ms_e2e / args.steps,
                    "gpu_launches": int(s3["kernel_launches"] - s2["kernel_launches"]),
                    "calls": "per step: frb_define_oscbank(pinned host arrays) + frb_fill_buffer(host out)"},
            "gpu_launches": int(s1["kernel_launches"] - s0["kernel_launches"]),
            "clocks": sampler.result(),
            "roofline": roofline,
        }
        if tensor_launches > 0:
            line["dtype"] = "f32 (operands split into 2 x fp16 = 22 bits, products exact, fp32 accumulation; parity below)"
        if not (args.voices == N_VOICES and args.partials == N_PARTIALS and args.samples == N_SAMPLES):
            line["config"]["workload"] = f"DEBUG reduced workload {n_voices}x{n_partials}x{n_samples}: not a valid bench number"
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline()
        if not args.no_parity:
            line["parity"] = parity_check({"device": host_dev, "host_e2e": host_e2e}, n_voices, n_partials, n_samples)
            parity_ok = line["parity"]["ok"]
        if world == 1 and not args.no_extra:
            line["extra"] = extra_configs(peak_fma, peaks.get("hbm_gbs", 6535.7))
        print(json.dumps(line), file=out, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if not parity_ok:
        sys.stderr.write("bench.py: PARITY FAILED: rendered block differs from the fp64 oracle by more than 1e-5 of full scale\n")
        sys.exit(3)


if __name__ == "__main__":
    main()
