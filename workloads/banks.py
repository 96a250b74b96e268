"""Synthetic oscillator banks of the BASELINE.json shapes, at any scale."""
import numpy as np

from .kinds import KIND_DELAY, KIND_MULTIPLY, KIND_OSCBANK, KIND_SUM2
from .graphs import GraphBuilder


def harmonic_bank(n_partials, sr=48000.0, f0=20.0):
    """cfg2: partial p = 1..n: f = f0*p, amp 1/p, phase 0, attack 48*(1 + p mod 7) samples, tau = sr*(0.2 + 2/p)."""
    p = np.arange(1, n_partials + 1, dtype=np.float64)
    freq = f0 * p
    amp = (1.0 / p).astype(np.float32)
    phase = np.zeros(n_partials, dtype=np.float32)
    attack = (48.0 * (1 + (p.astype(np.int64) % 7))).astype(np.float32)
    tau = (sr * (0.2 + 2.0 / p)).astype(np.float32)
    return dict(sample_rate=sr, voice_offsets=np.array([0, n_partials], dtype=np.uint64), freq_hz=freq, amp=amp,
                phase=phase, attack=attack, tau=tau)


def detuned_bank(n_voices, n_partials, sr=48000.0, seed=1, voices=None):
    """cfg4: voice v, partial k = 1..n: f = f0_v*k*(1+delta) wrapped below Nyquist, f0_v = 55*2^(v/12),
    delta ~ U(-0.002, 0.002) from PCG64(seed), amp 1/k, envelopes as cfg2.  `voices` selects a shard of voices
    (the random stream is drawn for all voices so shards are consistent across ranks)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    k = np.arange(1, n_partials + 1, dtype=np.float64)
    sel = list(range(n_voices)) if voices is None else list(voices)
    freqs, amps, phases, attacks, taus = [], [], [], [], []
    for v in range(n_voices):
        delta = rng.uniform(-0.002, 0.002, n_partials)
        if v not in sel:
            continue
        f0 = 55.0 * 2.0 ** (v / 12.0)
        f = f0 * k * (1.0 + delta)
        nyq = sr / 2.0
        f = np.where(f >= nyq, np.mod(f, nyq * 0.98) + 20.0, f)     # wrap below Nyquist
        freqs.append(f)
        amps.append((1.0 / k).astype(np.float32))
        phases.append(np.zeros(n_partials, dtype=np.float32))
        attacks.append((48.0 * (1 + (k.astype(np.int64) % 7))).astype(np.float32))
        taus.append((sr * (0.2 + 2.0 / k)).astype(np.float32))
    nv = len(sel)
    return dict(sample_rate=sr, voice_offsets=(np.arange(nv + 1, dtype=np.uint64) * n_partials),
                freq_hz=np.concatenate(freqs), amp=np.concatenate(amps), phase=np.concatenate(phases),
                attack=np.concatenate(attacks), tau=np.concatenate(taus)), sel


def full_scale(bank):
    """'Full scale' of a bank for the 1e-5 tolerance = sum |amp| of the loudest voice."""
    vo = bank["voice_offsets"].astype(np.int64)
    a = np.abs(bank["amp"].astype(np.float64))
    return max(float(a[vo[i]:vo[i + 1]].sum()) for i in range(len(vo) - 1))


def build_voice_mix_graph(r, bank, voice_ids, key=7, delay0=4800.0, delay_step=37.0, mix=0.3):
    """cfg4 graph: OscBank -> per voice  x_v + mix * Delay(x_v, delay0 + delay_step*v)  -> Sum2 chain -> out0.
    `voice_ids` are the global voice numbers of the bank's lanes (delays depend on the global number)."""
    r.define_oscbank(key, **bank)
    g = GraphBuilder(r)
    osc_h = g.next
    g.next += 1
    r.on_add_node(osc_h, KIND_OSCBANK, key)
    total = None
    for lane, v in enumerate(voice_ids):
        x = (osc_h, lane)
        d = g.node(KIND_DELAY, x, g.const(delay0 + delay_step * v))
        wet = g.node(KIND_MULTIPLY, d, g.const(mix))
        voice = g.node(KIND_SUM2, x, wet)
        total = voice if total is None else g.node(KIND_SUM2, total, voice)
    g.output(0, total)
    return g


def partial_signals(bank, n_times, idx=0, voice=0):
    """One voice of a bank spelled out in the reference's own vocabulary: the only way a time-varying signal enters the
    reference is as an external input, so every partial's unit-amplitude signal
        min(t/attack_p, 1) * exp(-t/tau_p) * sin(2*pi*f_p*t/sr + phase_p)          (include/friendship_b200.h)
    is evaluated here in fp64, rounded once to f32, and becomes input row p."""
    vo = bank["voice_offsets"].astype(np.int64)
    lo, hi = int(vo[voice]), int(vo[voice + 1])
    t = np.arange(idx, idx + n_times, dtype=np.float64)[None, :]
    f = bank["freq_hz"][lo:hi].astype(np.float64)[:, None]
    ph = bank["phase"][lo:hi].astype(np.float64)[:, None]
    att = bank["attack"][lo:hi].astype(np.float64)[:, None]
    tau = bank["tau"][lo:hi].astype(np.float64)[:, None]
    env = np.where(att > 0, np.minimum(t / np.where(att > 0, att, 1.0), 1.0), 1.0)
    env = env * np.where((tau > 0) & np.isfinite(tau), np.exp(-t / np.where(tau > 0, tau, 1.0)), 1.0)
    return (env * np.sin(2.0 * np.pi * f * t / bank["sample_rate"] + ph)).astype(np.float32)


def build_partial_sum_graph(r, amps):
    """out0 = (((in0*a0 + in1*a1) + in2*a2) + ...): Multiply(in_p, C(amp_p)) terms in a left Sum2 chain — the additive
    bank as a graph of the reference's seven primitives (reference.rs:221-234)."""
    from .kinds import KIND_MULTIPLY, KIND_SUM2
    g = GraphBuilder(r)
    total = None
    for p, a in enumerate(amps):
        term = g.node(KIND_MULTIPLY, g.input(p), g.const(a))
        total = term if total is None else g.node(KIND_SUM2, total, term)
    g.output(0, total)
    return g
