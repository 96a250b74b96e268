"""GPU: K1 oscillator bank (extension node, no reference counterpart — PARITY UNPINNED by any reference test)
against the fp64 closed-form oracle and the reference-style f32 per-sample evaluation.
Tolerance (BASELINE.json north_star): max abs error <= 1e-5 of full scale, full scale = sum |amp| of a voice."""
import numpy as np
import pytest

from workloads.banks import build_voice_mix_graph, detuned_bank, full_scale, harmonic_bank
from oracle.binding import OracleRenderer
from replay import assert_same_bits

pytestmark = pytest.mark.gpu
TOL = 1e-5


def render_bank(cls, bank, n, idx=0, n_voices=1, **kw):
    from libfriendship_b200 import KIND_OSCBANK
    r = cls(**kw)
    r.define_oscbank(5, **bank)
    r.on_add_node(1, KIND_OSCBANK, 5)
    for v in range(n_voices):
        r.on_add_edge((1, 0, v, v))
    return r.fill_buffer(n_voices, n, idx)


@pytest.mark.parametrize("idx", [0, 1000, 480000 - 2048, 11_520_000 - 2048])
def test_cfg2_harmonic_bank_vs_fp64(idx):
    from libfriendship_b200 import B200Renderer
    bank = harmonic_bank(1024)
    n = 2048
    gpu = render_bank(B200Renderer, bank, n, idx)
    ref = render_bank(OracleRenderer, bank, n, idx)
    err = np.abs(gpu.astype(np.float64) - ref.astype(np.float64)).max()
    assert err <= TOL * full_scale(bank), (err, full_scale(bank))


def test_cfg2_vs_reference_style_f32():
    """'against both the reference f32 output and an fp64 oracle': the f32 per-sample evaluation the reference
    would do (sinf of an f32 phase, f32 accumulate) is itself only ~1e-5-accurate; we must be at least as close
    to it as its own distance to fp64 plus our tolerance."""
    from libfriendship_b200 import B200Renderer
    bank = harmonic_bank(256)
    n = 1024
    gpu = render_bank(B200Renderer, bank, n, 0)
    ref64 = render_bank(OracleRenderer, bank, n, 0)
    ref32 = render_bank(OracleRenderer, bank, n, 0, ext_mode="f32")
    fs = full_scale(bank)
    d_ref = np.abs(ref32.astype(np.float64) - ref64).max()
    d_gpu = np.abs(gpu.astype(np.float64) - ref32).max()
    assert d_gpu <= d_ref + TOL * fs


@pytest.mark.parametrize("f", [0.0, 1.0, 20.0, 440.0, 11999.0, 12000.0, 12001.0, 20000.0, 23999.0, 24000.0, 30000.0, -440.0])
def test_single_partial_all_frequency_classes(f):
    """One partial, full scale = its amplitude: exercises both resonator classes, DC, Nyquist and aliased inputs."""
    from libfriendship_b200 import B200Renderer
    bank = dict(sample_rate=48000.0, voice_offsets=np.array([0, 1], dtype=np.uint64), freq_hz=np.array([f]),
                amp=np.array([0.8], dtype=np.float32), phase=np.array([0.7], dtype=np.float32),
                attack=np.array([100.0], dtype=np.float32), tau=np.array([30000.0], dtype=np.float32))
    for idx in (0, 123456):
        n = 1024
        gpu = render_bank(B200Renderer, bank, n, idx)
        ref = render_bank(OracleRenderer, bank, n, idx)
        err = np.abs(gpu.astype(np.float64) - ref).max()
        assert err <= TOL * 0.8, (f, idx, err)


def test_no_decay_no_attack_and_silent_partials():
    from libfriendship_b200 import B200Renderer
    rng = np.random.RandomState(3)
    P = 37
    bank = dict(sample_rate=44100.0, voice_offsets=np.array([0, 10, 10, P], dtype=np.uint64),     # voice 1 is empty
                freq_hz=rng.uniform(10, 22000, P), amp=rng.uniform(0, 1, P).astype(np.float32),
                phase=rng.uniform(-3, 3, P).astype(np.float32), attack=np.zeros(P, dtype=np.float32),
                tau=np.full(P, np.inf, dtype=np.float32))
    bank["amp"][5] = 0.0
    bank["tau"][::3] = 0.0          # <= 0 also means "no decay"
    gpu = render_bank(B200Renderer, bank, 777, 50000, n_voices=3)
    ref = render_bank(OracleRenderer, bank, 777, 50000, n_voices=3)
    assert np.abs(gpu.astype(np.float64) - ref).max() <= TOL * full_scale(bank)
    assert (gpu[1] == 0).all()


def test_block_invariance_bit_exact():
    """Segments are anchored to absolute time, so one call and many ragged block calls give identical bits."""
    from libfriendship_b200 import B200Renderer, KIND_OSCBANK
    bank = harmonic_bank(300)
    whole = render_bank(B200Renderer, bank, 5000, 0)
    r = B200Renderer()
    r.define_oscbank(5, **bank)
    r.on_add_node(1, KIND_OSCBANK, 5)
    r.on_add_edge((1, 0, 0, 0))
    parts, idx = [], 0
    for n in (1, 127, 128, 129, 1000, 3, 3612):
        parts.append(r.fill_buffer(1, n, idx))
        idx += n
    assert_same_bits(np.concatenate(parts, axis=1), whole, "osc block invariance")


def test_cfg4_shape_small_voice_mix_graph():
    """cfg4 at test scale: detuned partials x voices, per-voice Delay + mix, summed to one slot; the delay/mix
    nodes are reference primitives, so given the same voice planes they are bit-exact; end to end the bound is the
    oscillator tolerance times the mix gain."""
    from libfriendship_b200 import B200Renderer
    n_voices, n_partials, n = 6, 200, 3000
    bank, ids = detuned_bank(n_voices, n_partials)
    outs = []
    for cls in (B200Renderer, OracleRenderer):
        r = cls()
        build_voice_mix_graph(r, bank, ids, delay0=480.0)
        outs.append(r.fill_buffer(1, n, 0))
    fs = full_scale(bank) * n_voices * 1.3
    assert np.abs(outs[0].astype(np.float64) - outs[1]).max() <= TOL * fs


def test_many_groups_split_planes_path():
    """A voice with enough partials to take the split-planes + reduce path (> 512 groups)."""
    from libfriendship_b200 import B200Renderer
    bank, _ = detuned_bank(2, 4500, seed=4)
    n = 600
    gpu = render_bank(B200Renderer, bank, n, 20000, n_voices=2)
    ref = render_bank(OracleRenderer, bank, n, 20000, n_voices=2)
    assert np.abs(gpu.astype(np.float64) - ref).max() <= TOL * full_scale(bank)


def test_device_regroup_ragged_voices_mixed_classes():
    """The definition pipeline groups partials by (voice, resonator class) on the device: ragged voices (empty, 1, a
    few, several 1,024-partial tiles), classes interleaved partial by partial, silent partials in between."""
    from libfriendship_b200 import B200Renderer
    rng = np.random.RandomState(11)
    lens = [0, 1, 7, 3000, 0, 1025, 33]
    vo = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    P = int(vo[-1])
    freq = rng.uniform(10, 23900, P)
    freq[::2] = rng.uniform(10, 11900, len(freq[::2]))        # class 0 / class 1 alternate
    freq[1::2] = rng.uniform(12100, 23900, len(freq[1::2]))
    amp = (rng.uniform(0.1, 1, P) / 64).astype(np.float32)
    amp[rng.randint(0, P, 40)] = 0.0
    bank = dict(sample_rate=48000.0, voice_offsets=vo, freq_hz=freq, amp=amp,
                phase=rng.uniform(-3, 3, P).astype(np.float32), attack=rng.uniform(0, 300, P).astype(np.float32),
                tau=rng.uniform(2000, 90000, P).astype(np.float32))
    nv = len(lens)
    for idx in (0, 77777):
        gpu = render_bank(B200Renderer, bank, 700, idx, n_voices=nv)
        ref = render_bank(OracleRenderer, bank, 700, idx, n_voices=nv)
        assert np.abs(gpu.astype(np.float64) - ref).max() <= TOL * full_scale(bank)
        assert (gpu[0] == 0).all() and (gpu[4] == 0).all()


def test_redefine_bank_replaces_parameters():
    """frb_define_oscbank on an existing key replaces the node's parameters (the per-render input of a synthesis
    graph): the next render uses the new bank, also retroactively through a Delay (pure function of absolute time)."""
    from libfriendship_b200 import B200Renderer
    bank_a, ids = detuned_bank(3, 300, seed=5)
    bank_b, _ = detuned_bank(3, 300, seed=6)
    bank_b["amp"] = (bank_b["amp"] * 0.5).astype(np.float32)
    g = B200Renderer()
    build_voice_mix_graph(g, bank_a, ids, delay0=100.0)
    first = g.fill_buffer(1, 2000, 0)
    g.define_oscbank(7, **bank_b)
    second = g.fill_buffer(1, 2000, 2000)                     # continues at the head: the delay taps reach back before 2000
    for bank, got, idx in ((bank_a, first, 0), (bank_b, second, 2000)):
        o = OracleRenderer()
        build_voice_mix_graph(o, bank, ids, delay0=100.0)
        ref = o.fill_buffer(1, 2000, idx)
        assert np.abs(got.astype(np.float64) - ref).max() <= TOL * full_scale(bank) * 3 * 1.3
    # same parameters again: identical bits to a fresh renderer (nothing of the old definition survives)
    g.define_oscbank(7, **bank_a)
    again = g.fill_buffer(1, 2000, 0)
    assert_same_bits(again, first, "redefine")


def one_partial_bank(n_voices, seed=11, sr=48000.0):
    """One partial per voice (cfg3's exciters): both resonator classes, DC and near-Nyquist, with and without attack
    ramp and decay, and two silent voices (no partials)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    counts = np.ones(n_voices, dtype=np.uint64)
    counts[[3, n_voices - 2]] = 0
    P = int(counts.sum())
    f = rng.uniform(15.0, 23900.0, P)
    f[:4] = [0.0, 11999.5, 12000.5, 23999.0]
    attack = rng.uniform(0.0, 400.0, P).astype(np.float32)
    attack[::5] = 0.0
    tau = rng.uniform(2000.0, 90000.0, P).astype(np.float32)
    tau[::7] = np.inf
    return dict(sample_rate=sr, voice_offsets=np.concatenate([[0], np.cumsum(counts)]).astype(np.uint64), freq_hz=f,
                amp=rng.uniform(0.1, 1.0, P).astype(np.float32), phase=rng.uniform(-3, 3, P).astype(np.float32),
                attack=attack, tau=tau)


@pytest.mark.parametrize("anchor", [0, 48, 64])
def test_one_partial_voices_kernel_vs_fp64_and_block_invariance(anchor):
    """The dedicated kernel for banks of one-partial voices (osc_one_kernel, re-anchored every 8 samples whatever
    osc_anchor says): fp64 oracle at t = 0 (attack ramps) and deep into the render, an unaligned window, and ragged block
    cuts giving the bits of one call."""
    from libfriendship_b200 import B200Renderer, KIND_OSCBANK
    nv = 70
    bank = one_partial_bank(nv)
    for idx, n in ((0, 3000), (1_000_003, 2049)):
        gpu = render_bank(B200Renderer, bank, n, idx, n_voices=nv, osc_anchor=anchor)
        ref = render_bank(OracleRenderer, bank, n, idx, n_voices=nv)
        assert np.abs(gpu.astype(np.float64) - ref).max() <= TOL * full_scale(bank), (anchor, idx)
        assert (gpu[3] == 0).all() and (gpu[nv - 2] == 0).all()
    whole = render_bank(B200Renderer, bank, 4000, 0, n_voices=nv, osc_anchor=anchor)
    r = B200Renderer(osc_anchor=anchor)
    r.define_oscbank(5, **bank)
    r.on_add_node(1, KIND_OSCBANK, 5)
    for v in range(nv):
        r.on_add_edge((1, 0, v, v))
    parts, idx = [], 0
    for n in (1, 31, 32, 33, 255, 1, 1000, 2647):
        parts.append(r.fill_buffer(nv, n, idx))
        idx += n
    assert_same_bits(np.concatenate(parts, axis=1), whole, "one-partial voices, block invariance")
