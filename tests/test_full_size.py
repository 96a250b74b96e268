"""GPU: BASELINE.json's configurations at FULL size, checked through size-independent properties and oracle spot
checks (the CPU oracle cannot render 10 s of 4 million partials, but it can render any window of it exactly,
because the render is a pure function of absolute time)."""
import numpy as np
import pytest

from workloads.banks import build_voice_mix_graph, detuned_bank, full_scale, harmonic_bank
from oracle.binding import OracleRenderer
from replay import assert_same_bits

pytestmark = pytest.mark.gpu
TOL = 1e-5


def gpu():
    from libfriendship_b200 import B200Renderer
    return B200Renderer()


def test_cfg2_full_ten_seconds_spot_checked_against_fp64():
    """cfg2: 1,024 harmonic partials x 1 voice, 48 kHz x 10 s.  The whole render runs on the GPU; the oracle
    re-renders windows of it (start, attack region, middle, end)."""
    from libfriendship_b200 import KIND_OSCBANK
    bank = harmonic_bank(1024)
    n = 480000
    r = gpu()
    r.define_oscbank(5, **bank)
    r.on_add_node(1, KIND_OSCBANK, 5)
    r.on_add_edge((1, 0, 0, 0))
    out = r.fill_buffer(1, n, 0)
    assert np.isfinite(out).all()
    fs = full_scale(bank)
    o = OracleRenderer()
    o.define_oscbank(5, **bank)
    o.on_add_node(1, KIND_OSCBANK, 5)
    o.on_add_edge((1, 0, 0, 0))
    for start in (0, 250, 65536 - 64, 240000, n - 128):
        w = o.fill_buffer(1, 128, start)
        err = np.abs(out[:, start:start + 128].astype(np.float64) - w).max()
        assert err <= TOL * fs, (start, err)
    # envelope sanity at full length: the tail has decayed (tau <= 2.2 s) and the attack starts from silence
    assert abs(out[0, 0]) <= 1e-6 and np.abs(out[0, -4800:]).max() < 0.2 * np.abs(out[0, :48000]).max()


def test_cfg4_full_partial_count_windows_and_linearity():
    """cfg4 shape at full partial count (65,536 per voice) on a shortened render: two voices through the per-voice
    Delay/mix graph vs the oracle on windows around the delay taps, and linearity of the voice mix: rendering
    voices {0,1} equals the sum of rendering {0} and {1} (what the multi-GPU shard + reduce relies on)."""
    n_partials, n = 65536, 6000
    bank, ids = detuned_bank(2, n_partials)
    r = gpu()
    build_voice_mix_graph(r, bank, ids)
    both = r.fill_buffer(1, n, 0)
    fs = full_scale(bank) * 2 * 1.3
    o = OracleRenderer()
    build_voice_mix_graph(o, bank, ids)
    for start in (0, 4800 - 8, 4837 + 100):
        w = o.fill_buffer(1, 24, start)
        err = np.abs(both[:, start:start + 24].astype(np.float64) - w).max()
        assert err <= TOL * fs, (start, err)
    parts = []
    for v in (0, 1):
        bank_v, ids_v = detuned_bank(2, n_partials, voices=[v])
        rv = gpu()
        build_voice_mix_graph(rv, bank_v, ids_v)
        parts.append(rv.fill_buffer(1, n, 0))
    assert np.abs((parts[0].astype(np.float64) + parts[1]) - both).max() <= 4e-7 * fs   # f32 order of one add


def test_cfg4_one_call_equals_blocks_at_full_partial_count():
    n_partials = 65536
    bank, ids = detuned_bank(1, n_partials, seed=2)
    a = gpu()
    build_voice_mix_graph(a, bank, ids)
    whole = a.fill_buffer(1, 9000, 0)
    b = gpu()
    build_voice_mix_graph(b, bank, ids)
    parts, idx = [], 0
    for m in (512, 4096, 100, 4292):
        parts.append(b.fill_buffer(1, m, idx))
        idx += m
    assert_same_bits(np.concatenate(parts, axis=1), whole, "cfg4 blocks")


def test_cfg3_full_voice_count_selected_voices_vs_fp64():
    """cfg3: 4,096 voices, biquad + feedback delay each, excited by a 1-partial oscillator per voice, 1 s of audio
    at the full voice count; eight voices spread over the range are also routed to their own output slots and
    compared with the fp64 oracle, which renders just those voices."""
    from workloads.filters import build_cfg3_graph, cfg3_filters
    from libfriendship_b200 import KIND_DIRECTFORM, KIND_FBDELAY, KIND_OSCBANK
    n_voices, n = 4096, 48000
    bank, _ = detuned_bank(n_voices, 1, seed=5)
    r = gpu()
    build_cfg3_graph(r, n_voices, excitation="osc", bank=bank, mix_to_one=True)
    picks = [0, 1, 511, 900, 2047, 3000, 4000, 4095]
    for k, v in enumerate(picks):
        r.on_add_edge((11, 0, v, 1 + k))                       # FbDelay lane v -> output slot 1+k
    out = r.fill_buffer(1 + len(picks), n, 0)
    assert np.isfinite(out).all()
    # oracle: only the picked voices (each voice is independent)
    (b0, b1, b2, a1, a2), delay, gain = cfg3_filters(n_voices)
    sel = np.array(picks)
    o = OracleRenderer()
    vo = bank["voice_offsets"]
    sub = dict(sample_rate=bank["sample_rate"], voice_offsets=np.arange(len(picks) + 1, dtype=np.uint64),
               freq_hz=bank["freq_hz"][sel], amp=bank["amp"][sel], phase=bank["phase"][sel],
               attack=bank["attack"][sel], tau=bank["tau"][sel])
    o.define_oscbank(13, **sub)
    o.define_directform(11, b0[sel], b1[sel], b2[sel], a1[sel], a2[sel])
    o.define_fbdelay(12, delay[sel], gain[sel])
    o.on_add_node(12, KIND_OSCBANK, 13)
    o.on_add_node(10, KIND_DIRECTFORM, 11)
    o.on_add_node(11, KIND_FBDELAY, 12)
    for k in range(len(picks)):
        o.on_add_edge((12, 10, k, k))
        o.on_add_edge((10, 11, k, k))
        o.on_add_edge((11, 0, k, k))
    want = o.fill_buffer(len(picks), n, 0)
    scale = np.abs(want).max()
    assert np.abs(out[1:].astype(np.float64) - want).max() <= 1e-4 * scale
    # the mix slot is the left fold of all 4,096 voices: at least bounded and consistent with the picked voices' scale
    assert np.abs(out[0]).max() <= 4096 * scale


def test_cfg5_shape_one_voice_of_a_million_partials_at_the_end_of_a_minute():
    """cfg5: 2^20 partials per voice, 192 kHz, sample index ~11.5 M (60 s).  One voice, a 48-sample window at the
    very end of the render: exercises the exact fixed-point phase at large t and a 48 MB coefficient stream."""
    from libfriendship_b200 import KIND_OSCBANK
    sr, n_partials = 192000.0, 1 << 20
    rng = np.random.Generator(np.random.PCG64(2))
    k = np.arange(1, n_partials + 1, dtype=np.float64)
    f = 55.0 * k * (1.0 + rng.uniform(-0.002, 0.002, n_partials))
    f = np.where(f >= sr / 2, np.mod(f, sr / 2 * 0.98) + 20.0, f)
    bank = dict(sample_rate=sr, voice_offsets=np.array([0, n_partials], dtype=np.uint64), freq_hz=f,
                amp=(1.0 / k).astype(np.float32), phase=np.zeros(n_partials, dtype=np.float32),
                attack=(192.0 * (1 + (k.astype(np.int64) % 7))).astype(np.float32),
                tau=(sr * (20.0 + 200.0 / k)).astype(np.float32))       # slow decay: still audible after 60 s
    idx, n = 11_520_000 - 48, 48
    outs = []
    for cls in (gpu().__class__, OracleRenderer):
        r = cls()
        r.define_oscbank(5, **bank)
        r.on_add_node(1, KIND_OSCBANK, 5)
        r.on_add_edge((1, 0, 0, 0))
        outs.append(r.fill_buffer(1, n, idx))
    fs = full_scale(bank)
    assert np.abs(outs[1]).max() > 1e-3 * fs          # the signal is alive there
    assert np.abs(outs[0].astype(np.float64) - outs[1]).max() <= TOL * fs


def test_device_resident_inputs_and_outputs_match_the_host_path():
    """frb_fill_buffer_device: inputs and outputs stay in HBM (what bench.py's `value` times)."""
    import torch
    from workloads.graphs import build_cfg1_graph, cfg1_input
    n = 20000
    x = cfg1_input(n)
    host = gpu()
    build_cfg1_graph(host)
    want = host.fill_buffer(2, n, 0, [x])
    dev = gpu()
    build_cfg1_graph(dev)
    d_in = torch.from_numpy(x).cuda()
    d_out = torch.empty((2, n), dtype=torch.float32, device="cuda")
    dev.fill_buffer_device(d_out.data_ptr(), 2, n, 0, d_in.data_ptr(), [0, n])
    dev.sync()
    assert_same_bits(d_out.cpu().numpy(), want, "device path")
