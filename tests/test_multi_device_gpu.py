"""N B200s behind ONE renderer handle (frb_config::n_devices, csrc/multi.cu): the reference's caller owns one renderer
inside Dispatch (reference src/dispatch.rs:99-106, :147-153).  Needs at least two visible CUDA devices
(`gpurun --gpus 2`); skipped otherwise."""
import numpy as np
import pytest

from replay import assert_same_bits

pytestmark = pytest.mark.gpu


def _n_devices():
    import torch
    return torch.cuda.device_count()


needs2 = pytest.mark.skipif("_n_devices() < 2", reason="needs two CUDA devices")


@needs2
def test_sharded_render_equals_one_device_render():
    """cfg4-shaped graph (bank -> per-voice Delay/mix -> Sum2 chain): N devices vs one, the whole render, host path and
    device path, one call and consecutive calls, and after re-defining the bank.  Voices are the same bits on any device;
    only the order of the mix additions differs: <= 4e-7 of full scale."""
    from libfriendship_b200 import B200Renderer
    from workloads.banks import build_voice_mix_graph, detuned_bank, full_scale
    n = min(_n_devices(), 4)
    bank, ids = detuned_bank(10, 700)
    one, many = B200Renderer(), B200Renderer(n_devices=n)
    for r in (one, many):
        build_voice_mix_graph(r, bank, ids, delay0=480.0)
    nt = 20000
    a, b = one.fill_buffer(1, nt, 0), many.fill_buffer(1, nt, 0)
    fs = full_scale(bank) * len(ids) * 1.3
    assert float(np.abs(a.astype(np.float64) - b).max()) <= 4e-7 * fs
    assert float(np.abs(a).max()) > 0.01 * fs                       # not comparing silence with silence
    # consecutive calls (odd cut) continue the same render
    c = np.concatenate([many.fill_buffer(1, 9001, 0), many.fill_buffer(1, nt - 9001, 9001)], axis=1)
    assert float(np.abs(c.astype(np.float64) - b).max()) <= 4e-7 * fs
    # new parameters for the same node: every device takes over its share
    bank2, _ = detuned_bank(10, 700, seed=9)
    for r in (one, many):
        r.define_oscbank(7, **bank2)
    a2, b2 = one.fill_buffer(1, nt, 0), many.fill_buffer(1, nt, 0)
    assert float(np.abs(a2.astype(np.float64) - b2).max()) <= 4e-7 * fs
    assert float(np.abs(a2 - a).max()) > 1e-3 * fs
    st = many.stats()
    assert st["osc_launches"] >= 2 * n                              # every device ran its own bank kernels


@needs2
def test_graph_without_bank_lanes_is_bit_exact_on_several_devices():
    """Nothing to shard: the result is device 0's render of the whole graph, bit for bit the oracle's — including a Delay
    attached between calls that reads input history of the earlier call (reference tests/ext_input.rs:84-122)."""
    from libfriendship_b200 import B200Renderer, KIND_DELAY
    from oracle.binding import OracleRenderer
    from workloads.graphs import build_cfg1_graph, cfg1_input
    x = cfg1_input(9000)
    outs = []
    for r in (B200Renderer(n_devices=2), OracleRenderer()):
        g = build_cfg1_graph(r, delay=100.0)
        o1 = r.fill_buffer(2, 4000, 0, [x[:4000]])
        d = g.node(KIND_DELAY, g.input(0), g.const(777.0))
        g.output(1, d)
        o2 = r.fill_buffer(2, 5000, 4000, [x[4000:]])
        outs.append((o1, o2))
    assert_same_bits(outs[0][0], outs[1][0], "first call")
    assert_same_bits(outs[0][1], outs[1][1], "second call, after the edit")


@needs2
def test_graph_that_cannot_be_sharded_is_refused():
    from libfriendship_b200 import B200Renderer, KIND_MINIMUM, KIND_OSCBANK, RendererError
    from workloads.banks import detuned_bank
    from workloads.graphs import GraphBuilder
    bank, _ = detuned_bank(4, 64)
    r = B200Renderer(n_devices=2)
    r.define_oscbank(7, **bank)
    g = GraphBuilder(r)
    r.on_add_node(100, KIND_OSCBANK, 7)
    g.output(0, g.node(KIND_MINIMUM, (100, 0), (100, 1)))
    with pytest.raises(RendererError) as e:
        r.fill_buffer(1, 256, 0)
    assert e.value.code == -7


def test_more_devices_than_the_box_has_is_refused():
    from libfriendship_b200 import B200Renderer, RendererError
    with pytest.raises(RendererError):
        B200Renderer(n_devices=64)


@needs2
def test_external_inputs_reach_every_device_and_streaming_matches():
    """Lanes multiplied by an external envelope (linear in the lanes: shardable), the envelope also read through a Delay:
    the caller's input rows must reach every device's history, over ragged consecutive calls, and frb_render_stream on
    several devices must deliver the same blocks as the fill_buffer calls it stands for."""
    from libfriendship_b200 import B200Renderer, KIND_DELAY, KIND_MULTIPLY, KIND_OSCBANK, KIND_SUM2
    from workloads.banks import detuned_bank, full_scale
    from workloads.graphs import GraphBuilder
    bank, ids = detuned_bank(6, 300)

    def build(r):
        r.define_oscbank(7, **bank)
        g = GraphBuilder(r)
        r.on_add_node(100, KIND_OSCBANK, 7)
        env_late = g.node(KIND_DELAY, g.input(0), g.const(37.0))
        total = None
        for v in range(6):
            voice = g.node(KIND_MULTIPLY, (100, v), g.input(0) if v % 2 else env_late)
            total = voice if total is None else g.node(KIND_SUM2, total, voice)
        g.output(0, total)
        g.output(1, g.node(KIND_MULTIPLY, (100, 3), g.const(0.25)))

    one, many, strm = B200Renderer(), B200Renderer(n_devices=2), B200Renderer(n_devices=2)
    for r in (one, many, strm):
        build(r)
    assert many.lane_use(2) == 0
    rng = np.random.RandomState(3)
    env = np.abs(rng.randn(6000)).astype(np.float32)
    fs = full_scale(bank) * 6 * float(env.max())
    outs = {id(one): [], id(many): []}
    for idx, n in ((0, 2500), (2500, 1501), (4001, 1999)):
        row = env[idx:idx + n - (7 if idx else 0)]              # ragged: padded with its last value (reference.rs:72-73)
        for r in (one, many):
            outs[id(r)].append(r.fill_buffer(2, n, idx, [row]))
    a, b = np.concatenate(outs[id(one)], axis=1), np.concatenate(outs[id(many)], axis=1)
    assert float(np.abs(a.astype(np.float64) - b).max()) <= 4e-7 * fs
    assert float(np.abs(a).max()) > 1e-3 * fs
    # streaming: three blocks of 2,000 with the same (unragged) envelope
    want = B200Renderer(n_devices=2)
    build(want)
    w = np.concatenate([want.fill_buffer(2, 2000, k * 2000, [env[k * 2000:(k + 1) * 2000]]) for k in range(3)], axis=1)
    got = np.zeros_like(w)

    def sink(blk, t):
        got[:, t:t + blk.shape[1]] = blk

    strm.render_stream(2, 0, 6000, 2000, sink, n_in_rows=1, source=lambda t, n: env[None, t:t + n])
    assert np.array_equal(got.view(np.uint32), w.view(np.uint32))
