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
