"""GPU: N4 — frb_render_stream / Dispatch.render_stream (pipelined blocks, pinned double-buffered staging both ways)
give the same bits as the plain fill_buffer calls they stand for, which in turn match the oracle."""
import numpy as np
import pytest

from workloads.banks import build_voice_mix_graph, detuned_bank
from workloads.graphs import build_cfg1_graph, cfg1_input
from oracle.binding import OracleRenderer
from replay import assert_same_bits

pytestmark = pytest.mark.gpu


def collect(n_slots):
    got = []

    def sink(block, idx):
        assert block.shape[0] == n_slots
        got.append((idx, block.copy()))     # the view is only valid during the call
    return got, sink


@pytest.mark.parametrize("block", [512, 1000, 4096, 48000, 100000])
def test_stream_with_inputs_bit_exact_vs_oracle(block):
    """cfg1 graph (external input through Multiply/Delay/Sum + side chain), inputs fetched block by block."""
    from libfriendship_b200 import B200Renderer
    n = 48000
    x = cfg1_input(n)
    o = OracleRenderer()
    build_cfg1_graph(o)
    ref = o.fill_buffer(2, n, 0, [x])
    r = B200Renderer()
    build_cfg1_graph(r)
    got, sink = collect(2)
    r.render_stream(2, 0, n, block, sink, n_in_rows=1, source=lambda t, m: x[t:t + m][None, :])
    assert [g[0] for g in got] == list(range(0, n, min(block, n)))
    assert_same_bits(np.concatenate([g[1] for g in got], axis=1), ref, f"stream block={block}")


def test_stream_equals_fill_buffer_calls_and_continues():
    """Oscillator-bank graph: streamed blocks == one fill_buffer call; a stream that starts at the previous head
    continues (no seek), one that starts elsewhere is a seek — exactly like the calls it stands for."""
    from libfriendship_b200 import B200Renderer
    bank, ids = detuned_bank(3, 200)
    a, b = B200Renderer(), B200Renderer()
    build_voice_mix_graph(a, bank, ids, delay0=300.0)
    build_voice_mix_graph(b, bank, ids, delay0=300.0)
    whole = a.fill_buffer(1, 20000, 0)
    got, sink = collect(1)
    b.render_stream(1, 0, 12345, 3000, sink)
    b.render_stream(1, 12345, 20000 - 12345, 2048, sink)
    assert_same_bits(np.concatenate([g[1] for g in got], axis=1), whole, "stream vs fill")
    got2, sink2 = collect(1)
    b.render_stream(1, 5000, 1000, 256, sink2)      # seek back
    assert_same_bits(np.concatenate([g[1] for g in got2], axis=1), whole[:, 5000:6000], "stream seek")


def test_stream_zero_length_and_errors():
    from libfriendship_b200 import B200Renderer, RendererError
    r = B200Renderer()
    build_cfg1_graph(r)
    got, sink = collect(2)
    r.render_stream(2, 0, 0, 512, sink)
    assert len(got) == 1 and got[0][1].shape == (2, 0)
    with pytest.raises(RendererError):
        r.render_stream(2, 0, 100, 0, sink)

    def bad_sink(block, idx):
        raise KeyError("client failed")
    with pytest.raises(KeyError):
        r.render_stream(2, 0, 4096, 512, bad_sink)
    # the renderer is still usable after an aborted stream
    x = cfg1_input(1024)
    o = OracleRenderer()
    build_cfg1_graph(o)
    assert_same_bits(r.fill_buffer(2, 1024, 0, [x]), o.fill_buffer(2, 1024, 0, [x]), "after abort")


def test_dispatch_render_stream_to_wav(tmp_path):
    """Dispatch -> RouteGraph -> renderer -> WavClient: the file holds the bits the oracle renders."""
    import struct
    from libfriendship_b200.dispatch import Dispatch, EffectId, WavClient
    from workloads.graphs import f32_bits
    path = tmp_path / "const.wav"
    client = WavClient(path, 1, 48000)
    d = Dispatch(client)
    # Delay(C(0.5), C(7)) -> out0: zeros for 7 samples, then 0.5 (tests/render_prim.rs:101-129 at another delay)
    d.add_node(1, EffectId.primitive("F32Constant"))
    d.add_node(2, EffectId.primitive("Delay"))
    d.add_edge((1, 2, f32_bits(0.5), 0))
    d.add_edge((1, 2, f32_bits(7.0), 1))
    d.add_edge((2, 0, 0, 0))
    d.render_stream(0, 10000, 1, 1024)
    client.close()
    raw = open(path, "rb").read()
    data = np.frombuffer(raw[58:], dtype="<f4")
    assert struct.unpack("<I", raw[54:58])[0] == 40000 and len(data) == 10000
    assert (data[:7] == 0).all() and (data[7:] == 0.5).all()
