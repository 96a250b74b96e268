"""GPU: corners of the renderer contract that the reference's own tests do not pin (parity by source reading of
src/render/reference.rs), oracle vs CUDA, bit-exact."""
import numpy as np
import pytest

from workloads.graphs import build_cfg1_graph, cfg1_input, f32_bits
from oracle.binding import OracleRenderer
from replay import assert_same_bits

pytestmark = pytest.mark.gpu


def pair():
    from libfriendship_b200 import B200Renderer
    return B200Renderer(), OracleRenderer()


def test_seek_far_beyond_2_pow_32():
    """Absolute sample indices are u64 (renderer.rs:16): a seek to 2^33 + 12345, two contiguous blocks."""
    g, o = pair()
    build_cfg1_graph(g, delay=100.0)
    build_cfg1_graph(o, delay=100.0)
    idx = (1 << 33) + 12345
    x = cfg1_input(600)
    for k in range(2):
        blk = [x[300 * k:300 * (k + 1)]]
        assert_same_bits(g.fill_buffer(2, 300, idx + 300 * k, blk), o.fill_buffer(2, 300, idx + 300 * k, blk), f"block {k}")


def test_oscillator_phase_is_exact_beyond_2_pow_32():
    from workloads.banks import full_scale, harmonic_bank
    from libfriendship_b200 import KIND_OSCBANK
    bank = harmonic_bank(64)
    bank["tau"] = np.full(64, np.inf, dtype=np.float32)          # keep it audible at t = 2^32
    idx = (1 << 32) + 777
    outs = []
    for r in pair():
        r.define_oscbank(5, **bank)
        r.on_add_node(1, KIND_OSCBANK, 5)
        r.on_add_edge((1, 0, 0, 0))
        outs.append(r.fill_buffer(1, 400, idx))
    assert np.abs(outs[0].astype(np.float64) - outs[1]).max() <= 1e-5 * full_scale(bank)


def test_nested_effect_is_a_private_copy_and_nodes_can_be_deleted():
    """reference.rs:98-113: make_node deep-copies the nested graph when the node is added — redefining the effect
    afterwards must not change existing nodes; reference.rs:121-123,127-136: del_node / del_edge."""
    def body(factor):
        return [(1, 3, 0), (2, 1, 0)], [(0, 1, 0, 0), (1, 0, 0, 0), (2, 1, f32_bits(factor), 1)]
    g, o = pair()
    x = np.arange(1, 9, dtype=np.float32)
    for r in (g, o):
        r.define_effect(7, *body(5.0))
        r.on_add_node(1, 16, 7)                       # instance of "x5"
        r.on_add_edge((0, 1, 0, 0))
        r.on_add_edge((1, 0, 0, 0))
        r.define_effect(7, *body(7.0))                # same key, new definition
        r.on_add_node(2, 16, 7)                       # instance of "x7"
        r.on_add_edge((0, 2, 0, 0))
        r.on_add_edge((2, 0, 0, 1))
    a, b = g.fill_buffer(2, 8, 0, [x]), o.fill_buffer(2, 8, 0, [x])
    assert_same_bits(a, b)
    assert (a[0] == 5 * x).all() and (a[1] == 7 * x).all()
    for r in (g, o):
        r.on_del_edge((2, 0, 0, 1))                   # slot 1 becomes None -> renders 0
        r.on_del_edge((0, 2, 0, 0))
        r.on_del_node(2)
    a, b = g.fill_buffer(2, 8, 8, [x]), o.fill_buffer(2, 8, 8, [x])
    assert_same_bits(a, b)
    assert (a[1] == 0).all()


def test_more_outputs_requested_than_connected_and_zero_length_calls():
    g, o = pair()
    for r in (g, o):
        r.on_add_edge((0, 0, 0, 3))
    x = np.arange(5, dtype=np.float32)
    assert_same_bits(g.fill_buffer(6, 5, 0, [x]), o.fill_buffer(6, 5, 0, [x]))
    from libfriendship_b200 import RendererError
    for r in (g, o):
        assert r.fill_buffer(6, 0, 5).shape == (6, 0)          # zero samples: head stays 5
        assert r.fill_buffer(0, 4, 5).shape == (0, 4)          # zero slots: no slot is fed, head moves to 9 (reference.rs:84)
    for r in (g, o):
        with pytest.raises(RendererError) as e:                # slot 0 still has length 5 != 9 (reference.rs:69)
            r.fill_buffer(6, 3, 9, [x[:3]])
        assert e.value.code == -3
    assert_same_bits(g.fill_buffer(6, 3, 20, [x[:3]]), o.fill_buffer(6, 3, 20, [x[:3]]))   # a seek resets the slots


def test_edge_from_a_missing_node_is_an_error_not_a_crash():
    from libfriendship_b200 import RendererError
    g, o = pair()
    for r in (g, o):
        r.on_add_edge((5, 0, 0, 0))                   # node 5 was never added (reference.rs:186 would panic)
        with pytest.raises(RendererError) as e:
            r.fill_buffer(1, 4, 0)
        assert e.value.code == -1


def test_internal_block_length_does_not_change_the_result(monkeypatch):
    """One fill is cut into internal blocks (FRB_BLOCK_SAMPLES overrides their length): the primitive path and the
    oscillator bank are pure functions of absolute time, so 1,024-sample blocks give the bits of one block; the
    recurrences restart their scan tiles at every block start, so they agree within their tolerance."""
    from workloads.banks import build_voice_mix_graph, detuned_bank
    from workloads.filters import build_cfg3_graph
    from libfriendship_b200 import B200Renderer
    n = 20000
    x = cfg1_input(n)
    bank, ids = detuned_bank(3, 40)
    bank1, _ = detuned_bank(5, 1, seed=5)
    outs = {}
    for blk in (None, "1024"):
        if blk:
            monkeypatch.setenv("FRB_BLOCK_SAMPLES", blk)
        r = B200Renderer()
        build_cfg1_graph(r)
        a = r.fill_buffer(2, n, 0, [x])
        r = B200Renderer()
        build_voice_mix_graph(r, bank, ids, delay0=480.0)
        b = r.fill_buffer(1, n, 0)
        r = B200Renderer()
        build_cfg3_graph(r, 5, excitation="osc", bank=bank1, mix_to_one=True)
        c = r.fill_buffer(1, n, 0)
        outs[blk] = (a, b, c)
    assert_same_bits(outs[None][0], outs["1024"][0], "primitive graph, internal blocks")
    assert_same_bits(outs[None][1], outs["1024"][1], "oscillator bank + delay + mix, internal blocks")
    scale = np.abs(outs[None][2]).max()
    assert np.abs(outs[None][2].astype(np.float64) - outs["1024"][2]).max() <= 1e-5 * scale


def test_many_long_input_rows_ragged_and_unaligned():
    """Five input rows of > 200,000 samples each take the batched ingest (one kernel for all rows): rows of different
    lengths (the short ones are padded with their last value, reference.rs:72-73), a row of odd length (so the rows after
    it start at unaligned offsets), a Delay reaching back into the previous call's rows; bit-exact against the oracle."""
    from workloads.graphs import GraphBuilder
    from libfriendship_b200 import B200Renderer, KIND_DELAY, KIND_SUM2
    rng = np.random.Generator(np.random.PCG64(8))
    n = 262144
    lens = [n, n - 1, n - 4097, 7, n]
    outs = []
    for cls in (B200Renderer, OracleRenderer):
        r = cls()
        g = GraphBuilder(r)
        for s in range(5):
            d = g.node(KIND_DELAY, g.input(s), g.const(float(1000 + 37 * s)))
            g.output(s, g.node(KIND_SUM2, g.input(s), d))
        got = []
        for call in range(2):
            rows = [rng.uniform(-1, 1, m).astype(np.float32) for m in lens] if cls is B200Renderer else saved[call]
            got.append((rows, r.fill_buffer(5, n, call * n, rows)))
        if cls is B200Renderer:
            saved = [rows for rows, _ in got]
        outs.append(np.concatenate([o for _, o in got], axis=1))
    assert_same_bits(outs[0], outs[1], "batched ingest of ragged, unaligned rows")


def test_a_refused_call_changes_nothing():
    """The reference asserts on a row longer than n_times (reference.rs:71) or a slot fed after it was skipped (:69); the
    C ABI refuses such a call BEFORE any state changes: the next call is still contiguous and still reads the history
    the earlier calls supplied (here through a Delay of 2)."""
    from libfriendship_b200 import B200Renderer, KIND_DELAY, RendererError
    from workloads.graphs import GraphBuilder
    r = B200Renderer()
    g = GraphBuilder(r)
    g.output(0, g.node(KIND_DELAY, g.input(0), g.const(2.0)))
    a = r.fill_buffer(1, 4, 0, [[1, 2, 3, 4]])
    assert a[0].tolist() == [0, 0, 1, 2]
    for bad in ([[9, 9, 9, 9, 9]],):                               # longer than n_times
        with pytest.raises(RendererError) as e:
            r.fill_buffer(1, 4, 4, bad)
        assert e.value.code == -2
    b = r.fill_buffer(1, 4, 4, [[5, 6, 7, 8]])                     # contiguous: NOT a seek, history intact
    assert b[0].tolist() == [3, 4, 5, 6]
    c = r.fill_buffer(1, 4, 100, [[1, 1, 1, 1]])                   # a seek (idx != 8) is legal and forgets the history (renderer.rs:12-15)
    assert c[0].tolist() == [0, 0, 1, 1]
