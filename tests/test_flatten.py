"""CPU (host logic, no device): routing order, buffer indexing and delay-line offsets of the device schedule are
bit-exact against an independent restatement of the flattening rules (tests/pyflatten.py), on seeded random graphs
with nested effects, shared sub-graphs, constant and signal-driven delays and extension nodes."""
import numpy as np
import pytest

from pyflatten import FULL, PyGraph, flatten, parse_dump
from randgraph import random_graph


def planner():
    from libfriendship_b200 import B200Renderer
    return B200Renderer(device=-1)


def check(rec, n_slots):
    r, g = planner(), PyGraph()
    rec.apply(r)
    rec.apply(g)
    got = parse_dump(r.dump_schedule(n_slots))
    want = flatten(g, n_slots)
    assert got["values"] == [tuple(v) for v in want["values"]]            # routing order (dependency first)
    assert got["outputs"] == want["outputs"]
    assert got["stage"] == want["stage"]
    assert got["buffer"] == want["buffer"]                                # buffer indexing
    assert [b["lookback"] for b in got["buffers"]] == [want["lookback"][v] for v in range(len(want["values"])) if want["buffer"][v] >= 0]
    assert got["n_inputs"] == want["n_inputs"]
    assert got["from_zero"] == want["from_zero"] and got["full_history"] == want["full_history"]
    return got, want


@pytest.mark.parametrize("seed", range(60))
def test_schedule_matches_python_restatement(seed):
    rec = random_graph(seed, n_inputs=3, n_nodes=8 + seed % 17, n_outputs=3, nested_levels=1 + seed % 3)
    check(rec, 3)


def test_delay_line_offsets_and_stage_cut():
    """in0 -> *0.5 -> Delay(12000) -> ... : the Delay source is a computed signal, so it is materialised with a
    lookback of exactly 12000 samples and the Delay lands in the next stage."""
    from workloads.graphs import build_cfg1_graph
    from randgraph import Recorder
    rec = Recorder()
    build_cfg1_graph(rec)
    got, want = check(rec, 2)
    # the Delay source (in0 * 0.5) is a cheap elementwise expression: it is re-evaluated at t - 12000 from the input
    # history (V_TAP + V_GATE) instead of being written to a ring and read back, so there is no buffer and one stage
    assert got["buffers"] == [] and max(got["stage"]) == 0
    taps = [v for v in got["values"] if v[0] == 10]
    gates = [v for v in got["values"] if v[0] == 11]
    assert [(v[2] << 32) | v[3] for v in taps] == [12000] and [(v[2] << 32) | v[3] for v in gates] == [12000]


def test_expensive_delay_source_is_materialised_with_exact_lookback():
    """A Delay source with more than 8 operations is materialised: ring lookback = the delay, Delay in the next stage."""
    from workloads.graphs import GraphBuilder
    from randgraph import Recorder
    from libfriendship_b200 import KIND_DELAY, KIND_MULTIPLY, KIND_SUM2
    rec = Recorder()
    g = GraphBuilder(rec)
    x = g.input(0)
    for k in range(10):
        x = g.node(KIND_SUM2 if k % 2 else KIND_MULTIPLY, x, g.input(1 + k % 3))
    d = g.node(KIND_DELAY, x, g.const(777.0))
    g.output(0, g.node(KIND_SUM2, d, x))
    got, _ = check(rec, 1)
    assert sorted(b["lookback"] for b in got["buffers"]) == [777]
    assert max(got["stage"]) == 1


def test_signal_driven_delay_needs_full_history():
    from randgraph import Recorder
    rec = Recorder()
    rec.on_add_node(1, 0)          # Delay
    rec.on_add_node(2, 3)          # Multiply (computed source)
    rec.on_add_node(3, 1)          # const
    rec.on_add_edge((0, 2, 0, 0))
    rec.on_add_edge((3, 2, 0x3f000000, 1))
    rec.on_add_edge((2, 1, 0, 0))
    rec.on_add_edge((0, 1, 1, 1))  # frames driven by input 1
    rec.on_add_edge((1, 0, 0, 0))
    got, _ = check(rec, 1)
    assert got["full_history"] and got["from_zero"]
    assert [b["lookback"] for b in got["buffers"]] == [FULL]


def test_extension_nodes_lanes_and_lookbacks():
    from randgraph import Recorder
    rec = Recorder()
    rec.define_oscbank(5, 48000.0, [0, 2, 4, 6], [1.0] * 6, [1.0] * 6, [0.0] * 6, [0.0] * 6, [0.0] * 6)
    rec.define_directform(6, [1.0] * 3, [0.0] * 3, [0.0] * 3, [0.0] * 3, [0.0] * 3)
    rec.define_fbdelay(7, [10, 200, 30], [0.5] * 3)
    rec.on_add_node(1, 32, 5)
    rec.on_add_node(2, 33, 6)
    rec.on_add_node(3, 34, 7)
    for l in range(3):
        rec.on_add_edge((1, 2, l, l))
        rec.on_add_edge((2, 3, l, l))
        rec.on_add_edge((3, 0, l, l))
    got, want = check(rec, 3)
    by_ext = {}
    for b in got["buffers"]:
        by_ext.setdefault(b["ext"], []).append(b["lookback"])
    assert by_ext[0] == [2, 2, 2]          # OscBank lanes feed the biquad: x[n-1], x[n-2] must stay addressable
    assert by_ext[1] == [2, 2, 2]          # biquad output keeps y[n-1], y[n-2]
    assert by_ext[2] == [200, 200, 200]    # feedback delay keeps its longest delay line
    assert got["from_zero"]


def test_shared_subgraph_is_evaluated_once():
    from randgraph import Recorder
    rec = Recorder()
    rec.on_add_node(1, 1)
    rec.on_add_node(2, 3)
    rec.on_add_node(3, 2)
    rec.on_add_edge((0, 2, 0, 0))
    rec.on_add_edge((1, 2, 0x40000000, 1))
    rec.on_add_edge((2, 3, 0, 0))
    rec.on_add_edge((2, 3, 0, 1))        # Sum2(m, m): m shared
    rec.on_add_edge((3, 0, 0, 0))
    rec.on_add_edge((2, 0, 0, 1))        # and also an output
    got, _ = check(rec, 2)
    assert sum(1 for v in got["values"] if v[0] == 5) == 1


def _sum_chain(r, n):
    """out0 = (...((in0 + 1) + 1) ... + 1), n Sum2 nodes in a row: the longest path has n nodes."""
    r.on_add_node(1, 1)
    prev = 0
    for i in range(n):
        h = 2 + i
        r.on_add_node(h, 2)
        r.on_add_edge((prev, h, 0, 0))
        r.on_add_edge((1, h, 0x3F800000, 1))
        prev = h
    r.on_add_edge((prev, 0, 0, 0))


def test_long_chains_flatten_on_their_own_stack_and_may_have_more_than_65535_registers():
    """A Sum2 chain of 70,000 nodes — what a bank of partials written in the reference's own primitives looks like:
    the depth-first walk needs about 600 B of stack per node (it runs on a 512 MB stack of its own), and the stage has
    more than 65,535 virtual registers (0xFFFF once doubled as 'no register')."""
    from libfriendship_b200 import RendererError
    n = 70000
    r = planner()
    _sum_chain(r, n)
    got = parse_dump(r.dump_schedule(1))
    assert sum(1 for v in got["values"] if v[0] == 4) == n                 # V_SUM2, one per node
    assert len(set(got["stage"])) == 1 and len(got["stages"]) == 1       # one fused elementwise stage
    st = got["stages"][0]
    assert st["n_regs"] <= 2                                               # a chain keeps one value live
    assert n + 2 <= len(st["instrs"]) <= n + 6                             # load, n sums, store, end markers


def test_graphs_deeper_than_the_walk_allows_are_refused_not_crashed_on():
    from libfriendship_b200 import RendererError
    from libfriendship_b200._cabi import FRB_E_UNSUPPORTED
    r = planner()
    _sum_chain(r, 500100)
    with pytest.raises(RendererError) as e:
        r.dump_schedule(1)
    assert e.value.code == FRB_E_UNSUPPORTED


def test_input_slot_that_cannot_exist_is_the_zero_signal():
    """slot 0xFFFFFFFF + 1 used to wrap the slot count to 0 in u32 and index the device table out of bounds.
    Slots at or beyond the flattener's cap (above every slot that exists; at least 65,536) are the zero signal, so the
    table covers every slot a program names."""
    from libfriendship_b200 import B200Renderer
    r = B200Renderer(device=-1)
    r.on_add_edge((0, 0, 0xFFFFFFFF, 0))
    r.on_add_edge((0, 0, 70000, 1))
    r.on_add_edge((0, 0, 65535, 2))
    d = parse_dump(r.dump_schedule(3))
    assert d["n_inputs"] == 65536
    ops = [d["values"][o][0] for o in d["outputs"]]
    assert ops == [0, 0, 2]                      # V_ZERO, V_ZERO, V_INPUT
