"""Voice sharding behind the C ABI (frb_config::n_devices, csrc/multi.cu), host logic: what may be sharded, and what each
device's schedule is.  CPU only (planning handles)."""
import numpy as np
import pytest

from pyflatten import parse_dump, V_EXT, V_TAP, V_MUL, V_SUM2, V_MIN
from workloads.banks import build_voice_mix_graph
from workloads.graphs import GraphBuilder


def _bank(nv):
    return dict(sample_rate=48000.0, voice_offsets=np.arange(nv + 1, dtype=np.uint64) * 3, freq_hz=np.full(3 * nv, 440.0),
                amp=np.ones(3 * nv, np.float32), phase=np.zeros(3 * nv, np.float32), attack=np.zeros(3 * nv, np.float32),
                tau=np.zeros(3 * nv, np.float32))


def test_mix_graph_is_lane_linear_and_every_rank_gets_the_subgraph_of_its_voices():
    from libfriendship_b200 import B200Renderer
    nv, world = 10, 4
    r = B200Renderer(device=-1)
    build_voice_mix_graph(r, _bank(nv), list(range(nv)))
    assert r.lane_use(1) == 0
    whole = parse_dump(r.dump_schedule(1))
    n_ops_whole = sum(1 for v in whole["values"] if v[0] in (V_MUL, V_SUM2))
    seen_shifts = []
    n_ops = 0
    for rank in range(world):
        d = parse_dump(r.dump_schedule_shard(1, rank, world))
        owned = list(range(rank, nv, world))
        lanes = sorted(v[3] for v in d["values"] if v[0] == V_EXT)
        assert lanes == list(range(len(owned)))                              # compact lanes 0 .. n_owned - 1
        shifts = sorted(v[3] | (v[2] << 32) for v in d["values"] if v[0] == V_TAP)
        assert shifts == [4800 + 37 * v for v in owned]                      # the owned voices' delays, by GLOBAL voice number
        seen_shifts += shifts
        n_ops += sum(1 for v in d["values"] if v[0] in (V_MUL, V_SUM2))
    assert sorted(seen_shifts) == [4800 + 37 * v for v in range(nv)]
    assert n_ops == n_ops_whole - (world - 1)        # every voice's work exactly once; world - 1 mix additions move to the exchange


def test_what_is_not_linear_in_the_lanes_is_not_shardable():
    from libfriendship_b200 import B200Renderer, KIND_MINIMUM, KIND_MULTIPLY, KIND_OSCBANK, KIND_SUM2
    def mk():
        r = B200Renderer(device=-1)
        r.define_oscbank(7, **_bank(4))
        g = GraphBuilder(r)
        r.on_add_node(100, KIND_OSCBANK, 7)
        return r, g
    r, g = mk()                                                              # no lane at all
    g.output(0, g.node(KIND_MULTIPLY, g.input(0), g.const(0.5)))
    assert r.lane_use(1) == 1
    r, g = mk()                                                              # lane through Minimum
    g.output(0, g.node(KIND_MINIMUM, (100, 0), g.const(0.5)))
    assert r.lane_use(1) == 2
    r, g = mk()                                                              # product of two lanes
    g.output(0, g.node(KIND_MULTIPLY, (100, 0), (100, 1)))
    assert r.lane_use(1) == 2
    r, g = mk()                                                              # lanes + a lane-independent signal in one slot
    g.output(0, g.node(KIND_SUM2, (100, 0), g.input(0)))
    assert r.lane_use(1) == 2
    r, g = mk()                                                              # lane x external envelope, one slot silent: fine
    g.output(0, g.node(KIND_MULTIPLY, (100, 2), g.input(0)))
    assert r.lane_use(2) == 0


def test_rank_out_of_range_is_refused():
    from libfriendship_b200 import B200Renderer, RendererError
    r = B200Renderer(device=-1)
    with pytest.raises(RendererError):
        r.dump_schedule_shard(1, 3, 2)


def test_banks_larger_than_a_launch_can_hold_are_refused_at_definition():
    """Voices / lanes ride on grid.y: more than 65,535 per bank used to fail in the middle of a render, after the input
    history had advanced.  Now the definition is refused."""
    from libfriendship_b200 import B200Renderer, RendererError
    r = B200Renderer(device=-1)
    nv = 70000
    bank = dict(sample_rate=48000.0, voice_offsets=np.arange(nv + 1, dtype=np.uint64), freq_hz=np.full(nv, 440.0),
                amp=np.ones(nv, np.float32), phase=np.zeros(nv, np.float32), attack=np.zeros(nv, np.float32), tau=np.zeros(nv, np.float32))
    with pytest.raises(RendererError) as e:
        r.define_oscbank(3, **bank)
    assert e.value.code == -7
    with pytest.raises(RendererError) as e:
        r.define_fbdelay(4, np.full(nv, 100, np.uint32), np.full(nv, 0.5, np.float32))
    assert e.value.code == -7
