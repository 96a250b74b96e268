"""GPU parity proper: seeded random effect graphs (all 7 primitives, nested effects, constant and signal-driven
Delay amounts, NaN/inf/huge/negative constants, unconnected inputs, multi-slot outputs) rendered through the
C ABI on the B200 and by the CPU oracle — bit-exact, over several consecutive fill_buffer calls including
ragged inputs, seeks and graph edits between calls."""
import numpy as np
import pytest

from oracle.binding import OracleRenderer
from randgraph import random_graph, random_inputs
from replay import assert_same_bits

pytestmark = pytest.mark.gpu


def both():
    from libfriendship_b200 import B200Renderer
    return B200Renderer(), OracleRenderer()


@pytest.mark.parametrize("seed", range(40))
def test_random_graph_consecutive_calls(seed):
    gpu, orc = both()
    rec = random_graph(seed, n_inputs=2, n_nodes=10 + seed % 9, n_outputs=2, nested_levels=1 + seed % 2)
    rec.apply(gpu)
    rec.apply(orc)
    rng = np.random.RandomState(1000 + seed)
    idx = 0
    for call in range(3):
        n = int(rng.choice([1, 3, 4, 7, 64, 130, 257]))
        rows = random_inputs(rng, 2, n)
        a = gpu.fill_buffer(2, n, idx, rows)
        b = orc.fill_buffer(2, n, idx, rows)
        assert_same_bits(a, b, f"seed {seed} call {call} idx {idx} n {n}")
        idx += n


@pytest.mark.parametrize("seed", range(12))
def test_random_graph_seek_and_edit(seed):
    gpu, orc = both()
    rec = random_graph(500 + seed, n_inputs=2, n_nodes=14, n_outputs=3, nested_levels=2)
    rec.apply(gpu)
    rec.apply(orc)
    rng = np.random.RandomState(7 + seed)
    plan = [(0, 33), (33, 31), (200, 16), (216, 8), (5, 40)]       # contiguous, contiguous, seek, contiguous, seek back
    for i, (idx, n) in enumerate(plan):
        rows = random_inputs(rng, 2, n)
        if i == 2:
            # graph edit between calls: rewire output 0 to a delayed copy of input 0 (retroactive)
            for r in (gpu, orc):
                r.on_add_node(9001, 0)                              # Delay
                r.on_add_node(9002, 1)                              # F32Constant
                r.on_add_edge((0, 9001, 0, 0))
                r.on_add_edge((9002, 9001, 0x40400000, 1))          # 3.0 frames
                r.on_add_edge((9001, 0, 0, 0))
        if i == 4:
            for r in (gpu, orc):
                r.on_del_edge((9001, 0, 0, 0))
        a = gpu.fill_buffer(3, n, idx, rows)
        b = orc.fill_buffer(3, n, idx, rows)
        assert_same_bits(a, b, f"seed {seed} step {i}")


def test_cfg1_one_call_equals_blocks():
    """BASELINE configs[0]: 440 Hz sine through Multiply/Sum/Delay, 48 kHz x 1 s — one call and 94 x 512-sample
    calls are identical, and both equal the oracle bit for bit."""
    from workloads.graphs import build_cfg1_graph, cfg1_input
    n = 48000
    x = cfg1_input(n)
    gpu, orc = both()
    build_cfg1_graph(gpu)
    build_cfg1_graph(orc)
    whole = gpu.fill_buffer(2, n, 0, [x])
    ref = orc.fill_buffer(2, n, 0, [x])
    assert_same_bits(whole, ref, "cfg1 whole")
    from libfriendship_b200 import B200Renderer
    g2 = B200Renderer()
    build_cfg1_graph(g2)
    parts = []
    for s in range(0, n, 512):
        m = min(512, n - s)
        parts.append(g2.fill_buffer(2, m, s, [x[s:s + m]]))
    assert_same_bits(np.concatenate(parts, axis=1), ref, "cfg1 blocks")


def test_warmup_after_seek_constants_are_live():
    """SURVEY.md A.3 trap T1: C(0.5) -> Delay(2) rendered over 100..104 right after a seek is 0.5 x 4."""
    gpu, orc = both()
    for r in (gpu, orc):
        r.on_add_node(1, 0)
        r.on_add_node(2, 1)
        r.on_add_edge((2, 1, 0x3f000000, 0))
        r.on_add_edge((2, 1, 0x40000000, 1))
        r.on_add_edge((1, 0, 0, 0))
    a = gpu.fill_buffer(1, 4, 100)
    b = orc.fill_buffer(1, 4, 100)
    assert_same_bits(a, b)
    assert (a == 0.5).all()


def test_input_contract_errors():
    from libfriendship_b200 import RendererError
    gpu, orc = both()
    for r in (gpu, orc):
        r.on_add_edge((0, 0, 0, 0))
        r.on_add_edge((0, 0, 1, 1))
        r.fill_buffer(2, 4, 0, [[1, 2, 3, 4]])                     # slot 1 not fed: stays at length 0
        with pytest.raises(RendererError) as e:
            r.fill_buffer(2, 4, 4, [[1, 2], [5, 6]])               # slot 1 fed later: len 0 != idx 4 (reference.rs:69)
        assert e.value.code == -3
        with pytest.raises(RendererError) as e:
            r.fill_buffer(2, 2, 4, [[1, 2, 3]])                    # row longer than n_times (reference.rs:71)
        assert e.value.code == -2


@pytest.mark.parametrize("seed", range(10))
def test_sparkle_delay_flag_matches_sparkle_semantics(seed):
    """FRB_FLAG_SPARKLE_DELAY: negative / NaN delay amounts output 0.0 (reference sparkle.rs:525-542) instead of being
    clamped to delay 0 (reference.rs:205-210).  Random graphs rich in negative / NaN / signal-driven amounts."""
    from libfriendship_b200 import B200Renderer, FLAG_JIT_EAGER, FLAG_SPARKLE_DELAY
    pool = [-1.0, -0.5, float("nan"), 0.0, 1.0, 2.0, 3.5, -7.0, 1.8446744e19, 5.0]
    rec = random_graph(4000 + seed, n_inputs=2, n_nodes=14, n_outputs=2, nested_levels=1, signal_delay_prob=0.6, const_pool=pool)
    rng = np.random.RandomState(seed)
    rows = [(rng.randn(200) * 2).astype(np.float32) for _ in range(2)]
    outs = []
    for r in (B200Renderer(flags=FLAG_SPARKLE_DELAY), B200Renderer(flags=FLAG_SPARKLE_DELAY | FLAG_JIT_EAGER),
              OracleRenderer(flags=FLAG_SPARKLE_DELAY)):
        rec.apply(r)
        outs.append(r.fill_buffer(2, 200, 0, rows))
    assert_same_bits(outs[0], outs[2], f"sparkle interp seed {seed}")
    assert_same_bits(outs[1], outs[2], f"sparkle jit seed {seed}")


@pytest.mark.parametrize("seed", range(10))
def test_sparkle_min_flag_matches_sparkle_semantics(seed):
    """FRB_FLAG_SPARKLE_MIN: Minimum as select(a ULT b, a, b) — a NaN in either operand yields a (reference
    sparkle.rs:492-498) — instead of f32::min = minNum (reference.rs:242-248).  Random graphs rich in Minimum nodes,
    NaN constants and NaN-producing operations (0/0, x mod 0), interpreter and JIT against the oracle; the two
    semantics must actually differ on these graphs."""
    from libfriendship_b200 import B200Renderer, FLAG_JIT_EAGER, FLAG_SPARKLE_MIN, KIND_DIVIDE, KIND_MINIMUM, KIND_MODULO, KIND_SUM2
    from workloads.graphs import GraphBuilder
    rng = np.random.RandomState(7000 + seed)
    pool = [float("nan"), 0.0, -0.0, 1.0, -2.5, float("inf"), -float("inf"), 3.0]

    def build(r):
        g = GraphBuilder(r)
        rs = np.random.RandomState(100 + seed)
        vals = [g.input(0), g.input(1)] + [g.const(c) for c in pool]
        for k in range(24):
            kind = [KIND_MINIMUM, KIND_MINIMUM, KIND_DIVIDE, KIND_MODULO, KIND_SUM2][rs.randint(5)]
            a, b = vals[rs.randint(len(vals))], vals[rs.randint(len(vals))]
            vals.append(g.node(kind, a, b))
        mins = [v for v in vals[len(pool) + 2:]]
        g.output(0, g.node(KIND_MINIMUM, mins[-1], mins[-2]))
        g.output(1, g.node(KIND_MINIMUM, g.input(0), g.node(KIND_DIVIDE, g.input(1), g.input(1))))   # b = NaN where in1 == 0
        g.output(2, g.node(KIND_MINIMUM, g.node(KIND_DIVIDE, g.input(1), g.input(1)), g.input(0)))   # a = NaN where in1 == 0

    rows = [(rng.randn(200) * 2).astype(np.float32) for _ in range(2)]
    rows[1][::3] = 0.0
    outs = []
    for r in (B200Renderer(flags=FLAG_SPARKLE_MIN), B200Renderer(flags=FLAG_SPARKLE_MIN | FLAG_JIT_EAGER),
              OracleRenderer(flags=FLAG_SPARKLE_MIN), OracleRenderer()):
        build(r)
        outs.append(r.fill_buffer(3, 200, 0, rows))
    assert_same_bits(outs[0], outs[2], f"sparkle-min interp seed {seed}")
    assert_same_bits(outs[1], outs[2], f"sparkle-min jit seed {seed}")
    # slot 1: Sparkle keeps a (a number) where b is NaN, like minNum; slot 2: Sparkle yields NaN where a is NaN, minNum the number
    assert np.isnan(outs[2][2][::3]).all() and not np.isnan(outs[3][2][::3]).any()
