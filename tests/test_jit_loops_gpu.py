"""Long stage programs through the loop-emitting stage JIT (csrc/jit.cc) on the GPU: the programs that round 1 could only
interpret (NVRTC needed minutes for their straight-line form) compile in well under a second and agree bit for bit with
the interpreter, the CPU oracle and plain numpy f32 arithmetic in the same order."""
import time

import numpy as np
import pytest

from replay import assert_same_bits

pytestmark = pytest.mark.gpu


def _const_chain(r, n, c_bits=0x3F800000):
    from libfriendship_b200 import KIND_F32CONSTANT, KIND_SUM2
    r.on_add_node(1, KIND_F32CONSTANT)
    prev = 0
    for i in range(n):
        h = 2 + i
        r.on_add_node(h, KIND_SUM2)
        r.on_add_edge((prev, h, 0, 0))
        r.on_add_edge((1, h, c_bits, 1))
        prev = h
    r.on_add_edge((prev, 0, 0, 0))


@pytest.mark.parametrize("n_nodes,n_times", [(2000, 4096), (9000, 1001), (70000, 520)])
def test_long_sum2_chain_compiles_fast_and_is_bit_exact(n_nodes, n_times):
    """Sum2 = one f32 addition per node (reference.rs:228-234): numpy float32 additions in the same order round the same
    way.  The chain is one loop for the JIT whatever its length; 70,000 nodes was 'interpreted for good' in round 1."""
    from libfriendship_b200 import B200Renderer, FLAG_JIT_EAGER
    r = B200Renderer(flags=FLAG_JIT_EAGER)
    _const_chain(r, n_nodes)
    x = np.random.Generator(np.random.PCG64(n_nodes)).uniform(-1, 1, 2 * n_times).astype(np.float32)
    t0 = time.perf_counter()
    got0 = r.fill_buffer(1, n_times, 0, [x[:n_times]])
    first = time.perf_counter() - t0
    got1 = r.fill_buffer(1, n_times, n_times, [x[n_times:]])            # second block starts at an odd time for 1001
    st = r.stats()
    assert st["jit_launches"] >= 2, st
    assert first < 8.0, f"first fill_buffer took {first:.1f} s"        # flatten + NVRTC (a cache hit after the first case)
    want = x.copy()
    one = np.float32(1.0)
    for _ in range(n_nodes):
        want = want + one
    assert np.array_equal(got0[0].view(np.uint32), want[:n_times].view(np.uint32))
    assert np.array_equal(got1[0].view(np.uint32), want[n_times:].view(np.uint32))


def test_voices_of_chains_two_loop_levels_three_ways():
    """A chain of Multiply terms per voice, a gain per voice, voices summed: inner loops with per-voice trip counts inside
    an outer loop over voices.  JIT = interpreter = oracle, bit for bit, over ragged consecutive calls and a seek."""
    from libfriendship_b200 import B200Renderer, FLAG_JIT_EAGER, FLAG_NO_JIT, KIND_MULTIPLY, KIND_SUM2
    from oracle.binding import OracleRenderer
    from workloads.graphs import GraphBuilder

    def build(r):
        g = GraphBuilder(r)
        total, slot = None, 0
        for v in range(24):
            voice = None
            for p in range(12 + (v % 5)):
                term = g.node(KIND_MULTIPLY, g.input(slot % 7), g.const(1.0 / (1 + p) + 0.01 * v))
                slot += 1
                voice = term if voice is None else g.node(KIND_SUM2, voice, term)
            voice = g.node(KIND_MULTIPLY, voice, g.const(0.5 + v))
            total = voice if total is None else g.node(KIND_SUM2, total, voice)
        g.output(0, total)

    jit, itp, orc = B200Renderer(flags=FLAG_JIT_EAGER), B200Renderer(flags=FLAG_NO_JIT), OracleRenderer()
    for r in (jit, itp, orc):
        build(r)
    rng = np.random.RandomState(11)
    for idx, n in ((0, 130), (130, 257), (1000, 64)):
        rows = [rng.randn(n if k % 2 == 0 else n - 3).astype(np.float32) for k in range(7)]
        a, b, c = (r.fill_buffer(1, n, idx, rows) for r in (jit, itp, orc))
        assert_same_bits(a, c, f"jit vs oracle idx {idx}")
        assert_same_bits(b, c, f"interpreter vs oracle idx {idx}")
    assert jit.stats()["jit_launches"] >= 3 and itp.stats()["jit_launches"] == 0


@pytest.mark.parametrize("n_partials,n_times", [(256, 6000), (1024, 2048)])
def test_oscbank_is_pinned_to_reference_semantics(n_partials, n_times):
    """The headline kernel's extension node tied to the reference's own vocabulary (VERDICT r1 'weak' 1.ii): the cfg2
    bank rendered (a) by the OscBank node, (b) as a graph of the seven primitives — external inputs
    env_p(t)*sin(...) computed in fp64 -> f32, Multiply(in_p, C(amp_p)), left Sum2 chain (reference.rs:221-234) — on the
    GPU and on the pinned oracle.  (b) GPU == (b) oracle bit for bit (reference semantics: the pinned part of the oracle);
    (a) within 1e-5 of full scale of (b).  (b) is also a 3 x n_partials-instruction stage: one loop for the stage JIT."""
    from libfriendship_b200 import B200Renderer, FLAG_JIT_EAGER, KIND_OSCBANK
    from oracle.binding import OracleRenderer
    from workloads.banks import build_partial_sum_graph, full_scale, harmonic_bank, partial_signals
    bank = harmonic_bank(n_partials)
    rows = list(partial_signals(bank, n_times))
    outs = []
    for r in (B200Renderer(flags=FLAG_JIT_EAGER), B200Renderer(), OracleRenderer()):
        build_partial_sum_graph(r, bank["amp"])
        outs.append(r.fill_buffer(1, n_times, 0, rows))
    assert_same_bits(outs[0], outs[2], "reference-vocabulary bank: JIT vs oracle")
    assert_same_bits(outs[1], outs[2], "reference-vocabulary bank: default tiering vs oracle")
    r = B200Renderer()
    r.define_oscbank(7, **bank)
    r.on_add_node(2, KIND_OSCBANK, 7)
    r.on_add_edge((2, 0, 0, 0))
    a = r.fill_buffer(1, n_times, 0)
    fs = full_scale(bank)
    err = float(np.abs(a.astype(np.float64) - outs[2].astype(np.float64)).max())
    assert err <= 1e-5 * fs, f"OscBank vs reference-vocabulary graph: {err / fs:.2e} of full scale"


def test_reading_an_input_slot_that_can_never_be_fed_is_zero():
    """A toplevel edge may name any u32 input slot (RouteGraph does not check toplevel slots; RefRenderer returns 0 for a
    slot that was never fed, reference.rs:90-96).  Slot 0xFFFFFFFF used to wrap the slot count to 0 and index the
    device table out of bounds."""
    from libfriendship_b200 import B200Renderer, FLAG_JIT_EAGER, KIND_SUM2
    from oracle.binding import OracleRenderer
    outs = []
    for r in (B200Renderer(), B200Renderer(flags=FLAG_JIT_EAGER), OracleRenderer()):
        r.on_add_node(2, KIND_SUM2)
        r.on_add_edge((0, 2, 0, 0))
        r.on_add_edge((0, 2, 0xFFFFFFFF, 1))
        r.on_add_edge((2, 0, 0, 0))
        r.on_add_edge((0, 0, 4000000000, 1))
        outs.append(r.fill_buffer(2, 64, 0, [np.arange(64, dtype=np.float32)]))
    assert_same_bits(outs[0], outs[2], "interpreter")
    assert_same_bits(outs[1], outs[2], "jit")
    assert np.array_equal(outs[2][0], np.arange(64, dtype=np.float32)) and not outs[2][1].any()


def test_long_aperiodic_program_chunked_jit_three_ways():
    """1,200 nodes with no repetition to fold (random operations over four inputs and a shared side value, so that the
    loaded inputs live in registers across every chunk boundary): compiled as chunk functions, bit-exact vs the
    interpreter and the oracle over ragged consecutive calls.  (The chain only re-uses shallow values: the oracle, like
    the reference, re-evaluates a shared sub-graph once per consumer.)"""
    import random
    from libfriendship_b200 import (B200Renderer, FLAG_JIT_EAGER, FLAG_NO_JIT, KIND_DIVIDE, KIND_MINIMUM, KIND_MODULO,
                                    KIND_MULTIPLY, KIND_SUM2)
    from oracle.binding import OracleRenderer
    from workloads.graphs import GraphBuilder

    def build(r):
        rs = random.Random(2024)
        g = GraphBuilder(r)
        x = g.input(0)
        side = g.node(KIND_MULTIPLY, g.input(1), g.const(0.37))
        for k in range(1200):
            kind = rs.choice([KIND_SUM2, KIND_MULTIPLY, KIND_MINIMUM, KIND_SUM2, KIND_MODULO if k % 97 == 0 else KIND_SUM2,
                              KIND_DIVIDE if k % 89 == 0 else KIND_MULTIPLY])
            other = rs.choice([g.input(rs.randrange(1, 4)), side, g.const(rs.choice([0.5, -1.25, 2.0, 0.999]))])
            x = g.node(kind, x, other)
        g.output(0, g.node(KIND_MINIMUM, x, g.const(1e6)))
        g.output(1, g.node(KIND_SUM2, side, g.input(3)))

    jit, itp, orc = B200Renderer(flags=FLAG_JIT_EAGER), B200Renderer(flags=FLAG_NO_JIT), OracleRenderer()
    for r in (jit, itp, orc):
        build(r)
    assert "__noinline__ void frb_c0_" in jit.jit_source(2, 0)
    rng = np.random.RandomState(5)
    for idx, n in ((0, 300), (300, 1025)):
        rows = [(rng.rand(n if k != 2 else n - 5) * 0.9 + 0.05).astype(np.float32) for k in range(4)]
        a, b, c = (r.fill_buffer(2, n, idx, rows) for r in (jit, itp, orc))
        assert_same_bits(a, c, f"chunked jit vs oracle idx {idx}")
        assert_same_bits(b, c, f"interpreter vs oracle idx {idx}")
    assert jit.stats()["jit_launches"] >= 2 and itp.stats()["jit_launches"] == 0
