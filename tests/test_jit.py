"""The stage JIT (csrc/jit.cc): the fused elementwise + Delay path compiled by NVRTC into one straight-line sm_100a
kernel per stage — the B200 counterpart of the reference's LLVM JIT renderer (src/render/sparkle.rs)."""
import numpy as np
import pytest

from workloads.graphs import build_cfg1_graph, cfg1_input
from randgraph import random_graph, random_inputs
from replay import assert_same_bits, load_golden, replay


def test_generated_source_compiles_for_sm_100a_without_a_gpu():
    """CPU: codegen + NVRTC (--gpu-architecture=sm_100a --fmad=false): the cfg1 graph (single stage: the Delay
    source is re-evaluated at t - 12000) and a graph whose Delay source is materialised (two stages)."""
    from workloads.graphs import GraphBuilder
    from libfriendship_b200 import B200Renderer, KIND_DELAY, KIND_MULTIPLY, KIND_SUM2
    r = B200Renderer(device=-1)
    build_cfg1_graph(r)
    src0 = r.jit_source(2, 0)
    assert "frb_stage" in src0 and "f4tap_in" in src0 and "f4gate" in src0 and "f4st_out" in src0
    assert r.jit_cubin_size(2, 0) > 1000
    r2 = B200Renderer(device=-1)
    g = GraphBuilder(r2)
    x = g.input(0)
    for k in range(10):
        x = g.node(KIND_SUM2 if k % 2 else KIND_MULTIPLY, x, g.input(1 + k % 3))
    g.output(0, g.node(KIND_SUM2, g.node(KIND_DELAY, x, g.const(777.0)), x))
    s0, s1 = r2.jit_source(1, 0), r2.jit_source(1, 1)
    assert "f4st_buf" in s0                                        # stage 0 materialises the Delay source
    assert "f4delay<1>" in s1 and "f4st_out" in s1                 # stage 1 reads it back at t - 777
    assert r2.jit_cubin_size(1, 0) > 1000 and r2.jit_cubin_size(1, 1) > 1000


@pytest.mark.parametrize("seed", range(8))
def test_random_programs_compile(seed):
    from libfriendship_b200 import B200Renderer
    r = B200Renderer(device=-1)
    random_graph(300 + seed, n_inputs=2, n_nodes=16, n_outputs=3, nested_levels=2).apply(r)
    w = r.dump_schedule(3)
    n_stages = int(w[4])
    for s in range(n_stages):
        assert r.jit_cubin_size(3, s) > 0


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(24))
def test_jit_equals_interpreter_equals_oracle(seed):
    """Bit-exact three ways on seeded random graphs over several calls (ragged inputs, seek)."""
    from libfriendship_b200 import B200Renderer, FLAG_JIT_EAGER, FLAG_NO_JIT
    from oracle.binding import OracleRenderer
    rec = random_graph(900 + seed, n_inputs=2, n_nodes=10 + seed % 11, n_outputs=2, nested_levels=1 + seed % 2)
    jit, itp, orc = B200Renderer(flags=FLAG_JIT_EAGER), B200Renderer(flags=FLAG_NO_JIT), OracleRenderer()
    for r in (jit, itp, orc):
        rec.apply(r)
    rng = np.random.RandomState(50 + seed)
    for idx, n in ((0, 70), (70, 129), (500, 33)):
        rows = random_inputs(rng, 2, n)
        a, b, c = (r.fill_buffer(2, n, idx, rows) for r in (jit, itp, orc))
        assert_same_bits(a, c, f"jit vs oracle seed {seed} idx {idx}")
        assert_same_bits(b, c, f"interpreter vs oracle seed {seed} idx {idx}")
    assert jit.stats()["jit_launches"] > 0
    assert itp.stats()["jit_launches"] == 0


@pytest.mark.gpu
@pytest.mark.parametrize("test", load_golden(), ids=lambda t: t["name"])
def test_reference_golden_vectors_through_the_jit(test):
    from libfriendship_b200 import B200Renderer, FLAG_JIT_EAGER
    replay(B200Renderer(flags=FLAG_JIT_EAGER), test)


@pytest.mark.gpu
def test_hot_stage_gets_compiled_in_the_background():
    """Default policy when streaming short blocks: interpreted until hot (4th launch), then NVRTC compiles beside the
    render loop and the compiled kernel takes over without ever stalling a block — same bits before and after."""
    import time
    from libfriendship_b200 import B200Renderer
    from oracle.binding import OracleRenderer
    from libfriendship_b200 import KIND_DIVIDE, KIND_MINIMUM, KIND_MODULO, KIND_SUM2
    n = 512
    g, o = B200Renderer(), OracleRenderer()
    for r in (g, o):
        # a program structure no other test compiles (cubins are cached by structure process-wide: a cache hit loads at
        # once and there would be no background compile to watch)
        b = build_cfg1_graph(r, delay=100.0)
        x = b.node(KIND_MODULO, b.node(KIND_DIVIDE, b.node(KIND_MINIMUM, b.input(0), b.const(0.3)), b.const(0.7)), b.const(0.11))
        b.output(2, b.node(KIND_SUM2, b.node(KIND_MINIMUM, x, b.input(0)), b.node(KIND_DIVIDE, x, b.const(1.5))))
    k, deadline, switched_at = 0, time.time() + 60.0, None
    while time.time() < deadline:
        blk = [cfg1_input(n * (k + 1))[k * n:]]
        assert_same_bits(g.fill_buffer(3, n, k * n, blk), o.fill_buffer(3, n, k * n, blk), f"block {k}")
        k += 1
        s = g.stats()
        if s["jit_launches"] > 0 and switched_at is None:
            switched_at = k
        if switched_at is not None and k >= switched_at + 4:
            break
        if k > 12:
            time.sleep(0.05)
    s = g.stats()
    assert switched_at is not None and switched_at > 4, (switched_at, s)
    assert 0 < s["jit_launches"] < s["interp_launches"]


def _const_chain(r, n, slot_in=0, slot_out=0, first_handle=2, c=0x3F800000):
    """out[slot_out] = (...((in[slot_in] + c) + c) ... + c), n Sum2 nodes; constants from node 1"""
    from libfriendship_b200 import KIND_SUM2
    prev = None
    h = first_handle
    for i in range(n):
        r.on_add_node(h, KIND_SUM2)
        r.on_add_edge((0, h, slot_in, 0) if prev is None else (prev, h, 0, 0))
        r.on_add_edge((1, h, c, 1))
        prev = h
        h += 1
    r.on_add_edge((prev, 0, 0, slot_out))
    return h


def test_code_size_depends_on_structure_not_on_length():
    """NVRTC's time grows faster than linearly in straight-line code (500 statements: 16 s, 2,000: more than 5 minutes),
    so the generator folds runs of like instruction groups into loops whose trip counts and operands come from a table
    (csrc/jit.cc): what the compiler sees — `jit_code_instructions`, bounded by FRB_JIT_MAX_CODE — is the program's
    structure.  Strands of one structure share a body, whatever their length."""
    import time
    from libfriendship_b200 import B200Renderer, KIND_F32CONSTANT
    # 1) one long chain: a load, ONE loop, a store
    r = B200Renderer(device=-1)
    r.on_add_node(1, KIND_F32CONSTANT)
    _const_chain(r, 3000)
    assert r.jit_code_instructions(1, 0) <= 16
    src = r.jit_source(1, 0)
    assert src.count("const unsigned n_") == 1 and "3000" not in src[src.index("frb_stage"):]
    t0 = time.time()
    assert r.jit_cubin_size(1, 0) > 1000
    first = time.time() - t0
    # 2) another length, other constants: the same source text, so the cubin comes from the cache
    r2 = B200Renderer(device=-1)
    r2.on_add_node(1, KIND_F32CONSTANT)
    _const_chain(r2, 9000, c=0x40000000)
    assert r2.jit_source(1, 0) == src
    t0 = time.time()
    assert r2.jit_cubin_size(1, 0) > 1000
    assert time.time() - t0 < max(0.1, first / 2)
    # 3) chains of different lengths on 48 slots: 48 strands, one body
    r = B200Renderer(device=-1)
    r.on_add_node(1, KIND_F32CONSTANT)
    h = 2
    for slot in range(48):
        h = _const_chain(r, 20 + slot, slot_in=slot, slot_out=slot, first_handle=h, c=0x3F800000 + slot)
    one = B200Renderer(device=-1)
    one.on_add_node(1, KIND_F32CONSTANT)
    _const_chain(one, 20)
    assert r.jit_code_instructions(48, 0) == one.jit_code_instructions(1, 0) <= 16
    assert r.jit_cubin_size(48, 0) > 1000


def test_repeated_groups_fold_on_two_levels():
    """A chain per voice, voices summed: the inner chains are loops, and the run of like (loop + glue) groups is a loop
    again — the operand cursor walks the table, so inner trip counts may differ per voice."""
    from workloads.graphs import GraphBuilder
    from libfriendship_b200 import B200Renderer, KIND_MULTIPLY, KIND_SUM2
    r = B200Renderer(device=-1)
    g = GraphBuilder(r)
    total, slot = None, 0
    for v in range(24):
        voice = None
        for p in range(12 + (v % 5)):
            term = g.node(KIND_MULTIPLY, g.input(slot), g.const(1.0 / (1 + p)))
            slot += 1
            voice = term if voice is None else g.node(KIND_SUM2, voice, term)
        voice = g.node(KIND_MULTIPLY, voice, g.const(0.5 + v))
        total = voice if total is None else g.node(KIND_SUM2, total, voice)
    g.output(0, total)
    src = r.jit_source(1, 0)
    body = src[src.index("frb_stage"):]
    assert body.count("const unsigned n_") >= 2 and "#pragma unroll 1\n" in body       # an outer loop around an inner one
    assert "const float4 h" in body                                                     # inner loop: loads first, then arithmetic
    assert r.jit_code_instructions(1, 0) < 120                                          # ~1,000 instructions in the program
    assert r.jit_cubin_size(1, 0) > 1000


def test_long_aperiodic_program_is_cut_into_chunk_functions():
    """What folding cannot shorten (no repetition) is compiled in __noinline__ chunks of 64 statements with the live
    registers handed over through local memory: NVRTC's time stays linear (400 statements: 2 s instead of 8)."""
    import random
    import time
    from libfriendship_b200 import B200Renderer, KIND_MINIMUM, KIND_MULTIPLY, KIND_SUM2
    from workloads.graphs import GraphBuilder
    random.seed(77)
    r = B200Renderer(device=-1)
    g = GraphBuilder(r)
    x = g.input(0)
    for _ in range(420):
        x = g.node(random.choice([KIND_SUM2, KIND_MULTIPLY, KIND_MINIMUM, KIND_SUM2]), x, g.input(random.randrange(1, 4)))
    g.output(0, x)
    assert 400 < r.jit_code_instructions(1, 0) < 4096
    src = r.jit_source(1, 0)
    assert src.count("__noinline__ void frb_c0_") >= 6 and "float4 R[" in src
    t0 = time.time()
    assert r.jit_cubin_size(1, 0) > 1000
    assert time.time() - t0 < 30.0
