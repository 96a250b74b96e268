"""Stage programs longer than the 48 KB the interpreter kernel stages in shared memory (interp.cu, launch_interp:
`prog_in_smem`): the program is then read from global memory, one broadcast load per interpreted instruction.
The graphs of the other GPU tests are tens to hundreds of instructions, so this file is the only one on that path
(DESIGN.md §9).  Named test_zz_* so that it runs after every other parity test.

The graph is the Sum2 chain of tests/test_flatten.py: out0 = (...((in0 + 1) + 1) ... + 1), reference semantics of
Sum2 = one f32 addition per node (reference src/render/reference.rs:232-238), so numpy's float32 additions in the
same order round the same way: bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _chain(n):
    from libfriendship_b200 import B200Renderer, FLAG_NO_JIT, KIND_F32CONSTANT, KIND_SUM2
    r = B200Renderer(flags=FLAG_NO_JIT)        # the interpreter is what is under test
    r.on_add_node(1, KIND_F32CONSTANT)
    prev = 0
    for i in range(n):
        h = 2 + i
        r.on_add_node(h, KIND_SUM2)
        r.on_add_edge((prev, h, 0, 0))
        r.on_add_edge((1, h, 0x3F800000, 1))          # + 1.0f
        prev = h
    r.on_add_edge((prev, 0, 0, 0))
    return r


def _want(x, n):
    want = x.copy()
    one = np.float32(1.0)
    for _ in range(n):
        want = want + one
    return want


@pytest.mark.parametrize("n_nodes,n_times", [(4000, 4096), (4000, 1001), (9000, 520)])
def test_program_read_from_global_memory_is_bit_exact(n_nodes, n_times):
    # (n_nodes + 3) x 16 B > 48 KB from 3,070 nodes on
    assert (n_nodes + 3) * 16 > 48 * 1024
    r = _chain(n_nodes)
    rng = np.random.Generator(np.random.PCG64(n_nodes + n_times))
    x = rng.uniform(-1, 1, 2 * n_times).astype(np.float32)
    # two consecutive blocks (the second starts at an odd time for n_times = 1001)
    got0 = r.fill_buffer(1, n_times, 0, [x[:n_times]])
    got1 = r.fill_buffer(1, n_times, n_times, [x[n_times:]])
    st = r.stats()
    assert st["interp_launches"] >= 2 and st["jit_launches"] == 0
    want = _want(x, n_nodes)
    assert np.array_equal(got0[0].view(np.uint32), want[:n_times].view(np.uint32))
    assert np.array_equal(got1[0].view(np.uint32), want[n_times:].view(np.uint32))
