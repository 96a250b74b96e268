"""Graph builders shared by tests, smoke() and bench.py.  They drive any object with the renderer interface
(on_add_node / on_add_edge / define_* / fill_buffer): the B200 renderer or the CPU oracle."""
import struct

import numpy as np

from .kinds import (KIND_DELAY, KIND_DIVIDE, KIND_F32CONSTANT, KIND_MINIMUM, KIND_MODULO, KIND_MULTIPLY,
                                KIND_SUM2)


def f32_bits(x):
    """f32::to_bits — how tests encode F32Constant values in from_slot (reference tests/render_prim.rs:88)."""
    return struct.unpack("<I", struct.pack("<f", float(np.float32(x))))[0]


class GraphBuilder:
    """Allocates handles and wires primitives; constants are one shared F32Constant node (handle 1) whose value
    is selected by from_slot, exactly as the reference's tests do."""

    def __init__(self, r):
        self.r = r
        self.next = 2
        r.on_add_node(1, KIND_F32CONSTANT)

    def const(self, x):
        return (1, f32_bits(x))

    def input(self, slot):
        return (0, slot)

    def node(self, kind, *ins):
        h = self.next
        self.next += 1
        self.r.on_add_node(h, kind)
        for to_slot, src in enumerate(ins):
            if src is None:
                continue
            self.r.on_add_edge((src[0], h, src[1], to_slot))
        return (h, 0)

    def output(self, slot, src):
        self.r.on_add_edge((src[0], 0, src[1], slot))


def cfg1_input(n, sr=48000.0, f=440.0):
    t = np.arange(n, dtype=np.float64)
    return np.sin(2.0 * np.pi * f * t / sr).astype(np.float32)


def build_cfg1_graph(r, delay=12000.0):
    """BASELINE configs[0] / SURVEY.md §8d cfg1: 440 Hz sine (external input 0) through Multiply/Sum/Delay,
    plus a Minimum/Modulo/Divide side chain on output slot 1."""
    g = GraphBuilder(r)
    x = g.input(0)
    gain = g.node(KIND_MULTIPLY, x, g.const(0.5))
    d = g.node(KIND_DELAY, gain, g.const(delay))
    e = g.node(KIND_MULTIPLY, d, g.const(0.35))
    g.output(0, g.node(KIND_SUM2, gain, e))
    m = g.node(KIND_MINIMUM, x, g.const(0.25))
    mo = g.node(KIND_MODULO, m, g.const(0.1))
    g.output(1, g.node(KIND_DIVIDE, mo, g.const(3.0)))
    return g
