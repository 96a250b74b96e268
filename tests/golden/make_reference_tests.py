"""Transcribes the reference's 11 integration tests into tests/golden/reference_tests.json.

The reference is a Rust crate that cannot be built in this image (no rustc/cargo; it needs 2017 nightly + LLVM 3.8),
so these vectors are hand-transcribed message by message from the test sources; every entry cites file:line.
The expected arrays are the literal `assert_eq!(rendered, array![[...]])` operands (bit-exact f32 equality).

Steps:  add_node(handle, kind) / add_edge(from, to, from_slot, to_slot) / define_effect(key, nodes, edges)
        / render(idx, n_times, n_slots, inputs) -> expect
Kinds follow include/friendship_b200.h (0 Delay, 1 F32Constant, 2 Sum2, 3 Multiply, 4 Divide, 5 Modulo, 6 Minimum, 16 nested effect).
"""
import json
import os
import struct


def bits(x):
    return struct.unpack("<I", struct.pack("<f", x))[0]


DELAY, CONST, SUM2, MUL, DIV, MOD, MIN, EFFECT = 0, 1, 2, 3, 4, 5, 6, 16


def binary_test(name, lines, kind, a, b, expect):
    # tests/render_prim.rs: node 1 = the primitive -> out0, node 2 = C(a) -> slot 0, node 3 = C(b) -> slot 1
    return {
        "name": name, "source": f"tests/render_prim.rs:{lines}",
        "steps": [
            {"op": "add_node", "handle": 1, "kind": kind},
            {"op": "add_edge", "edge": [1, 0, 0, 0]},
            {"op": "add_node", "handle": 2, "kind": CONST},
            {"op": "add_edge", "edge": [2, 1, bits(a), 0]},
            {"op": "add_node", "handle": 3, "kind": CONST},
            {"op": "add_edge", "edge": [3, 1, bits(b), 1]},
            {"op": "render", "idx": 0, "n_times": 4, "n_slots": 1, "inputs": [], "expect_bits": [[bits(expect)] * 4]},
        ],
    }


def f32_div(a, b):
    import numpy as np
    return float(np.float32(a) / np.float32(b))


tests = [
    {"name": "render_zeros", "source": "tests/render_prim.rs:70-80", "steps": [
        {"op": "render", "idx": 0, "n_times": 4, "n_slots": 1, "inputs": [], "expect_bits": [[0, 0, 0, 0]]}]},
    {"name": "render_const", "source": "tests/render_prim.rs:83-98", "steps": [
        {"op": "add_node", "handle": 1, "kind": CONST},
        {"op": "add_edge", "edge": [1, 0, bits(0.5), 0]},
        {"op": "render", "idx": 0, "n_times": 4, "n_slots": 1, "inputs": [], "expect_bits": [[bits(0.5)] * 4]}]},
    {"name": "render_delay", "source": "tests/render_prim.rs:101-129", "steps": [
        {"op": "add_node", "handle": 1, "kind": DELAY},
        {"op": "add_edge", "edge": [1, 0, 0, 0]},
        {"op": "add_node", "handle": 2, "kind": CONST},
        {"op": "add_edge", "edge": [2, 1, bits(0.5), 0]},
        {"op": "add_node", "handle": 3, "kind": CONST},
        {"op": "add_edge", "edge": [3, 1, bits(2.0), 1]},
        {"op": "render", "idx": 0, "n_times": 4, "n_slots": 1, "inputs": [],
         "expect_bits": [[0, 0, bits(0.5), bits(0.5)]]}]},
    binary_test("render_mult", "132-162", MUL, 0.5, -3.0, -1.5),
    binary_test("render_sum2", "165-195", SUM2, 0.5, -3.0, -2.5),
    binary_test("render_div", "198-227", DIV, 0.5, -3.0, f32_div(0.5, -3.0)),   # 0.5f32/-3f32 == 0xBE2AAAAB
    binary_test("render_mod", "230-259", MOD, -3.5, 2.0, 0.5),
    binary_test("render_min", "262-291", MIN, -3.5, 2.0, -3.5),
    {"name": "render_passthrough", "source": "tests/ext_input.rs:47-81", "steps": [
        {"op": "add_edge", "edge": [0, 0, 0, 0]},                                        # Edge::new_to_null(toplevel, (0,0))
        {"op": "render", "idx": 0, "n_times": 4, "n_slots": 1, "inputs": [[1.0, 2.0, 3.0, 4.0]],
         "expect_bits": [[bits(1.0), bits(2.0), bits(3.0), bits(4.0)]]},
        {"op": "render", "idx": 4, "n_times": 4, "n_slots": 1, "inputs": [[0.0, 1.0, 2.0]],   # pad with last value
         "expect_bits": [[bits(0.0), bits(1.0), bits(2.0), bits(2.0)]]},
        {"op": "render", "idx": 0, "n_times": 4, "n_slots": 1, "inputs": [],                  # seek zeroes the inputs
         "expect_bits": [[0, 0, 0, 0]]}]},
    {"name": "ext_render_delay", "source": "tests/ext_input.rs:84-122", "steps": [
        {"op": "add_node", "handle": 1, "kind": DELAY},
        {"op": "add_edge", "edge": [1, 0, 0, 0]},
        {"op": "add_edge", "edge": [0, 1, 0, 0]},                                        # Edge::new_from_null(delay, (0,0))
        {"op": "render", "idx": 0, "n_times": 4, "n_slots": 1, "inputs": [[1.0, 2.0, 3.0, 4.0]],
         "expect_bits": [[bits(1.0), bits(2.0), bits(3.0), bits(4.0)]]},
        {"op": "add_node", "handle": 2, "kind": CONST},
        {"op": "add_edge", "edge": [2, 1, bits(1.0), 1]},                                # graph edit between calls
        {"op": "render", "idx": 4, "n_times": 4, "n_slots": 1, "inputs": [[1.0, 2.0, 3.0, 4.0]],
         "expect_bits": [[bits(4.0), bits(1.0), bits(2.0), bits(3.0)]]}]},
    {"name": "load_multby2", "source": "tests/load_effect.rs:42-65,68-112", "steps": [
        # effect "MulBy2": Multiply(in0, C(5.0)) -> out0   (tests/load_effect.rs:42-65)
        {"op": "define_effect", "key": 1, "nodes": [[1, MUL, 0], [2, CONST, 0]],
         "edges": [[0, 1, 0, 0], [1, 0, 0, 0], [2, 1, bits(5.0), 1]]},
        {"op": "add_node", "handle": 1, "kind": EFFECT, "key": 1},
        {"op": "add_edge", "edge": [1, 0, 0, 0]},
        {"op": "add_node", "handle": 2, "kind": CONST},
        {"op": "add_edge", "edge": [2, 1, bits(0.5), 0]},
        {"op": "render", "idx": 0, "n_times": 4, "n_slots": 1, "inputs": [], "expect_bits": [[bits(2.5)] * 4]}]},
]

if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_tests.json")
    with open(out, "w") as f:
        json.dump({"comment": "transcribed from /root/reference/tests/*.rs by make_reference_tests.py", "tests": tests}, f, indent=1)
    print("wrote", out, len(tests), "tests")
