"""Replays a transcribed reference test (tests/golden/reference_tests.json) against a renderer object."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_tests.json")


def load_golden():
    with open(GOLDEN) as f:
        return json.load(f)["tests"]


def replay(r, test):
    for step in test["steps"]:
        op = step["op"]
        if op == "add_node":
            r.on_add_node(step["handle"], step["kind"], step.get("key", 0))
        elif op == "add_edge":
            r.on_add_edge(tuple(step["edge"]))
        elif op == "del_edge":
            r.on_del_edge(tuple(step["edge"]))
        elif op == "del_node":
            r.on_del_node(step["handle"])
        elif op == "define_effect":
            r.define_effect(step["key"], [tuple(n) for n in step["nodes"]], [tuple(e) for e in step["edges"]])
        elif op == "render":
            out = r.fill_buffer(step["n_slots"], step["n_times"], step["idx"], step["inputs"])
            expect = np.array(step["expect_bits"], dtype=np.uint32)
            got = out.view(np.uint32)
            assert np.array_equal(got, expect), (test["name"], test["source"], out, expect.view(np.float32))
        else:
            raise ValueError(op)


def assert_same_bits(a, b, what=""):
    """Bit-exact f32 equality; any NaN equals any NaN (x86 and CUDA produce different NaN payloads, and the
    reference's own assert_eq! on f32 arrays cannot pin a payload)."""
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    both_nan = np.isnan(a) & np.isnan(b)
    same = (a.view(np.uint32) == b.view(np.uint32)) | both_nan
    if not same.all():
        bad = np.argwhere(~same)
        i = tuple(bad[0])
        raise AssertionError(f"{what}: {len(bad)} mismatches, first at {i}: {a[i]!r} ({a.view(np.uint32)[i]:#x}) vs "
                             f"{b[i]!r} ({b.view(np.uint32)[i]:#x})")
