"""CPU (host logic; planning-only renderer, no compute): the control path around the hot path — RouteGraph
validation errors, effect loading through ResMan (JSON wire format, SHA-256), query callbacks — restated from
reference src/dispatch.rs, src/routing/routegraph.rs, src/routing/effect.rs, src/resman.rs."""
import hashlib
import json

import pytest

from libfriendship_b200.dispatch import Client, Dispatch, DispatchError, EffectId, sha256_file


def bits(x):
    import struct
    return struct.unpack("<I", struct.pack("<f", x))[0]


def planner(client=None):
    return Dispatch(client, device=-1)


def mulby_desc(factor=5.0, name="MulBy2"):
    """The EffectDesc tests/load_effect.rs:42-65 builds (serde JSON)."""
    h = lambda n: {"node_handle": n}
    return {"meta": {"id": {"name": name, "sha256": None, "urls": []},
                     "inputs": [{"name": "source", "channel": 0}], "outputs": [{"name": "result", "channel": 0}]},
            "adjlist": {"nodes": [[h(1), {"name": "Multiply", "sha256": None, "urls": ["primitive:///Multiply"]}],
                                  [h(2), {"name": "Constant", "sha256": None, "urls": ["primitive:///F32Constant"]}]],
                        "edges": [{"from": h(0), "to": h(1), "weight": {"from_slot": 0, "to_slot": 0}},
                                  {"from": h(1), "to": h(0), "weight": {"from_slot": 0, "to_slot": 0}},
                                  {"from": h(2), "to": h(1), "weight": {"from_slot": bits(factor), "to_slot": 1}}]}}


def test_sha256_known_answers(tmp_path):
    for payload in (b"", b"abc", b"abcdbcdecdefdefgefghfghighijhijkijkljklmklmnlmnomnopnopq", b"x" * 1000):
        p = tmp_path / "f.bin"
        p.write_bytes(payload)
        assert sha256_file(p) == hashlib.sha256(payload).digest()


def test_routegraph_errors():
    d = planner()
    d.add_node(1, EffectId.primitive("Delay"))
    with pytest.raises(DispatchError) as e:
        d.add_node(1, EffectId.primitive("Sum2"))
    assert e.value.variant == "NodeExists"                                   # routegraph.rs:156
    with pytest.raises(DispatchError) as e:
        d.add_edge((1, 7, 0, 0))
    assert e.value.variant == "NoSuchNode"                                   # routegraph.rs:171
    with pytest.raises(DispatchError) as e:
        d.add_edge((9, 1, 0, 0))
    assert e.value.variant == "NoSuchNode"                                   # routegraph.rs:188
    with pytest.raises(DispatchError) as e:
        d.add_edge((0, 1, 0, 2))                                             # Delay has inputs 0 (source) and 1 (frames)
    assert e.value.variant == "NoSuchSlot"
    with pytest.raises(DispatchError) as e:
        d.add_edge((1, 0, 1, 0))                                             # Delay has one output
    assert e.value.variant == "NoSuchSlot"
    d.add_edge((0, 1, 0, 0))
    with pytest.raises(DispatchError) as e:
        d.add_edge((0, 1, 3, 0))
    assert e.value.variant == "SlotAlreadyConnected"                         # routegraph.rs:176
    with pytest.raises(DispatchError) as e:
        d.del_node(1)
    assert e.value.variant == "NodeInUse"                                    # routegraph.rs:273
    d.del_edge((0, 1, 0, 0))
    d.del_node(1)
    d.del_node(1)                                                            # already deleted: Ok (routegraph.rs:265)
    d.add_node(2, EffectId.primitive("F32Constant"))
    d.add_edge((2, 0, 0xFFFFFFFE, 0))                                        # every u32 but the last is an output
    with pytest.raises(DispatchError) as e:
        d.add_edge((2, 0, 0xFFFFFFFF, 1))
    assert e.value.variant == "NoSuchSlot"                                   # effect.rs:390-393: range 0..u32::MAX


def test_cycles_are_rejected():
    d = planner()
    for h in (1, 2, 3):
        d.add_node(h, EffectId.primitive("Sum2"))
    d.add_edge((1, 2, 0, 0))
    d.add_edge((2, 3, 0, 0))
    with pytest.raises(DispatchError) as e:
        d.add_edge((3, 1, 0, 0))
    assert e.value.variant == "WouldCycle"
    with pytest.raises(DispatchError) as e:
        d.add_edge((1, 1, 0, 1))
    assert e.value.variant == "WouldCycle"
    d.add_edge((3, 0, 0, 0))                                                 # outputs never close a loop
    d.add_edge((0, 1, 0, 0))


def test_cycle_check_sees_through_nested_effects(tmp_path):
    """A two-input / two-output effect whose input 0 only reaches output 0 and input 1 only output 1: feeding its
    output 1 back into its own input 0 is legal, output 0 back into input 0 is a cycle (routegraph.rs:239-262)."""
    h = lambda n: {"node_handle": n}
    prim = lambda n: {"name": n, "sha256": None, "urls": [f"primitive:///{n}"]}
    desc = {"meta": {"id": {"name": "TwoLanes", "sha256": None, "urls": []},
                     "inputs": [{"name": "a", "channel": 0}, {"name": "b", "channel": 0}],
                     "outputs": [{"name": "x", "channel": 0}, {"name": "y", "channel": 0}]},
            "adjlist": {"nodes": [[h(1), prim("Multiply")], [h(2), prim("Multiply")], [h(3), prim("F32Constant")]],
                        "edges": [{"from": h(0), "to": h(1), "weight": {"from_slot": 0, "to_slot": 0}},
                                  {"from": h(3), "to": h(1), "weight": {"from_slot": bits(2.0), "to_slot": 1}},
                                  {"from": h(1), "to": h(0), "weight": {"from_slot": 0, "to_slot": 0}},
                                  {"from": h(0), "to": h(2), "weight": {"from_slot": 1, "to_slot": 0}},
                                  {"from": h(3), "to": h(2), "weight": {"from_slot": bits(3.0), "to_slot": 1}},
                                  {"from": h(2), "to": h(0), "weight": {"from_slot": 0, "to_slot": 1}}]}}
    (tmp_path / "two.fnd").write_text(json.dumps(desc))
    d = planner()
    d.add_dir(tmp_path)
    d.add_node(1, EffectId("TwoLanes"))
    d.add_edge((1, 1, 1, 0))                                                 # y -> a : lanes are independent, no cycle
    with pytest.raises(DispatchError) as e:
        d.add_edge((1, 1, 1, 1))                                             # y -> b : b feeds y
    assert e.value.variant == "WouldCycle"


def test_effect_loading_by_name_and_hash(tmp_path):
    path = tmp_path / "mulby2.fnd"
    path.write_text(json.dumps(mulby_desc()))
    (tmp_path / "garbage.fnd").write_text("{not json")
    (tmp_path / "other.fnd").write_text(json.dumps(mulby_desc(7.0, name="MulBy7")))
    sha = sha256_file(path)
    assert sha == hashlib.sha256(path.read_bytes()).digest()

    got = {}

    class C(Client):
        def node_meta(self, handle, meta):
            got["meta"] = (handle, meta)

        def node_id(self, handle, id):
            got["id"] = (handle, id)

    d = planner(C())
    with pytest.raises(DispatchError) as e:
        d.add_node(1, EffectId("MulBy2", sha, []))                            # no search dir yet
    assert e.value.variant == "NoMatchingEffect"
    d.add_dir(tmp_path)
    d.add_node(1, EffectId("MulBy2", sha, []))                                # tests/load_effect.rs:90-93
    d.add_node(2, EffectId("MulBy7"))                                         # by name only
    with pytest.raises(DispatchError) as e:
        d.add_node(3, EffectId("MulBy2", bytes(32), []))                      # wrong hash
    assert e.value.variant == "NoMatchingEffect"
    with pytest.raises(DispatchError) as e:
        d.add_node(3, EffectId("Delay", bytes(32), ["primitive:///Delay"]))   # primitive with a hash: not a primitive (effect.rs:152-154)
    assert e.value.variant == "NoMatchingEffect"
    # slot validation uses the loaded metadata: MulBy2 has exactly one input and one output
    with pytest.raises(DispatchError) as e:
        d.add_edge((0, 1, 0, 1))
    assert e.value.variant == "NoSuchSlot"
    d.add_edge((1, 0, 0, 0))
    d.query_meta(1)
    d.query_id(1)
    d.query_meta(99)                                                          # unknown handle: warning only
    assert got["meta"][0] == 1 and got["meta"][1]["id"]["name"] == "MulBy2"
    assert [i["name"] for i in got["meta"][1]["inputs"]] == ["source"]
    assert got["id"][1]["name"] == "MulBy2" and len(got["id"][1]["sha256"]) == 32
    adj = d.adjlist()
    assert sorted(n[0]["node_handle"] for n in adj["nodes"]) == [1, 2]
    assert {"from": {"node_handle": 1}, "to": {"node_handle": 0}, "weight": {"from_slot": 0, "to_slot": 0}} in adj["edges"]


def test_effect_metadata_must_agree_with_its_graph(tmp_path):
    bad = mulby_desc()
    bad["meta"]["outputs"].append({"name": "second", "channel": 0})           # declared output 1 is not driven (effect.rs:168-175)
    (tmp_path / "bad.fnd").write_text(json.dumps(bad))
    bad2 = mulby_desc(name="Undriven")
    bad2["adjlist"]["edges"].pop()                                            # Multiply input 1 not driven (effect.rs:189-195)
    (tmp_path / "bad2.fnd").write_text(json.dumps(bad2))
    bad3 = mulby_desc(name="UndeclaredInput")
    bad3["adjlist"]["edges"][0]["weight"]["from_slot"] = 3                    # reads input 3, only 1 declared (effect.rs:179-188)
    (tmp_path / "bad3.fnd").write_text(json.dumps(bad3))
    d = planner()
    d.add_dir(tmp_path)
    for name in ("MulBy2", "Undriven", "UndeclaredInput"):
        with pytest.raises(DispatchError) as e:
            d.add_node(1, EffectId(name))
        assert e.value.variant == "NoMatchingEffect"


def test_render_without_a_device_fails_loudly():
    from libfriendship_b200 import RendererError
    d = planner()
    with pytest.raises(DispatchError) as e:
        d.render_range(0, 4, 1)
    assert e.value.code == -9


def test_messages_by_osc_address(tmp_path):
    """`Dispatch::dispatch(OscToplevel)` (dispatch.rs:109-160): every message through the one entry point, by the
    address its `#[osc_address]` attributes spell (dispatch.rs:31-84)."""
    (tmp_path / "mulby2.fnd").write_text(json.dumps(mulby_desc()))
    got = []

    class C(Client):
        def node_meta(self, handle, meta):
            got.append(("meta", handle, meta["id"]["name"]))

        def node_id(self, handle, id):
            got.append(("id", handle, id["name"]))

    d = planner(C())
    d.dispatch("/resman/add_dir", tmp_path)
    d.dispatch("/routegraph/add_node", 1, EffectId("MulBy2"))
    d.dispatch("/routegraph/add_node", 2, EffectId.primitive("Delay"))
    d.dispatch("/routegraph/add_edge", (0, 1, 0, 0))
    d.dispatch("routegraph/add_edge/", (1, 2, 0, 0))                          # leading / trailing separators do not matter
    d.dispatch("/routegraph/query_meta", 1)
    d.dispatch("/routegraph/query_id", 2)
    assert got == [("meta", 1, "MulBy2"), ("id", 2, "Delay")]
    with pytest.raises(DispatchError) as e:
        d.dispatch("/routegraph/del_node", 1)
    assert e.value.variant == "NodeInUse"
    d.dispatch("/routegraph/del_edge", (1, 2, 0, 0))
    d.dispatch("/routegraph/del_edge", (0, 1, 0, 0))
    d.dispatch("/routegraph/del_node", 1)
    assert [n[0]["node_handle"] for n in d.adjlist()["nodes"]] == [2]
    with pytest.raises(DispatchError) as e:
        d.dispatch("/renderer/render", range(0, 4), 1, None)                  # planning only: refused, not emulated
    assert e.value.code == -9
    with pytest.raises(DispatchError) as e:
        d.dispatch("/renderer/stop")
    assert e.value.variant == "BadMessage"


def test_hostile_effect_files_are_rejected_not_crashed_on(tmp_path):
    """Effect files come from disk (resman.rs:64-97): nesting deeper than serde_json's recursion limit, numbers out of
    range, and effects that contain themselves (directly, or through another file, by name only) end up as
    NoMatchingEffect — the reference's parser errors out on the first two and overflows its stack on the third."""
    d = planner()
    d.add_dir(tmp_path)
    hostile = {"deep_arr": "[" * 200000, "deep_obj": '{"a":' * 100000, "big_num": '{"meta": 1e999999}',
               "huge_handle": json.dumps(mulby_desc()).replace('"node_handle": 2', '"node_handle": 99999999999999999999')}
    for name, payload in hostile.items():
        f = tmp_path / "a.fnd"
        f.write_text(payload)
        with pytest.raises(DispatchError) as e:
            d.add_node(1, EffectId("MulBy2"))
        assert e.value.variant == "NoMatchingEffect", name
        f.unlink()
    nested = lambda name, inner: dict(mulby_desc(name=name), adjlist=dict(
        mulby_desc()["adjlist"], nodes=mulby_desc()["adjlist"]["nodes"] + [[{"node_handle": 3}, {"name": inner, "sha256": None, "urls": []}]]))
    (tmp_path / "a.fnd").write_text(json.dumps(nested("A", "A")))
    (tmp_path / "b.fnd").write_text(json.dumps(nested("B", "C")))
    (tmp_path / "c.fnd").write_text(json.dumps(nested("C", "B")))
    for name in ("A", "B", "C"):
        with pytest.raises(DispatchError) as e:
            d.add_node(1, EffectId(name))
        assert e.value.variant == "NoMatchingEffect", name
    # depth within the limit still loads: 100 nested arrays are not an effect, but they parse (the error is the shape's)
    (tmp_path / "ok.fnd").write_text(json.dumps(mulby_desc(name="Fine")))
    d.add_node(1, EffectId("Fine"))
