"""GPU: the reference's 11 integration tests replayed message for message through Dispatch -> RouteGraph/ResMan ->
B200 renderer -> Client::audio_rendered, including the effect file on disk of tests/load_effect.rs."""
import json

import numpy as np
import pytest

from replay import load_golden

pytestmark = pytest.mark.gpu
KIND_NAME = {0: "Delay", 1: "F32Constant", 2: "Sum2", 3: "Multiply", 4: "Divide", 5: "Modulo", 6: "Minimum"}


def _n_devices():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("via", ["methods", "osc_address", "two_devices"])
@pytest.mark.parametrize("test", load_golden(), ids=lambda t: t["name"])
def test_dispatch_replays_reference_tests(test, via, tmp_path):
    from libfriendship_b200.dispatch import Client, Dispatch, EffectId, sha256_file
    from test_dispatch_cpu import bits  # noqa: F401

    rendered = []

    class MyClient(Client):                                  # tests/render_prim.rs:18-27
        def audio_rendered(self, buffer, idx):
            rendered.append(buffer)

    if via == "two_devices":
        # the same Dispatch with ITS ONE renderer spanning two B200s (frb_config.n_devices, csrc/multi.cu): the reference's
        # graphs use no oscillator-bank lane, so the result must stay bit for bit what one device renders
        if _n_devices() < 2:
            pytest.skip("needs two CUDA devices")
        d = Dispatch(MyClient(), n_devices=2)
    else:
        d = Dispatch(MyClient())
    if via == "osc_address":
        # the same messages through the one entry point, by address (Dispatch::dispatch, dispatch.rs:109-160)
        disp = d

        class ByAddress:
            add_dir = staticmethod(lambda p: disp.dispatch("/resman/add_dir", p))
            add_node = staticmethod(lambda h, eid: disp.dispatch("/routegraph/add_node", h, eid))
            add_edge = staticmethod(lambda e: disp.dispatch("/routegraph/add_edge", e))
            render_range = staticmethod(lambda a, b, n, inputs: disp.dispatch("/renderer/render", range(a, b), n, inputs))
        d = ByAddress
    effect_ids = {}
    for step in test["steps"]:
        op = step["op"]
        if op == "define_effect":
            # tests/load_effect.rs:68-93: serialise the EffectDesc into <tmp>/mulby2.fnd, AddDir, hash the file
            h = lambda n: {"node_handle": n}
            desc = {"meta": {"id": {"name": "MulBy2", "sha256": None, "urls": []},
                             "inputs": [{"name": "source", "channel": 0}], "outputs": [{"name": "result", "channel": 0}]},
                    "adjlist": {"nodes": [[h(n[0]), {"name": KIND_NAME[n[1]], "sha256": None,
                                                     "urls": [f"primitive:///{KIND_NAME[n[1]]}"]}] for n in step["nodes"]],
                                "edges": [{"from": h(e[0]), "to": h(e[1]), "weight": {"from_slot": e[2], "to_slot": e[3]}}
                                          for e in step["edges"]]}}
            path = tmp_path / "mulby2.fnd"
            path.write_text(json.dumps(desc, separators=(",", ":")))
            d.add_dir(tmp_path)
            effect_ids[step["key"]] = EffectId("MulBy2", sha256_file(path), [])
        elif op == "add_node":
            eid = effect_ids[step["key"]] if step["kind"] == 16 else EffectId.primitive(KIND_NAME[step["kind"]])
            d.add_node(step["handle"], eid)
        elif op == "add_edge":
            d.add_edge(tuple(step["edge"]))
        elif op == "render":
            d.render_range(step["idx"], step["idx"] + step["n_times"], step["n_slots"], step["inputs"])
            got = rendered.pop().view(np.uint32)
            assert np.array_equal(got, np.array(step["expect_bits"], dtype=np.uint32)), (test["name"], got)
        else:
            raise ValueError(op)
