"""Seeded random effect graphs for differential tests (oracle vs CUDA).  Pure construction logic: the same
sequence of GraphWatcher calls is applied to any renderer object."""
import numpy as np

from workloads.graphs import f32_bits
from libfriendship_b200 import (KIND_DELAY, KIND_DIVIDE, KIND_EFFECT, KIND_F32CONSTANT, KIND_MINIMUM, KIND_MODULO,
                                KIND_MULTIPLY, KIND_SUM2)

BINARY = [KIND_SUM2, KIND_MULTIPLY, KIND_DIVIDE, KIND_MODULO, KIND_MINIMUM]
SPECIAL = [0.0, -0.0, 1.0, -1.0, 0.5, 2.0, 3.0, -2.5, 7.25, 1e-3, 100.0, float("inf"), float("-inf"), float("nan"),
           1e30, -1e30, 1e-40, 1.8446744e19, 3.9999, 12.0]


class Recorder:
    """Records GraphWatcher / definition calls so that they can be replayed on several renderers."""

    def __init__(self):
        self.calls = []

    def __getattr__(self, name):
        def f(*a):
            self.calls.append((name, a))
        return f

    def apply(self, r):
        for name, a in self.calls:
            getattr(r, name)(*a)


def random_body(rng, rec_nodes, rec_edges, n_inputs, n_nodes, n_outputs, effect_keys, allow_delay=True,
                signal_delay_prob=0.3, const_pool=None):
    """Builds a random DAG body.  rec_nodes(handle, kind, key) / rec_edges(edge) receive the structure.
    Returns nothing; sources are node outputs, constants and graph inputs."""
    const_pool = const_pool or SPECIAL
    CONST_H = 1
    rec_nodes(CONST_H, KIND_F32CONSTANT, 0)
    sources = [(0, s) for s in range(n_inputs)]          # (handle, from_slot)
    def pick(prefer_signal=False):
        r = rng.rand()
        if sources and (prefer_signal or r < 0.65):
            return sources[rng.randint(len(sources))]
        if r < 0.97:
            c = const_pool[rng.randint(len(const_pool))] if rng.rand() < 0.5 else float(np.float32(rng.randn() * 4))
            return (CONST_H, f32_bits(c))
        return None                                      # unconnected input
    h = 2
    for _ in range(n_nodes):
        r = rng.rand()
        if allow_delay and r < 0.25:
            rec_nodes(h, KIND_DELAY, 0)
            src = pick(prefer_signal=True)
            if rng.rand() < signal_delay_prob:
                amt = pick()
            else:
                d = [0.0, 1.0, 2.0, 3.0, 5.0, 17.0, 64.0, 2.5, 100.0, -1.0][rng.randint(10)]
                amt = (CONST_H, f32_bits(d))
            for to_slot, s in enumerate((src, amt)):
                if s is not None:
                    rec_edges((s[0], h, s[1], to_slot))
            sources.append((h, 0))
        elif effect_keys and r < 0.40:
            key, n_in, n_out = effect_keys[rng.randint(len(effect_keys))]
            rec_nodes(h, KIND_EFFECT, key)
            for to_slot in range(n_in):
                s = pick()
                if s is not None:
                    rec_edges((s[0], h, s[1], to_slot))
            for o in range(n_out):
                sources.append((h, o))
        else:
            kind = BINARY[rng.randint(len(BINARY))]
            rec_nodes(h, kind, 0)
            for to_slot in range(2):
                s = pick()
                if s is not None:
                    rec_edges((s[0], h, s[1], to_slot))
            sources.append((h, 0))
        h += 1
    for o in range(n_outputs):
        if rng.rand() < 0.9:
            s = sources[rng.randint(len(sources))] if rng.rand() < 0.9 else pick()
            if s is not None:
                rec_edges((s[0], 0, s[1], o))
    return h


def random_graph(seed, n_inputs=2, n_nodes=12, n_outputs=2, nested_levels=1, allow_delay=True,
                 signal_delay_prob=0.3, const_pool=None):
    """Returns a Recorder holding definitions + the top-level graph."""
    rng = np.random.RandomState(seed)
    rec = Recorder()
    effect_keys = []
    key = 100
    for level in range(nested_levels):
        for _ in range(2):
            nodes, edges = [], []
            n_in, n_out = rng.randint(1, 3), rng.randint(1, 3)
            random_body(rng, lambda h, k, ky: nodes.append((h, k, ky)), edges.append, n_in, rng.randint(2, 6), n_out,
                        list(effect_keys) if level > 0 else [], allow_delay, signal_delay_prob, const_pool)
            rec.define_effect(key, nodes, edges)
            effect_keys.append((key, n_in, n_out))
            key += 1
    random_body(rng, lambda h, k, ky: rec.on_add_node(h, k, ky), rec.on_add_edge, n_inputs, n_nodes, n_outputs,
                effect_keys, allow_delay, signal_delay_prob, const_pool)
    return rec


def random_inputs(rng, n_rows, n_times, ragged=True):
    rows = []
    for _ in range(n_rows):
        ln = rng.randint(0, n_times + 1) if ragged and rng.rand() < 0.4 else n_times
        rows.append((rng.randn(ln) * 3).astype(np.float32))
    return rows
