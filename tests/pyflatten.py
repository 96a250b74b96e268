"""Independent Python restatement of the flattening rules (test infrastructure): routing order, stage cut,
buffer indexing and delay-line lookbacks, with the same deterministic tie-break as csrc/flatten.cc
(DFS from output slot 0 upward, inbound slot 0 before slot 1; extension lanes numbered together).
Semantics follow reference src/render/reference.rs:158-266."""
import copy
import struct

V_ZERO, V_CONST, V_INPUT, V_DELAY, V_SUM2, V_MUL, V_DIV, V_MOD, V_MIN, V_EXT, V_TAP, V_GATE = range(12)
CLONE_MAX_OPS = 8
KIND_TO_OP = {2: V_SUM2, 3: V_MUL, 4: V_DIV, 5: V_MOD, 6: V_MIN}
FULL = (1 << 64) - 1


class PyGraph:
    """Graph mirror with the renderer interface (GraphWatcher + definitions)."""

    def __init__(self):
        self.nodes = {}          # handle -> dict(kind, key, body, inbound=[edge or None])
        self.outputs = []        # edge or None by to_slot
        self.effects = {}
        self.ext = {}            # (kind, key) -> (lanes, max_delay)

    def define_effect(self, key, nodes, edges):
        g = PyGraph()
        g.effects, g.ext = self.effects, self.ext
        for (h, kind, k) in nodes:
            g.on_add_node(h, kind, k)
        for e in edges:
            g.on_add_edge(e)
        self.effects[key] = g

    def define_oscbank(self, key, sample_rate, voice_offsets, *a, **kw):
        self.ext[(32, key)] = (len(voice_offsets) - 1, 0)

    def define_directform(self, key, b0, *a):
        self.ext[(33, key)] = (len(b0), 0)

    def define_fbdelay(self, key, delay, gain):
        self.ext[(34, key)] = (len(delay), int(max(delay)) if len(delay) else 0)

    def on_add_node(self, handle, kind, key=0):
        body = copy.deepcopy(self.effects[key]) if kind == 16 else None
        self.nodes[handle] = dict(kind=kind, key=key, body=body, inbound=[])

    def on_del_node(self, handle):
        self.nodes.pop(handle, None)

    def _slot(self, e):
        return self.outputs if e[1] == 0 else self.nodes[e[1]]["inbound"]

    def on_add_edge(self, e):
        v = self._slot(e)
        while len(v) <= e[3]:
            v.append(None)
        v[e[3]] = tuple(e)

    def on_del_edge(self, e):
        v = self._slot(e)
        if e[3] < len(v):
            v[e[3]] = None


def const_delay(bits):
    d = struct.unpack("<f", struct.pack("<I", bits))[0]
    if d >= 18446744073709551616.0:
        return None
    if not (d >= 0.0):
        return 0
    return int(d)


def sat_add(a, b):
    if a == FULL or b == FULL:
        return FULL
    r = a + b
    return FULL if r > (1 << 40) else r


def flatten(top, n_slots):
    values, cons, ext = [], {}, []
    edge_memo, ext_memo = {}, {}
    n_inputs = [0]

    def mk(op, a=0, b=0, imm=0):
        k = (op, a, b, imm)
        if k not in cons:
            cons[k] = len(values)
            values.append(k)
        return cons[k]

    def mk64(op, a, v):
        return mk(op, a, v >> 32, v & 0xFFFFFFFF)

    def shift_of(x):
        return (x[2] << 32) | x[3]

    def is_ti(v):
        o, a, b, _ = values[v]
        if o in (V_ZERO, V_CONST):
            return True
        if V_SUM2 <= o <= V_MIN:
            return is_ti(a) and is_ti(b)
        return False

    def const_delay_of(amt):
        """(reads, d) for a constant `frames` value"""
        if values[amt][0] == V_ZERO:
            return 0
        return const_delay(values[amt][3])

    def clone_cost(v, budget):
        o, a, b, _ = values[v]
        if o in (V_ZERO, V_CONST, V_INPUT, V_EXT, V_TAP):
            return 0
        if o == V_GATE:
            return clone_cost(a, budget)
        if V_SUM2 <= o <= V_MIN:
            if budget == 0:
                return None
            ca = clone_cost(a, budget - 1)
            if ca is None:
                return None
            cb = clone_cost(b, budget - 1 - min(ca, budget - 1))
            if cb is None or 1 + ca + cb > budget:
                return None
            return 1 + ca + cb
        if o == V_DELAY:
            if values[b][0] not in (V_CONST, V_ZERO):
                return None
            if values[a][0] not in (V_INPUT, V_EXT) and not is_ti(a):
                return None
            return 0
        return None

    def clone_shifted(v, shift):
        o, a, b, imm = values[v]
        if o in (V_ZERO, V_CONST):
            return v
        if o in (V_INPUT, V_EXT):
            return mk64(V_TAP, v, shift)
        if o == V_TAP:
            return mk64(V_TAP, a, shift_of(values[v]) + shift)
        if o == V_GATE:
            return mk64(V_GATE, clone_shifted(a, shift), shift_of(values[v]) + shift)
        if o == V_DELAY:
            d = const_delay_of(b)
            if d is None:
                return mk(V_ZERO)
            if is_ti(a):
                return mk64(V_GATE, a, shift + d)
            return mk64(V_GATE, mk64(V_TAP, a, shift + d), shift + d)
        ca = clone_shifted(a, shift)
        cb = clone_shifted(b, shift)
        return mk(o, ca, cb)

    def maybe(ctx, vec, slot):
        if slot < len(vec) and vec[slot] is not None:
            return resolve(ctx, vec[slot])
        return mk(V_ZERO)

    # ctx = (graph, parent ctx, node in the parent) ; identity by id() chain
    def resolve(ctx, e):
        g, parent, pnode, cid = ctx
        frm, _, from_slot, _ = e
        if frm == 0:
            if parent is None:
                n_inputs[0] = max(n_inputs[0], from_slot + 1)
                return mk(V_INPUT, 0, 0, from_slot)
            return maybe(parent, pnode["inbound"], from_slot)
        key = (cid, frm, from_slot)
        if key in edge_memo:
            return edge_memo[key]
        n = g.nodes[frm]
        kind = n["kind"]
        if kind == 1:
            v = mk(V_CONST, 0, 0, from_slot)
        elif kind == 0:
            assert from_slot == 0
            src = maybe(ctx, n["inbound"], 0)
            amt = maybe(ctx, n["inbound"], 1)
            if values[src][0] == V_ZERO:
                v = src
            else:
                sop = values[src][0]
                computed = (V_SUM2 <= sop <= V_MIN) or sop == V_GATE
                v = None
                if sop in (V_INPUT, V_EXT, V_TAP) and values[amt][0] in (V_CONST, V_ZERO):
                    # a stored signal delayed by a constant: read at t - d (a TAP)
                    d = const_delay_of(amt)
                    base = shift_of(values[src]) if sop == V_TAP else 0
                    if d is None:
                        v = mk(V_ZERO)
                    elif d < (1 << 40) and base < (1 << 40):
                        v = mk64(V_TAP, values[src][1] if sop == V_TAP else src, base + d)
                if computed and values[amt][0] in (V_CONST, V_ZERO):
                    d = const_delay_of(amt)
                    if d is not None and d < (1 << 40) and clone_cost(src, CLONE_MAX_OPS) is not None and not is_ti(src):
                        v = mk64(V_GATE, clone_shifted(src, d), d)
                if v is None:
                    v = mk(V_DELAY, src, amt)
        elif kind in KIND_TO_OP:
            assert from_slot == 0
            a = maybe(ctx, n["inbound"], 0)
            b = maybe(ctx, n["inbound"], 1)
            v = mk(KIND_TO_OP[kind], a, b)
        elif kind == 16:
            child = (n["body"], ctx, n, cid + (frm,))
            v = maybe(child, n["body"].outputs, from_slot)
        else:
            xk = (cid, frm)
            if xk not in ext_memo:
                lanes, maxd = top.ext[(kind, n["key"])]
                inputs = [maybe(ctx, n["inbound"], l) for l in range(lanes)] if kind != 32 else []
                inst = len(ext)
                ext.append(dict(kind=kind - 32, key=n["key"], lanes=lanes, inputs=inputs, maxd=maxd))
                ext_memo[xk] = inst
                for l in range(lanes):
                    mk(V_EXT, inst, 0, l)
            inst = ext_memo[xk]
            v = mk(V_EXT, inst, 0, from_slot) if from_slot < ext[inst]["lanes"] else mk(V_ZERO)
        edge_memo[key] = v
        return v

    mk(V_ZERO)
    root = (top, None, None, ())
    outputs = [maybe(root, top.outputs, s) for s in range(n_slots)]
    nv = len(values)
    op = [v[0] for v in values]
    is_leaf = lambda v: op[v] in (V_ZERO, V_CONST, V_INPUT)
    ti = [False] * nv
    for v in range(nv):
        if op[v] in (V_ZERO, V_CONST):
            ti[v] = True
        elif V_SUM2 <= op[v] <= V_MIN:
            ti[v] = ti[values[v][1]] and ti[values[v][2]]
    st = [0] * nv
    for v in range(nv):
        o, a, b, imm = values[v]
        if V_SUM2 <= o <= V_MIN:
            st[v] = max(st[a], st[b])
        elif o == V_DELAY:
            s = st[b]
            if not is_leaf(a) and not ti[a]:
                s = max(s, st[a] if op[a] == V_EXT else st[a] + 1)
            st[v] = s
        elif o in (V_TAP, V_GATE):
            st[v] = st[a]
        elif o == V_EXT:
            x = ext[a]
            if "stage" not in x:
                x["stage"] = max([st[i] if op[i] == V_EXT else st[i] + 1 for i in x["inputs"]] or [0])
            st[v] = x["stage"]
    need = [False] * nv
    for v in range(nv):
        o, a, b, imm = values[v]
        if o == V_EXT:
            need[v] = True
            continue

        def use(u):
            if not is_leaf(u) and st[u] < st[v]:
                need[u] = True
        if o == V_DELAY:
            if not is_leaf(a) and not ti[a]:
                need[a] = True
            else:
                use(a)
            use(b)
        elif V_SUM2 <= o <= V_MIN:
            use(a)
            use(b)
        elif o == V_GATE:
            use(a)
    for x in ext:
        for i in x["inputs"]:
            need[i] = True
    buf = [-1] * nv
    nb = 0
    for v in range(nv):
        if need[v]:
            buf[v] = nb
            nb += 1
    L = [0] * nv

    def raise_(v, l):
        if L[v] == FULL:
            return
        if l == FULL or l > L[v]:
            L[v] = l
    from_zero = False
    for v in range(nv - 1, -1, -1):
        o, a, b, imm = values[v]
        if V_SUM2 <= o <= V_MIN:
            raise_(a, L[v])
            raise_(b, L[v])
        elif o == V_DELAY:
            raise_(b, L[v])
            if op[b] == V_CONST:
                d = const_delay(values[b][3])
            elif op[b] == V_ZERO:
                d = 0
            else:
                d = FULL
            if d is not None:
                raise_(a, sat_add(L[v], d))
        elif o == V_GATE:
            raise_(a, L[v])
        elif o == V_TAP:
            raise_(a, sat_add(L[v], (b << 32) | imm))
        elif o == V_EXT and imm == 0:
            x = ext[a]
            li = 0
            for l in range(x["lanes"]):
                ll = L[v + l]
                li = FULL if (ll == FULL or li == FULL) else max(li, ll)
            own, inl = li, li
            if x["kind"] == 1:
                own = li if li == FULL else max(li, 2)
                inl = sat_add(li, 2)
            if x["kind"] == 2:
                own = li if li == FULL else max(li, x["maxd"])
            for l in range(x["lanes"]):
                L[v + l] = own
            for i in x["inputs"]:
                raise_(i, inl)
            if x["kind"] != 0:
                from_zero = True
    full_history = any(l == FULL for l in L)
    return dict(values=values, outputs=outputs, stage=st, buffer=buf, lookback=L, n_buffers=nb, n_inputs=n_inputs[0],
                n_stages=max(st) + 1 if st else 1, n_ext=len(ext), from_zero=from_zero or full_history,
                full_history=full_history, ext=ext)


def parse_dump(w):
    """Decodes frb_dump_schedule words (layout: libfriendship_b200/csrc/schedule.hpp)."""
    w = [int(x) for x in w]
    assert w[0] == 0x53425246
    nv, no, nb, ns, nx, nin, flags = w[1:8]
    p = 8
    values, stage, buf = [], [], []
    for _ in range(nv):
        values.append(tuple(w[p:p + 4]))
        stage.append(w[p + 4])
        buf.append(w[p + 5] - 1)
        p += 6
    outputs = w[p:p + no]
    p += no
    buffers = []
    for _ in range(nb):
        buffers.append(dict(value=w[p], lookback=w[p + 1] | (w[p + 2] << 32), ext=w[p + 3] - 1, lane=w[p + 4]))
        p += 5
    stages = []
    for _ in range(ns):
        n_ext, n_instr, n_regs = w[p:p + 3]
        p += 3
        ext_ids = w[p:p + n_ext]
        p += n_ext
        instrs = [tuple(w[p + 4 * i:p + 4 * i + 4]) for i in range(n_instr)]
        p += 4 * n_instr
        stages.append(dict(ext=ext_ids, instrs=instrs, n_regs=n_regs))
    assert p == len(w)
    return dict(values=values, stage=stage, buffer=buf, outputs=outputs, buffers=buffers, stages=stages, n_inputs=nin,
                from_zero=bool(flags & 1), full_history=bool(flags & 2), n_ext=nx)
