// flatten.cc — routed effect tree -> topologically ordered schedule (see schedule.hpp).
//
// Semantics being flattened (reference src/render/reference.rs:178-266 and SURVEY.md Appendix A.5):
//   * an edge from the toplevel of a nested effect reads the OUTER node's inbound edge of that slot (:189-193),
//   * an edge from a nested-effect node reads the inner graph's output edge of that slot (:188-190),
//   * a missing edge or out-of-range slot is 0.0f (:164-173),
//   * every node is a pure function of (absolute time, inputs), so a shared sub-graph is evaluated once here
//     (hash-consing) where the reference re-evaluates it per consumer: same values, less work.
// Order: depth-first from output slot 0 upward, inbound slot 0 before slot 1 (Delay: source before frames) —
// the deterministic counterpart of RouteGraph::iter_nodes_dep_first (routegraph.rs:105-126), whose own order
// depends on HashMap iteration.
#include "flatten.hpp"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>
#include <functional>
#include <exception>
#include <map>
#include <set>

#include <pthread.h>

namespace frb {

namespace {

constexpr int kMaxDepthGuard = 500000;            // levels of recursion the walk's own stack is sized for (flatten() below)

struct Ctx {
    const Graph* g;
    int parent;                 // context of the enclosing graph, -1 at the top
    const GraphNode* pnode;     // the nested-effect node in the parent graph
};

struct Flattener {
    const FlattenEnv& env;
    Schedule& s;
    std::vector<Ctx> ctxs;
    std::map<std::pair<int, uint32_t>, int> child_ctx;              // (ctx, node handle) -> ctx
    std::map<std::array<uint32_t, 4>, uint32_t> cons;               // hash-consing of values
    std::map<std::array<uint32_t, 3>, uint32_t> edge_memo;          // (ctx, from, from_slot) -> value
    std::map<std::pair<int, uint32_t>, uint32_t> ext_memo;          // (ctx, node handle) -> ext instance
    int depth = 0;

    Flattener(const FlattenEnv& e, Schedule& sch) : env(e), s(sch) {}

    uint32_t mk(uint8_t op, uint32_t a, uint32_t b, uint32_t imm) {
        std::array<uint32_t, 4> k{op, a, b, imm};
        auto it = cons.find(k);
        if (it != cons.end()) return it->second;
        uint32_t id = (uint32_t)s.values.size();
        s.values.push_back(Value{op, a, b, imm});
        cons.emplace(k, id);
        return id;
    }
    uint32_t zero() { return mk(V_ZERO, 0, 0, 0); }

    // ---- Delay of a cheap, purely elementwise source: re-evaluate the source at the shifted time ----
    // Delay(src, d)(t) = (t >= d) ? src(t - d) : 0.  When src is a small expression over stored signals (external
    // inputs, extension lanes), constants and constant Delays, src(t - d) is the same expression over those signals
    // read at t - d: same operations on the same operands, hence bit-identical, and no ring write + read of src.
    static constexpr uint32_t CLONE_MAX_OPS = 8;
    static uint64_t shift_of(const Value& x) { return ((uint64_t)x.b << 32) | x.imm; }
    uint32_t mk64(uint8_t op, uint32_t a, uint64_t v) { return mk(op, a, (uint32_t)(v >> 32), (uint32_t)(v & 0xffffffffu)); }

    // number of operations a shifted copy of v would add, or UINT32_MAX if v cannot be shifted
    uint32_t clone_cost(uint32_t v, uint32_t budget) {
        const Value x = s.values[v];
        switch (x.op) {
            case V_ZERO: case V_CONST: return 0;
            case V_INPUT: case V_EXT: case V_TAP: return 0;
            case V_GATE: return clone_cost(x.a, budget);
            case V_SUM2: case V_MUL: case V_DIV: case V_MOD: case V_MIN: {
                if (budget == 0) return UINT32_MAX;
                uint32_t ca = clone_cost(x.a, budget - 1);
                if (ca == UINT32_MAX) return ca;
                uint32_t cb = clone_cost(x.b, budget - 1 - std::min(ca, budget - 1));
                if (cb == UINT32_MAX || 1 + ca + cb > budget) return UINT32_MAX;
                return 1 + ca + cb;
            }
            case V_DELAY: {
                // a Delay that stayed a Delay has a leaf / constant-expression source or a signal-driven amount
                const Value amt = s.values[x.b];
                if (amt.op != V_CONST && amt.op != V_ZERO) return UINT32_MAX;
                const uint8_t sop = s.values[x.a].op;
                if (sop != V_INPUT && sop != V_EXT && !is_time_invariant(x.a)) return UINT32_MAX;
                return 0;
            }
            default: return UINT32_MAX;
        }
    }
    bool const_delay_of(const Value& amt, uint64_t* d) const;   // defined below
    uint32_t clone_shifted(uint32_t v, uint64_t shift) {
        const Value x = s.values[v];
        switch (x.op) {
            case V_ZERO: case V_CONST: return v;
            case V_INPUT: case V_EXT: return mk64(V_TAP, v, shift);
            case V_TAP: return mk64(V_TAP, x.a, shift_of(x) + shift);
            case V_GATE: return mk64(V_GATE, clone_shifted(x.a, shift), shift_of(x) + shift);
            case V_DELAY: {
                uint64_t d = 0;
                if (!const_delay_of(s.values[x.b], &d)) return zero();          // never reads its source
                if (is_time_invariant(x.a)) return mk64(V_GATE, x.a, shift + d);  // a constant expression is the same at any time
                return mk64(V_GATE, mk64(V_TAP, x.a, shift + d), shift + d);
            }
            default: {
                uint32_t a = clone_shifted(x.a, shift), b = clone_shifted(x.b, shift);
                return mk(x.op, a, b, 0);
            }
        }
    }

    bool is_time_invariant(uint32_t v) {
        const Value x = s.values[v];
        if (x.op == V_ZERO || x.op == V_CONST) return true;
        if (x.op >= V_SUM2 && x.op <= V_MIN) return is_time_invariant(x.a) && is_time_invariant(x.b);
        return false;
    }

    uint32_t resolve_maybe(int ctx, const std::vector<std::optional<frb_edge>>& vec, size_t slot) {
        if (slot < vec.size() && vec[slot].has_value()) return resolve(ctx, *vec[slot]);
        return zero();
    }

    uint32_t resolve(int ctx, const frb_edge& e) {
        if (++depth > kMaxDepthGuard) throw Error{FRB_E_UNSUPPORTED, "graph too deep to flatten (longest path over 500,000 nodes)"};
        struct Guard { int& d; ~Guard() { --d; } } guard{depth};
        const Ctx c = ctxs[ctx];
        if (e.from == 0) {
            // reading an input of this graph level
            if (c.parent < 0) {
                if (e.from_slot >= env.input_slot_cap) return zero();     // no such slot exists (FlattenEnv::input_slot_cap)
                s.n_input_slots = std::max<uint64_t>(s.n_input_slots, (uint64_t)e.from_slot + 1);
                return mk(V_INPUT, 0, 0, e.from_slot);
            }
            return resolve_maybe(c.parent, c.pnode->inbound, e.from_slot);
        }
        std::array<uint32_t, 3> mk_key{(uint32_t)ctx, e.from, e.from_slot};
        auto mit = edge_memo.find(mk_key);
        if (mit != edge_memo.end()) return mit->second;

        auto nit = c.g->nodes.find(e.from);
        if (nit == c.g->nodes.end())
            throw Error{FRB_E_BAD_HANDLE, "edge reads node " + std::to_string(e.from) + " which does not exist"};
        const GraphNode& n = nit->second;
        uint32_t v;
        auto need_slot0 = [&](const char* what) {
            if (e.from_slot != 0)
                throw Error{FRB_E_BAD_SLOT, std::string(what) + ": from_slot must be 0 (reference.rs:199,223,230,237,244,251)"};
        };
        switch (n.kind) {
            case FRB_KIND_F32CONSTANT: v = mk(V_CONST, 0, 0, e.from_slot); break;
            case FRB_KIND_DELAY: {
                need_slot0("Delay");
                uint32_t src = resolve_maybe(ctx, n.inbound, 0);
                uint32_t amt = resolve_maybe(ctx, n.inbound, 1);
                // Delay of the constant-zero signal is zero at every t (exactly +0.0f either way).
                if (s.values[src].op == V_ZERO) { v = src; break; }
                {
                    // Delay of a stored signal (an external input, an extension lane, or one of those already shifted)
                    // by a constant amount is that signal read at t - d, 0 before t = d: a TAP — 128-bit reads wherever
                    // t - d is a multiple of 4 — instead of the general Delay with its per-sample clamps.
                    const Value sv = s.values[src];
                    const Value amtv = s.values[amt];
                    uint64_t d = 0;
                    if ((sv.op == V_INPUT || sv.op == V_EXT || sv.op == V_TAP) && (amtv.op == V_CONST || amtv.op == V_ZERO)) {
                        if (!const_delay_of(amtv, &d)) { v = zero(); break; }              // never reads its source
                        const uint64_t base = sv.op == V_TAP ? shift_of(sv) : 0;
                        if (d < (1ull << 40) && base < (1ull << 40)) {
                            v = mk64(V_TAP, sv.op == V_TAP ? sv.a : src, base + d);
                            break;
                        }
                    }
                }
                {
                    const uint8_t sop = s.values[src].op;
                    const bool computed = (sop >= V_SUM2 && sop <= V_MIN) || sop == V_GATE;
                    uint64_t d = 0;
                    const Value amtv = s.values[amt];
                    if (computed && (amtv.op == V_CONST || amtv.op == V_ZERO) && const_delay_of(amtv, &d) && d < (1ull << 40) &&
                        clone_cost(src, CLONE_MAX_OPS) != UINT32_MAX && !is_time_invariant(src)) {
                        v = mk64(V_GATE, clone_shifted(src, d), d);
                        break;
                    }
                }
                v = mk(V_DELAY, src, amt, 0);
                break;
            }
            case FRB_KIND_SUM2: case FRB_KIND_MULTIPLY: case FRB_KIND_DIVIDE: case FRB_KIND_MODULO: case FRB_KIND_MINIMUM: {
                need_slot0("binary primitive");
                uint32_t a = resolve_maybe(ctx, n.inbound, 0);
                uint32_t b = resolve_maybe(ctx, n.inbound, 1);
                uint8_t op = n.kind == FRB_KIND_SUM2 ? V_SUM2 : n.kind == FRB_KIND_MULTIPLY ? V_MUL
                           : n.kind == FRB_KIND_DIVIDE ? V_DIV : n.kind == FRB_KIND_MODULO ? V_MOD : V_MIN;
                if (env.shard_world > 1) {
                    // the lanes of other ranks are zero signals here: a sum keeps its other operand, a product or a
                    // quotient of a zero lane is that rank's (zero) share of a linear output
                    const bool za = s.values[a].op == V_ZERO, zb = s.values[b].op == V_ZERO;
                    if (op == V_SUM2 && za) { v = b; break; }
                    if (op == V_SUM2 && zb) { v = a; break; }
                    if (op == V_MUL && (za || zb)) { v = zero(); break; }
                    if (op == V_DIV && za) { v = zero(); break; }
                }
                v = mk(op, a, b, 0);
                break;
            }
            case FRB_KIND_EFFECT: {
                auto key = std::make_pair(ctx, e.from);
                auto cit = child_ctx.find(key);
                int child;
                if (cit == child_ctx.end()) {
                    child = (int)ctxs.size();
                    ctxs.push_back(Ctx{n.body.get(), ctx, &n});
                    child_ctx.emplace(key, child);
                } else {
                    child = cit->second;
                }
                v = resolve_maybe(child, n.body->output_edges, e.from_slot);
                break;
            }
            case FRB_KIND_OSCBANK: case FRB_KIND_DIRECTFORM: case FRB_KIND_FBDELAY: {
                auto key = std::make_pair(ctx, e.from);
                auto xit = ext_memo.find(key);
                uint32_t inst;
                if (xit == ext_memo.end()) {
                    int64_t lanes = env.ext_lanes(n.kind, n.key);
                    if (lanes < 0) throw Error{FRB_E_BAD_HANDLE, "extension node with undefined key"};
                    ExtInstance x;
                    x.kind = n.kind == FRB_KIND_OSCBANK ? EXT_OSCBANK : n.kind == FRB_KIND_DIRECTFORM ? EXT_DIRECTFORM : EXT_FBDELAY;
                    x.key = n.key;
                    x.n_lanes = (uint32_t)lanes;
                    if (x.kind != EXT_OSCBANK)
                        for (uint32_t l = 0; l < x.n_lanes; l++) x.inputs.push_back(resolve_maybe(ctx, n.inbound, l));
                    inst = (uint32_t)s.ext.size();
                    s.ext.push_back(std::move(x));
                    ext_memo.emplace(key, inst);
                    // all lanes get consecutive value ids right after the instance's inputs
                    for (uint32_t l = 0; l < s.ext[inst].n_lanes; l++) mk(V_EXT, inst, 0, l);
                } else {
                    inst = xit->second;
                }
                if (env.shard_world > 1 && n.kind == FRB_KIND_OSCBANK) {
                    // this rank's bank holds the voices it owns, compacted: global voice = shard_rank + lane * shard_world
                    // (osc_create); everybody else's voice is the zero signal here
                    const uint32_t lane = e.from_slot / env.shard_world;
                    v = (e.from_slot % env.shard_world == env.shard_rank && lane < s.ext[inst].n_lanes) ? mk(V_EXT, inst, 0, lane) : zero();
                    break;
                }
                v = (e.from_slot < s.ext[inst].n_lanes) ? mk(V_EXT, inst, 0, e.from_slot) : zero();
                if (env.shard_world > 1 && s.values[v].op == V_EXT && s.values[s.ext[inst].inputs[e.from_slot]].op == V_ZERO)
                    v = zero();                              // a zero-state linear filter of silence
                break;
            }
            default: throw Error{FRB_E_INVALID, "unknown node kind " + std::to_string(n.kind)};
        }
        edge_memo.emplace(mk_key, v);
        return v;
    }
};

bool const_delay(uint32_t bits, bool sparkle, uint64_t* out);

bool Flattener::const_delay_of(const Value& amt, uint64_t* d) const {
    if (amt.op == V_ZERO) { *d = 0; return true; }
    return const_delay(amt.imm, env.sparkle_delay, d);
}

// Delay amount of a constant `frames` signal, with the clamps of reference.rs:200-212.
// Returns false when the Delay can never read its source (frames >= 2^64).
bool const_delay(uint32_t bits, bool sparkle, uint64_t* out) {
    float d;
    std::memcpy(&d, &bits, 4);
    if (d >= 18446744073709551616.0f) return false;
    if (!(d >= 0.0f)) {                             // negative or NaN
        if (sparkle) return false;                  // sparkle.rs:525-542: the Delay outputs 0.0 and never reads its source
        *out = 0;                                   // reference.rs:206-207: delay 0
        return true;
    }
    *out = (uint64_t)d;
    return true;
}

uint64_t sat_add(uint64_t a, uint64_t b) {
    if (a == LOOKBACK_FULL || b == LOOKBACK_FULL) return LOOKBACK_FULL;
    uint64_t r = a + b;
    if (r < a || r > (1ull << 40)) return LOOKBACK_FULL;
    return r;
}

}  // namespace

namespace {
Schedule flatten_here(const Graph& top, uint32_t n_slots, const FlattenEnv& env);
}

// The depth-first walk recurses once per node of the longest path (about 600 B of stack per level): a Sum2 chain of
// 15,000 nodes overflows the 8 MB of a default thread — the reference's per-sample recursion (reference.rs:178-266)
// has the same limit.  The walk therefore runs on a thread of its own whose stack (reserved, committed only as far as
// it is touched) holds kMaxDepthGuard (500,000) levels, and deeper graphs are refused with FRB_E_UNSUPPORTED instead of crashing
// the host process.
constexpr size_t kFlattenStack = 512ull << 20;

Schedule flatten(const Graph& top, uint32_t n_slots, const FlattenEnv& env) {
    struct Job { const Graph* top; uint32_t n_slots; const FlattenEnv* env; Schedule out; std::exception_ptr err; } job{&top, n_slots, &env, {}, nullptr};
    auto body = [](void* p) -> void* {
        Job& j = *static_cast<Job*>(p);
        try { j.out = flatten_here(*j.top, j.n_slots, *j.env); } catch (...) { j.err = std::current_exception(); }
        return nullptr;
    };
    pthread_attr_t attr;
    pthread_t th;
    bool started = false;
    if (pthread_attr_init(&attr) == 0) {
        if (pthread_attr_setstacksize(&attr, kFlattenStack) == 0) started = pthread_create(&th, &attr, body, &job) == 0;
        pthread_attr_destroy(&attr);
    }
    if (!started) throw Error{FRB_E_UNSUPPORTED, "cannot reserve the stack for the graph walk"};
    pthread_join(th, nullptr);
    if (job.err) std::rethrow_exception(job.err);
    return std::move(job.out);
}

namespace {
Schedule flatten_here(const Graph& top, uint32_t n_slots, const FlattenEnv& env) {
    Schedule s;
    Flattener f(env, s);
    f.ctxs.push_back(Ctx{&top, -1, nullptr});
    f.zero();   // value 0 is always the zero signal
    for (uint32_t slot = 0; slot < n_slots; slot++) s.outputs.push_back(f.resolve_maybe(0, top.output_edges, slot));

    const size_t nv = s.values.size();
    auto& V = s.values;
    auto is_leaf = [&](uint32_t v) { return V[v].op == V_ZERO || V[v].op == V_CONST || V[v].op == V_INPUT; };

    // time-invariant values: constant expressions (no inputs, delays or extension outputs underneath)
    std::vector<uint8_t> ti(nv, 0);
    for (size_t v = 0; v < nv; v++) {
        switch (V[v].op) {
            case V_ZERO: case V_CONST: ti[v] = 1; break;
            case V_SUM2: case V_MUL: case V_DIV: case V_MOD: case V_MIN: ti[v] = ti[V[v].a] && ti[V[v].b]; break;
            default: ti[v] = 0;
        }
    }

    // ---- stages ----
    s.value_stage.assign(nv, 0);
    auto& st = s.value_stage;
    std::vector<uint8_t> ext_staged(s.ext.size(), 0);
    for (size_t v = 0; v < nv; v++) {
        const Value& x = V[v];
        switch (x.op) {
            case V_ZERO: case V_CONST: case V_INPUT: st[v] = 0; break;
            case V_SUM2: case V_MUL: case V_DIV: case V_MOD: case V_MIN: st[v] = std::max(st[x.a], st[x.b]); break;
            case V_DELAY: {
                uint32_t sg = st[x.b];
                if (!is_leaf(x.a) && !ti[x.a]) sg = std::max(sg, V[x.a].op == V_EXT ? st[x.a] : st[x.a] + 1);
                st[v] = sg;
                break;
            }
            case V_TAP: case V_GATE: st[v] = st[x.a]; break;
            case V_EXT: {
                ExtInstance& inst = s.ext[x.a];
                if (!ext_staged[x.a]) {
                    uint32_t sg = 0;
                    for (uint32_t in : inst.inputs) sg = std::max(sg, V[in].op == V_EXT ? st[in] : st[in] + 1);
                    inst.stage = sg;
                    ext_staged[x.a] = 1;
                }
                st[v] = inst.stage;
                break;
            }
        }
    }
    uint32_t n_stages = 1;
    for (size_t v = 0; v < nv; v++) n_stages = std::max(n_stages, st[v] + 1);
    s.stages.resize(n_stages);
    for (uint32_t i = 0; i < s.ext.size(); i++) s.stages[s.ext[i].stage].ext.push_back(i);

    // ---- which values are materialised in HBM ----
    std::vector<uint8_t> need_buf(nv, 0);
    for (size_t v = 0; v < nv; v++) {
        const Value& x = V[v];
        if (x.op == V_EXT) { need_buf[v] = 1; continue; }
        auto use = [&](uint32_t u) {
            if (!is_leaf(u) && st[u] < st[v]) need_buf[u] = 1;          // consumed by a later stage
        };
        if (x.op == V_DELAY) {
            if (!is_leaf(x.a) && !ti[x.a]) need_buf[x.a] = 1;           // Delay source becomes an indexed read
            else use(x.a);                                              // constant expression from an earlier stage
            use(x.b);
        } else if (x.op >= V_SUM2 && x.op <= V_MIN) {
            use(x.a); use(x.b);
        } else if (x.op == V_GATE) {
            use(x.a);
        }
    }
    for (auto& inst : s.ext)
        for (uint32_t in : inst.inputs) need_buf[in] = 1;               // filters read planes (leaves included)

    s.value_buffer.assign(nv, -1);
    for (size_t v = 0; v < nv; v++) {
        if (!need_buf[v]) continue;
        s.value_buffer[v] = (int32_t)s.buffers.size();
        BufferInfo b;
        b.value = (uint32_t)v;
        b.lookback = 0;
        if (V[v].op == V_EXT) { b.ext = V[v].a; b.lane = V[v].imm; }
        s.buffers.push_back(b);
    }
    for (uint32_t i = 0; i < s.ext.size(); i++) {
        ExtInstance& inst = s.ext[i];
        std::array<uint32_t, 4> k{V_EXT, i, 0, 0};
        uint32_t v0 = f.cons.at(k);
        inst.first_out_buf = (uint32_t)s.value_buffer[v0];
        for (uint32_t in : inst.inputs) inst.in_bufs.push_back((uint32_t)s.value_buffer[in]);
    }

    // ---- lookback: how far before the block start each value must stay addressable ----
    std::vector<uint64_t> L(nv, 0);
    auto raise = [&](uint32_t v, uint64_t l) {
        if (L[v] == LOOKBACK_FULL) return;
        if (l == LOOKBACK_FULL || l > L[v]) L[v] = l;
    };
    std::vector<uint8_t> ext_done(s.ext.size(), 0);
    for (size_t vi = nv; vi-- > 0;) {
        const Value& x = V[vi];
        switch (x.op) {
            case V_SUM2: case V_MUL: case V_DIV: case V_MOD: case V_MIN: raise(x.a, L[vi]); raise(x.b, L[vi]); break;
            case V_DELAY: {
                raise(x.b, L[vi]);
                uint64_t d = 0;
                bool reads = true;
                if (V[x.b].op == V_CONST) reads = const_delay(V[x.b].imm, env.sparkle_delay, &d);
                else if (V[x.b].op == V_ZERO) d = 0;
                else d = LOOKBACK_FULL;                                  // signal-driven delay (reference.rs:200)
                if (reads) raise(x.a, sat_add(L[vi], d));
                break;
            }
            case V_GATE: raise(x.a, L[vi]); break;
            case V_TAP: raise(x.a, sat_add(L[vi], Flattener::shift_of(x))); break;
            case V_EXT: {
                if (x.imm != 0) break;                                   // lanes are consecutive; lane 0 has the lowest id
                ExtInstance& inst = s.ext[x.a];
                uint64_t li = 0;
                for (uint32_t l = 0; l < inst.n_lanes; l++) {
                    uint64_t ll = L[vi + l];
                    if (ll == LOOKBACK_FULL) li = LOOKBACK_FULL;
                    else if (li != LOOKBACK_FULL) li = std::max(li, ll);
                }
                uint64_t own = li, in_l = li;
                if (inst.kind == EXT_DIRECTFORM) {                       // needs x[n-1], x[n-2], y[n-1], y[n-2]
                    own = (li == LOOKBACK_FULL) ? li : std::max<uint64_t>(li, 2);
                    in_l = sat_add(li, 2);
                }
                if (inst.kind == EXT_FBDELAY) {
                    uint64_t dmax = env.ext_max_delay ? env.ext_max_delay(inst.key) : 0;
                    own = (li == LOOKBACK_FULL) ? li : std::max<uint64_t>(li, dmax);
                }
                for (uint32_t l = 0; l < inst.n_lanes; l++) L[vi + l] = own;
                for (uint32_t in : inst.inputs) raise(in, in_l);
                if (inst.kind != EXT_OSCBANK) s.from_zero = true;         // recurrences are defined from t = 0
                break;
            }
            default: break;
        }
    }
    for (size_t v = 0; v < nv; v++) {
        if (L[v] == LOOKBACK_FULL) { s.full_history = true; s.from_zero = true; }
        else s.max_lookback = std::max(s.max_lookback, L[v]);
        if (s.value_buffer[v] >= 0) s.buffers[s.value_buffer[v]].lookback = L[v];
    }

    // ---- one register program per stage ----
    std::vector<std::vector<uint32_t>> out_slots_of(nv);
    for (uint32_t k = 0; k < s.outputs.size(); k++) out_slots_of[s.outputs[k]].push_back(k);

    // Sum2 chains over consecutive lanes of an extension node — the "voices summed to one slot" mix,
    // ((x0 + x1) + x2) + ... — become ONE fold instruction with the same left-to-right order (bit-exact), so a
    // 4,096-voice mix is a streaming loop over planes instead of 8,192 interpreted instructions per thread.
    std::vector<uint32_t> n_uses(nv, 0);
    for (size_t v = 0; v < nv; v++) {
        if (V[v].op == V_DELAY || (V[v].op >= V_SUM2 && V[v].op <= V_MIN)) { n_uses[V[v].a]++; n_uses[V[v].b]++; }
        if (V[v].op == V_TAP || V[v].op == V_GATE) n_uses[V[v].a]++;
    }
    for (uint32_t o : s.outputs) n_uses[o]++;
    for (auto& inst : s.ext) for (uint32_t in : inst.inputs) n_uses[in]++;
    std::vector<uint32_t> fold_first(nv, 0), fold_len(nv, 0);
    std::vector<uint8_t> absorbed(nv, 0);
    for (size_t v = 0; v < nv; v++) {
        const Value& x = V[v];
        if (x.op != V_SUM2 || V[x.b].op != V_EXT || st[x.b] > st[v]) continue;
        const uint32_t bb = (uint32_t)s.value_buffer[x.b];
        if (V[x.a].op == V_EXT && V[x.a].a == V[x.b].a && st[x.a] <= st[v] && (uint32_t)s.value_buffer[x.a] + 1 == bb) {
            fold_first[v] = (uint32_t)s.value_buffer[x.a];
            fold_len[v] = 2;
        } else if (fold_len[x.a] > 0 && fold_first[x.a] + fold_len[x.a] == bb && s.buffers[fold_first[x.a]].ext == V[x.b].a && n_uses[x.a] == 1 &&
                   s.value_buffer[x.a] < 0 && st[x.a] == st[v]) {
            fold_first[v] = fold_first[x.a];
            fold_len[v] = fold_len[x.a] + 1;
            absorbed[x.a] = 1;
        }
    }

    for (uint32_t sg = 0; sg < n_stages; sg++) {
        struct VI { uint8_t op; uint32_t flags; uint32_t dst, a, b, aux; };   // with virtual registers
        std::vector<VI> code;
        std::vector<int64_t> vreg_of(nv, -1);
        uint32_t n_vreg = 0;
        auto new_vreg = [&]() { return n_vreg++; };
        const uint32_t NOREG = 0xFFFFFFFFu;   // not a virtual register (a stage may have more than 65,535); packs as 0xFFFF in Instr::make

        // operand: immediate for constants, register otherwise (loading leaves / planes on first use)
        auto operand = [&](uint32_t u, uint32_t imm_flag, uint32_t* flags) -> uint32_t {
            if (V[u].op == V_ZERO) { *flags |= imm_flag; return 0u; }
            if (V[u].op == V_CONST) { *flags |= imm_flag; return V[u].imm; }
            if (vreg_of[u] >= 0) return (uint32_t)vreg_of[u];
            uint32_t r = new_vreg();
            if (V[u].op == V_INPUT) code.push_back(VI{I_LDIN, 0, r, 0, 0, V[u].imm});
            else {
                if (s.value_buffer[u] < 0) throw Error{FRB_E_UNSUPPORTED, "internal: value needed across stages has no buffer"};
                code.push_back(VI{I_LDBUF, 0, r, 0, 0, (uint32_t)s.value_buffer[u]});
            }
            vreg_of[u] = r;
            return r;
        };
        auto emit_sinks = [&](uint32_t v, uint32_t flags_a, uint32_t a) {
            if (s.value_buffer[v] >= 0 && V[v].op != V_EXT)
                code.push_back(VI{I_STBUF, flags_a, NOREG, a, 0, (uint32_t)s.value_buffer[v]});
            for (uint32_t k : out_slots_of[v]) code.push_back(VI{I_STOUT, flags_a, NOREG, a, 0, k});
        };

        for (size_t v = 0; v < nv; v++) {
            if (st[v] != sg || absorbed[v]) continue;
            const Value& x = V[v];
            if (fold_len[v] >= 3) {
                // the fold kernel materialises the chain's value; the program only routes it to its sinks
                if (s.value_buffer[v] < 0) {
                    s.value_buffer[v] = (int32_t)s.buffers.size();
                    BufferInfo b;
                    b.value = (uint32_t)v;
                    b.lookback = L[v];
                    s.buffers.push_back(b);
                }
                s.stages[sg].folds.push_back(FoldJob{fold_first[v], fold_len[v], (uint32_t)s.value_buffer[v]});
                if (!out_slots_of[v].empty()) {
                    uint32_t dst = new_vreg();
                    code.push_back(VI{I_LDBUF, 0, dst, 0, 0, (uint32_t)s.value_buffer[v]});
                    vreg_of[v] = dst;
                    for (uint32_t k : out_slots_of[v]) code.push_back(VI{I_STOUT, 0, NOREG, dst, 0, k});
                }
                continue;
            }
            if (is_leaf((uint32_t)v) || x.op == V_EXT) {
                // leaves and extension outputs only appear in a program when they feed a sink directly
                bool sink = !out_slots_of[v].empty() || (s.value_buffer[v] >= 0 && x.op != V_EXT);
                if (!sink) continue;
                uint32_t fl = 0;
                uint32_t a = operand((uint32_t)v, IF_A_IMM, &fl);
                emit_sinks((uint32_t)v, fl, a);
                continue;
            }
            uint32_t fl = 0, dst;
            if (x.op == V_TAP) {
                dst = new_vreg();
                if (V[x.a].op == V_INPUT) code.push_back(VI{I_TAP_IN, 0, dst, x.imm, x.b, V[x.a].imm});
                else code.push_back(VI{I_TAP_BUF, 0, dst, x.imm, x.b, (uint32_t)s.value_buffer[x.a]});
            } else if (x.op == V_GATE) {
                uint32_t a = operand(x.a, IF_A_IMM, &fl);
                dst = new_vreg();
                code.push_back(VI{I_GATE, fl, dst, a, x.imm, x.b});
            } else if (x.op == V_DELAY) {
                uint32_t amt = operand(x.b, IF_A_IMM, &fl);
                const Value& src = V[x.a];
                if (src.op == V_INPUT) {
                    dst = new_vreg();
                    code.push_back(VI{I_DLY_IN, fl, dst, amt, 0, src.imm});
                } else if (ti[x.a]) {
                    uint32_t b = operand(x.a, IF_B_IMM, &fl);
                    dst = new_vreg();
                    code.push_back(VI{I_DLY_TI, fl, dst, amt, b, 0});
                } else {
                    dst = new_vreg();
                    code.push_back(VI{I_DLY_BUF, fl, dst, amt, 0, (uint32_t)s.value_buffer[x.a]});
                }
            } else {
                uint32_t a = operand(x.a, IF_A_IMM, &fl);
                uint32_t b = operand(x.b, IF_B_IMM, &fl);
                dst = new_vreg();
                uint8_t op = x.op == V_SUM2 ? I_ADD : x.op == V_MUL ? I_MUL : x.op == V_DIV ? I_DIV : x.op == V_MOD ? I_MOD : I_MIN;
                code.push_back(VI{op, fl, dst, a, b, 0});
            }
            vreg_of[v] = dst;
            emit_sinks((uint32_t)v, 0, dst);
        }

        // ---- strands: independent sub-programs (connected components over registers) run as separate thread
        // blocks (grid.y), so a stage that drives 64 unrelated output slots exposes 64x the memory-level
        // parallelism of one long straight-line program; each strand gets its own register assignment.
        auto uses_a = [](const VI& c) { return !(c.flags & IF_A_IMM) && c.op != I_LDIN && c.op != I_LDBUF && c.op != I_TAP_IN && c.op != I_TAP_BUF; };
        auto uses_b = [](const VI& c) {
            return !(c.flags & IF_B_IMM) && (c.op == I_ADD || c.op == I_MUL || c.op == I_DIV || c.op == I_MOD || c.op == I_MIN || c.op == I_DLY_TI);
        };
        std::vector<uint32_t> parent(n_vreg);
        for (uint32_t i = 0; i < n_vreg; i++) parent[i] = i;
        std::function<uint32_t(uint32_t)> find = [&](uint32_t x) { while (parent[x] != x) x = parent[x] = parent[parent[x]]; return x; };
        auto unite = [&](uint32_t a, uint32_t b) { a = find(a); b = find(b); if (a != b) parent[std::max(a, b)] = std::min(a, b); };
        for (const VI& c : code) {
            uint32_t anchor = c.dst != NOREG ? c.dst : (uses_a(c) ? c.a : NOREG);
            if (anchor == NOREG) continue;
            if (uses_a(c)) unite(anchor, c.a);
            if (uses_b(c)) unite(anchor, c.b);
        }
        constexpr uint32_t MAX_STRANDS = 4096;
        std::map<uint32_t, uint32_t> strand_of_root;          // component root -> strand (in order of first appearance)
        std::vector<std::vector<VI>> strands;
        std::vector<VI> immediates_only;                      // sinks of immediates (e.g. an unconnected output slot)
        for (const VI& c : code) {
            uint32_t anchor = c.dst != NOREG ? c.dst : (uses_a(c) ? c.a : NOREG);
            if (anchor == NOREG) { immediates_only.push_back(c); continue; }
            uint32_t root = find(anchor);
            auto it = strand_of_root.find(root);
            if (it == strand_of_root.end()) {
                uint32_t id = (uint32_t)strand_of_root.size() % MAX_STRANDS;
                it = strand_of_root.emplace(root, id).first;
                if (id >= strands.size()) strands.emplace_back();
            }
            strands[it->second].push_back(c);
        }
        if (!immediates_only.empty()) {
            if (strands.empty()) strands.emplace_back();
            strands[0].insert(strands[0].end(), immediates_only.begin(), immediates_only.end());
        }

        Stage& stage = s.stages[sg];
        uint32_t max_phys = 0;
        for (auto& sc : strands) {
            stage.strand_offsets.push_back((uint32_t)stage.program.size());
            // linear-scan register assignment over the strand's straight-line program
            std::map<uint32_t, int64_t> last_use;
            for (size_t i = 0; i < sc.size(); i++) {
                if (uses_a(sc[i])) last_use[sc[i].a] = (int64_t)i;
                if (uses_b(sc[i])) last_use[sc[i].b] = (int64_t)i;
            }
            // lowest free register first: the allocator's state after a group of instructions depends only on which
            // values are live, so a repeated group (a term of a Sum2 chain, a voice of a mix) gets the same registers
            // every time — which is what lets the stage JIT fold the repetition into a loop (jit.cc)
            std::map<uint32_t, uint32_t> phys;
            std::set<uint32_t> free_list;
            uint32_t n_phys = 0;
            for (size_t i = 0; i < sc.size(); i++) {
                VI c = sc[i];
                const bool a_reg = uses_a(c), b_reg = uses_b(c);
                const uint32_t va = c.a, vb = c.b;
                if (a_reg) c.a = phys.at(va);
                if (b_reg) c.b = phys.at(vb);
                if (a_reg && last_use[va] == (int64_t)i) free_list.insert(phys.at(va));
                if (b_reg && last_use[vb] == (int64_t)i && !(a_reg && vb == va)) free_list.insert(phys.at(vb));
                if (c.dst != NOREG) {
                    const uint32_t vd = c.dst;
                    uint32_t pr;
                    if (!free_list.empty()) { pr = *free_list.begin(); free_list.erase(free_list.begin()); }
                    else pr = n_phys++;
                    phys[vd] = pr;
                    c.dst = pr;
                    if (!last_use.count(vd)) free_list.insert(pr);   // dead value
                }
                stage.program.push_back(Instr::make(c.op, c.flags, c.dst, c.a, c.b, c.aux));
            }
            stage.program.push_back(Instr::make(I_END, 0, 0, 0, 0, 0));
            stage.program.push_back(Instr::make(I_END, 0, 0, 0, 0, 0));   // pad: the kernel prefetches one instruction ahead
            max_phys = std::max(max_phys, n_phys);
        }
        if (strands.empty()) {
            stage.strand_offsets.push_back(0);
            stage.program.push_back(Instr::make(I_END, 0, 0, 0, 0, 0));
            stage.program.push_back(Instr::make(I_END, 0, 0, 0, 0, 0));
        }
        stage.strand_offsets.push_back((uint32_t)stage.program.size());
        stage.n_regs = max_phys;
        if (max_phys > env.max_regs)
            throw Error{FRB_E_UNSUPPORTED, "stage needs " + std::to_string(max_phys) + " live registers; limit " + std::to_string(env.max_regs)};
    }
    return s;
}
}  // namespace

LaneUse lane_use_of_outputs(const Schedule& s) {
    enum : uint8_t { ZERO = 0, IND = 1, LIN = 2, BAD = 3 };      // independent of the lanes / linear in them / neither
    const auto& V = s.values;
    std::vector<uint8_t> c(V.size(), BAD);
    for (size_t v = 0; v < V.size(); v++) {
        const Value& x = V[v];
        switch (x.op) {
            case V_ZERO: c[v] = ZERO; break;
            case V_CONST: case V_INPUT: c[v] = IND; break;
            case V_EXT: {
                const ExtInstance& inst = s.ext[x.a];
                c[v] = inst.kind == EXT_OSCBANK ? LIN : c[inst.inputs[x.imm]];   // DirectForm / FbDelay: linear, zero state
                break;
            }
            case V_SUM2: {
                const uint8_t a = c[x.a], b = c[x.b];
                c[v] = a == ZERO ? b : b == ZERO ? a : (a == b && a != BAD) ? a : BAD;
                break;
            }
            case V_MUL: {
                const uint8_t a = c[x.a], b = c[x.b];
                c[v] = (a == ZERO || b == ZERO) ? (a == BAD || b == BAD ? BAD : ZERO) : (a == IND && b == IND) ? IND
                     : ((a == LIN && b == IND) || (a == IND && b == LIN)) ? LIN : BAD;
                break;
            }
            case V_DIV: {
                const uint8_t a = c[x.a], b = c[x.b];
                c[v] = b != IND ? BAD : a;                   // x / (lane-independent): the class of x
                break;
            }
            case V_MOD: case V_MIN: c[v] = (c[x.a] <= IND && c[x.b] <= IND) ? IND : BAD; break;
            case V_DELAY: c[v] = c[x.b] <= IND ? c[x.a] : BAD; break;          // the amount must not depend on a lane
            case V_TAP: case V_GATE: c[v] = c[x.a]; break;
            default: c[v] = BAD;
        }
    }
    bool lin = false, ind = false;
    for (uint32_t o : s.outputs) {
        if (c[o] == BAD) return LANES_OTHER;
        lin |= c[o] == LIN;
        ind |= c[o] == IND;
    }
    if (!lin) return LANES_UNUSED;
    return ind ? LANES_OTHER : LANES_LINEAR;
}

std::vector<uint32_t> Schedule::dump() const {
    std::vector<uint32_t> w;
    w.push_back(0x53425246u);   // 'FRBS'
    w.push_back((uint32_t)values.size());
    w.push_back((uint32_t)outputs.size());
    w.push_back((uint32_t)buffers.size());
    w.push_back((uint32_t)stages.size());
    w.push_back((uint32_t)ext.size());
    w.push_back((uint32_t)std::min<uint64_t>(n_input_slots, 0xFFFFFFFFull));
    w.push_back((from_zero ? 1u : 0u) | (full_history ? 2u : 0u));
    for (size_t v = 0; v < values.size(); v++) {
        w.push_back(values[v].op); w.push_back(values[v].a); w.push_back(values[v].b); w.push_back(values[v].imm);
        w.push_back(value_stage[v]); w.push_back((uint32_t)(value_buffer[v] + 1));
    }
    for (uint32_t o : outputs) w.push_back(o);
    for (auto& b : buffers) {
        w.push_back(b.value); w.push_back((uint32_t)(b.lookback & 0xFFFFFFFFu)); w.push_back((uint32_t)(b.lookback >> 32));
        w.push_back(b.ext + 1); w.push_back(b.lane);
    }
    for (auto& sg : stages) {
        w.push_back((uint32_t)sg.ext.size()); w.push_back((uint32_t)sg.program.size()); w.push_back(sg.n_regs);
        for (uint32_t e : sg.ext) w.push_back(e);
        for (auto& i : sg.program) { w.push_back(i.w0); w.push_back(i.a); w.push_back(i.b); w.push_back(i.aux); }
    }
    return w;
}

}  // namespace frb
