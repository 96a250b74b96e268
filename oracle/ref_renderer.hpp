// oracle/ref_renderer.hpp — TEST INFRASTRUCTURE ONLY.  Not part of the product; never linked into
// libfriendship_b200.so.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may build or call this.
//
// CPU restatement (C++17, scalar, single-threaded) of the reference's `RefRenderer`
// (reference: src/render/reference.rs, all 291 lines): per-sample recursive pull, no memoisation,
// f32 arithmetic, full external-input history.  Each function cites the lines it follows.
//
// Parity pin: the reference cannot be compiled here (no rustc/cargo; needs 2017 nightly + LLVM 3.8,
// SURVEY.md F4), so this oracle is pinned by the reference's own 11 integration tests' literal
// expected arrays (tests/golden/reference_tests.json) and otherwise by source
// reading.  Behaviour NOT covered by those tests (negative/NaN/huge/time-varying delay, NaN Minimum,
// negative Modulo divisor, nested-in-nested effects, del_*) is "parity pinned by source reading only".
//
// Extension nodes (OscBank / DirectForm / FbDelay) do not exist in the reference:
// their oracles are an fp64 closed form / fp64 sequential recurrence, plus a "reference-style" f32
// per-sample variant.  PARITY UNPINNED by any reference test for those.
//
// Build: g++ -O2 -std=c++17 -ffp-contract=off -fno-fast-math  (no FMA contraction: Rust does not contract).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

namespace oracle {

// reference src/routing/routegraph.rs:38-44 (+ EdgeWeight :20-25); handle 0 == toplevel (nullable_int.rs:27-31)
struct Edge {
    uint32_t from, to, from_slot, to_slot;
};

// reference src/routing/effect.rs:86-112; numbering matches include/friendship_b200.h FRB_KIND_*
enum Kind : uint32_t {
    Delay = 0, F32Constant = 1, Sum2 = 2, Multiply = 3, Divide = 4, Modulo = 5, Minimum = 6,
    UserEffect = 16, OscBank = 32, DirectForm = 33, FbDelay = 34,
};

// A "panic" in the reference (assert!/unwrap/expect) becomes this exception; the C shim maps it to a status.
struct Panic : std::runtime_error {
    int code;
    Panic(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

struct OscBankDef {
    double sample_rate = 48000.0;
    std::vector<uint64_t> voice_offsets;
    std::vector<double> freq_hz;
    std::vector<float> amp, phase, attack, tau;
};
struct DirectFormDef {
    std::vector<float> b0, b1, b2, a1, a2;
};
struct FbDelayDef {
    std::vector<uint32_t> delay;
    std::vector<float> gain;
};

enum class ExtMode { FP64, F32_REFSTYLE };

struct NodeMap;

// reference.rs:31-44  (Node, MyNodeData)
struct Node {
    uint32_t kind = F32Constant;
    std::shared_ptr<NodeMap> user;                   // MyNodeData::UserNode(NodeMap), deep copy (reference.rs:98-113)
    std::shared_ptr<const OscBankDef> osc;
    std::shared_ptr<const DirectFormDef> df;
    std::shared_ptr<const FbDelayDef> fb;
    std::vector<std::optional<Edge>> inbound;        // indexed by to_slot
    // Recurrence memo for the extension recurrences only (they are defined sequentially from t = 0);
    // cleared whenever the graph or the input history changes.  lane -> y history (fp64 or f32 kept as double).
    mutable std::unordered_map<uint32_t, std::vector<double>> memo_y;
    mutable std::unordered_map<uint32_t, std::vector<float>> memo_x;
};

using GetInput = std::function<float(uint64_t, uint32_t)>;

// reference.rs:14-18
struct NodeMap {
    std::unordered_map<uint32_t, Node> nodes;
    std::vector<std::optional<Edge>> output_edges;
    ExtMode ext_mode = ExtMode::FP64;
    bool sparkle_delay = false;
    bool sparkle_min = false;

    // reference.rs:141-153
    void add_edge(const Edge& e) {
        std::vector<std::optional<Edge>>* inbound;
        if (e.to == 0) {
            inbound = &output_edges;
        } else {
            auto it = nodes.find(e.to);
            if (it == nodes.end()) throw Panic(-1, "add_edge: no such node (reference.rs:145 unwrap)");
            inbound = &it->second.inbound;
        }
        size_t slot = e.to_slot;
        if (inbound->size() <= slot) inbound->resize(slot + 1);
        (*inbound)[slot] = e;
    }

    // reference.rs:158-161
    float get_output(uint64_t time, uint32_t slot, const GetInput& get_input) const {
        const std::optional<Edge>* out_edge = slot < output_edges.size() ? &output_edges[slot] : nullptr;
        return get_maybe_edge_value(time, out_edge, get_input);
    }

    // reference.rs:164-173
    float get_maybe_edge_value(uint64_t time, const std::optional<Edge>* maybe_edge, const GetInput& get_input) const {
        if (maybe_edge && maybe_edge->has_value()) return get_edge_value(time, **maybe_edge, get_input);
        return 0.0f;
    }

    static const std::optional<Edge>* slot_of(const Node& n, size_t i) {
        return i < n.inbound.size() ? &n.inbound[i] : nullptr;
    }

    // reference.rs:178-266
    float get_edge_value(uint64_t time, const Edge& edge, const GetInput& get_input) const {
        uint32_t from = edge.from;
        uint32_t from_slot = edge.from_slot;
        if (from == 0) return get_input(time, from_slot);                      // :181-183
        auto it = nodes.find(from);
        if (it == nodes.end()) throw Panic(-1, "edge from unknown node (reference.rs:186 index panic)");
        const Node& node = it->second;
        auto in = [&](size_t i, uint64_t t) { return get_maybe_edge_value(t, slot_of(node, i), get_input); };
        switch (node.kind) {
            case UserEffect: {                                                  // :188-194
                return node.user->get_output(time, from_slot, [&](uint64_t time2, uint32_t slot2) {
                    return get_maybe_edge_value(time2, slot_of(node, slot2), get_input);
                });
            }
            case Delay: {                                                       // :197-216
                if (from_slot != 0) throw Panic(-4, "Delay: from_slot != 0 (reference.rs:199)");
                float delay_frames = in(1, time);
                if (delay_frames >= 18446744073709551616.0f) return 0.0f;      // :202-205
                // SparkleRenderer's variant (reference sparkle.rs:525-542): `fp ULT 0` (negative OR NaN) returns 0.0
                if (sparkle_delay && !(delay_frames >= 0.0f)) return 0.0f;
                uint64_t delay_int;
                if (delay_frames < 0.0f) delay_int = 0;                         // :206-207
                else if (delay_frames != delay_frames) delay_int = 0;           // NaN: Rust saturating `as u64` gives 0
                else delay_int = (uint64_t)delay_frames;                        // :208-210 truncation toward zero
                if (time < delay_int) return 0.0f;                              // :212 checked_sub -> None -> 0
                return in(0, time - delay_int);
            }
            case F32Constant: {                                                 // :217-220
                float f; std::memcpy(&f, &from_slot, 4); return f;
            }
            case Multiply: {                                                    // :221-227
                if (from_slot != 0) throw Panic(-4, "Multiply: from_slot != 0 (reference.rs:223)");
                float l = in(0, time), r = in(1, time);
                return l * r;
            }
            case Sum2: {                                                        // :228-234
                if (from_slot != 0) throw Panic(-4, "Sum2: from_slot != 0 (reference.rs:230)");
                float l = in(0, time), r = in(1, time);
                return l + r;
            }
            case Divide: {                                                      // :235-241
                if (from_slot != 0) throw Panic(-4, "Divide: from_slot != 0 (reference.rs:237)");
                float l = in(0, time), r = in(1, time);
                return l / r;
            }
            case Minimum: {                                                     // :242-248  f32::min == IEEE minNum
                if (from_slot != 0) throw Panic(-4, "Minimum: from_slot != 0 (reference.rs:244)");
                float l = in(0, time), r = in(1, time);
                // SparkleRenderer's variant (reference sparkle.rs:492-498): select(fcmp ULT l r, l, r) — ULT is
                // "unordered or less than", so a NaN in either operand yields l
                if (sparkle_min) return !(l >= r) ? l : r;
                return std::fmin(l, r);
            }
            case Modulo: {                                                      // :249-262
                if (from_slot != 0) throw Panic(-4, "Modulo: from_slot != 0 (reference.rs:251)");
                float dividend = in(0, time), divisor = in(1, time);
                float rem = std::fmod(dividend, divisor);                       // Rust `%` on f32 == fmodf
                if (rem < 0.0f) return rem + divisor;
                return rem;
            }
            case OscBank: return osc_value(node, time, from_slot);
            case DirectForm: return directform_value(node, time, from_slot, get_input);
            case FbDelay: return fbdelay_value(node, time, from_slot, get_input);
        }
        throw Panic(-6, "unknown node kind");
    }

    // ---------------- extension oracles (no reference counterpart; SURVEY.md §8c "Extension oracle") -------------
    // out_v(t) = sum_p amp_p * min(t/A_p,1) * exp(-t/tau_p) * sin(2 pi f_p t / sr + phi_p)
    float osc_value(const Node& node, uint64_t time, uint32_t voice) const {
        const OscBankDef& b = *node.osc;
        if (voice + 1 >= b.voice_offsets.size()) return 0.0f;
        const double two_pi = 6.283185307179586476925286766559;
        double t = (double)time;
        if (ext_mode == ExtMode::FP64) {
            double acc = 0.0;
            for (uint64_t p = b.voice_offsets[voice]; p < b.voice_offsets[voice + 1]; p++) {
                double env = 1.0;
                if (b.attack[p] > 0.0f) env = std::fmin(t / (double)b.attack[p], 1.0);
                if (b.tau[p] > 0.0f && std::isfinite(b.tau[p])) env *= std::exp(-t / (double)b.tau[p]);
                // exact-ish range reduction: turns = frac(f/sr * t) via fmod on the product in long double
                long double turns = (long double)b.freq_hz[p] / (long double)b.sample_rate * (long double)time;
                turns -= floorl(turns);
                acc += (double)b.amp[p] * env * std::sin(two_pi * (double)turns + (double)b.phase[p]);
            }
            return (float)acc;
        }
        // reference-style: what RefRenderer would do had the primitive existed — f32, one sinf per partial-sample
        float acc = 0.0f;
        float tf = (float)time;
        for (uint64_t p = b.voice_offsets[voice]; p < b.voice_offsets[voice + 1]; p++) {
            float env = 1.0f;
            if (b.attack[p] > 0.0f) env = std::fmin(tf / b.attack[p], 1.0f);
            if (b.tau[p] > 0.0f && std::isfinite(b.tau[p])) env *= std::exp(-tf / b.tau[p]);
            double turns = b.freq_hz[p] / b.sample_rate * t;
            turns -= std::floor(turns);
            acc += b.amp[p] * env * std::sin((float)(two_pi * turns) + b.phase[p]);
        }
        return acc;
    }

    // y[n] = b0 x[n] + b1 x[n-1] + b2 x[n-2] - a1 y[n-1] - a2 y[n-2]; sequential from n = 0.
    float directform_value(const Node& node, uint64_t time, uint32_t lane, const GetInput& get_input) const {
        const DirectFormDef& d = *node.df;
        if (lane >= d.b0.size()) return 0.0f;
        auto& ys = node.memo_y[lane];
        auto& xs = node.memo_x[lane];
        while (ys.size() <= time) {
            uint64_t n = ys.size();
            float x0 = get_maybe_edge_value(n, slot_of(node, lane), get_input);
            xs.push_back(x0);
            if (ext_mode == ExtMode::FP64) {
                double x1 = n >= 1 ? xs[n - 1] : 0.0, x2 = n >= 2 ? xs[n - 2] : 0.0;
                double y1 = n >= 1 ? ys[n - 1] : 0.0, y2 = n >= 2 ? ys[n - 2] : 0.0;
                ys.push_back((double)d.b0[lane] * x0 + (double)d.b1[lane] * x1 + (double)d.b2[lane] * x2
                             - (double)d.a1[lane] * y1 - (double)d.a2[lane] * y2);
            } else {
                float x1 = n >= 1 ? xs[n - 1] : 0.0f, x2 = n >= 2 ? xs[n - 2] : 0.0f;
                float y1 = n >= 1 ? (float)ys[n - 1] : 0.0f, y2 = n >= 2 ? (float)ys[n - 2] : 0.0f;
                float y = d.b0[lane] * x0;
                y = y + d.b1[lane] * x1;
                y = y + d.b2[lane] * x2;
                y = y - d.a1[lane] * y1;
                y = y - d.a2[lane] * y2;
                ys.push_back((double)y);
            }
        }
        return (float)ys[time];
    }

    // y[n] = x[n] + g y[n-D]; sequential from n = 0.
    float fbdelay_value(const Node& node, uint64_t time, uint32_t lane, const GetInput& get_input) const {
        const FbDelayDef& d = *node.fb;
        if (lane >= d.delay.size()) return 0.0f;
        auto& ys = node.memo_y[lane];
        uint64_t D = d.delay[lane];
        while (ys.size() <= time) {
            uint64_t n = ys.size();
            float x0 = get_maybe_edge_value(n, slot_of(node, lane), get_input);
            if (ext_mode == ExtMode::FP64) {
                double yd = n >= D ? ys[n - D] : 0.0;
                ys.push_back((double)x0 + (double)d.gain[lane] * yd);
            } else {
                float yd = n >= D ? (float)ys[n - D] : 0.0f;
                float prod = d.gain[lane] * yd;
                ys.push_back((double)(x0 + prod));
            }
        }
        return (float)ys[time];
    }

    void clear_memo() {
        for (auto& kv : nodes) {
            kv.second.memo_y.clear();
            kv.second.memo_x.clear();
            if (kv.second.user) kv.second.user->clear_memo();
        }
    }
    void set_ext_mode(ExtMode m) {
        ext_mode = m;
        for (auto& kv : nodes) if (kv.second.user) kv.second.user->set_ext_mode(m);
    }
    void set_sparkle_delay(bool f) {
        sparkle_delay = f;
        for (auto& kv : nodes) if (kv.second.user) kv.second.user->set_sparkle_delay(f);
    }
    void set_sparkle_min(bool f) {
        sparkle_min = f;
        for (auto& kv : nodes) if (kv.second.user) kv.second.user->set_sparkle_min(f);
    }
};

// One external-input slot vector of the reference (`inputs[slot]: Vec<f32>`, reference.rs:22-25), stored as
// "base zeros, then data": identical reads, but a seek (which rewrites every slot to `idx` zeros, :52-58) is O(1).
struct InputSlot {
    uint64_t base = 0;            // number of leading zeros
    std::vector<float> data;      // values for times [base, base + data.size())
    uint64_t len() const { return base + data.size(); }
    float at(uint64_t t) const { return (t >= base && t - base < data.size()) ? data[t - base] : 0.0f; }
    float last_or_zero() const { return data.empty() ? 0.0f : data.back(); }   // Vec::last of base zeros is 0 too
};

// reference.rs:20-29
class RefRenderer {
public:
    NodeMap nodes;
    // The reference grows `inputs` to buff.len() (= n_slots * n_times, :60) slot vectors.  Only slots that are
    // ever fed a row hold data, so the tail is kept as (count, base) epochs and materialised on first feed.
    std::vector<InputSlot> inputs;                              // materialised prefix
    uint64_t n_slot_vectors = 0;                                // what `self.inputs.len()` is in the reference
    std::vector<std::pair<uint64_t, uint64_t>> epochs;          // (slot index upper bound, base) for the tail
    uint64_t head = 0;                                          // reference.rs:26-28
    ExtMode ext_mode = ExtMode::FP64;
    bool sparkle_delay = false;      // evaluate Delay the way SparkleRenderer does (sparkle.rs:525-542)
    bool sparkle_min = false;        // evaluate Minimum the way SparkleRenderer does (sparkle.rs:492-498)

    // definitions registry (the reference passes Rc<Effect>; the C ABI passes keys)
    std::unordered_map<uint64_t, std::shared_ptr<NodeMap>> effect_defs;
    std::unordered_map<uint64_t, std::shared_ptr<const OscBankDef>> osc_defs;
    std::unordered_map<uint64_t, std::shared_ptr<const DirectFormDef>> df_defs;
    std::unordered_map<uint64_t, std::shared_ptr<const FbDelayDef>> fb_defs;

    // reference.rs:98-113 make_node (deep copy of nested graphs)
    Node make_node(uint32_t kind, uint64_t key) const {
        Node n;
        n.kind = kind;
        switch (kind) {
            case Delay: case F32Constant: case Sum2: case Multiply: case Divide: case Modulo: case Minimum: break;
            case UserEffect: {
                auto it = effect_defs.find(key);
                if (it == effect_defs.end()) throw Panic(-1, "unknown effect definition key");
                n.user = deep_copy(*it->second);
                break;
            }
            case OscBank: {
                auto it = osc_defs.find(key);
                if (it == osc_defs.end()) throw Panic(-1, "unknown oscbank key");
                n.osc = it->second; break;
            }
            case DirectForm: {
                auto it = df_defs.find(key);
                if (it == df_defs.end()) throw Panic(-1, "unknown directform key");
                n.df = it->second; break;
            }
            case FbDelay: {
                auto it = fb_defs.find(key);
                if (it == fb_defs.end()) throw Panic(-1, "unknown fbdelay key");
                n.fb = it->second; break;
            }
            default: throw Panic(-6, "unknown node kind");
        }
        return n;
    }
    static std::shared_ptr<NodeMap> deep_copy(const NodeMap& src) {
        auto dst = std::make_shared<NodeMap>();
        dst->output_edges = src.output_edges;
        for (auto& kv : src.nodes) {
            Node n;
            n.kind = kv.second.kind;
            n.osc = kv.second.osc; n.df = kv.second.df; n.fb = kv.second.fb;
            n.inbound = kv.second.inbound;
            if (kv.second.user) n.user = deep_copy(*kv.second.user);
            dst->nodes.emplace(kv.first, std::move(n));
        }
        return dst;
    }
    void define_effect(uint64_t key, const std::vector<std::pair<uint32_t, std::pair<uint32_t, uint64_t>>>& nds,
                       const std::vector<Edge>& edges) {
        auto nm = std::make_shared<NodeMap>();
        for (auto& n : nds) nm->nodes.emplace(n.first, make_node(n.second.first, n.second.second));  // :103-105
        for (auto& e : edges) nm->add_edge(e);                                                       // :106-108
        effect_defs[key] = nm;
    }

    // reference.rs:116-137 GraphWatcher
    void on_add_node(uint32_t handle, uint32_t kind, uint64_t key) {
        Node n = make_node(kind, key);
        nodes.nodes.erase(handle);             // HashMap::insert replaces
        nodes.nodes.emplace(handle, std::move(n));
        nodes.clear_memo();
    }
    void on_del_node(uint32_t handle) { nodes.nodes.erase(handle); nodes.clear_memo(); }
    void on_add_edge(const Edge& e) { nodes.add_edge(e); nodes.clear_memo(); }
    void on_del_edge(const Edge& e) {
        std::vector<std::optional<Edge>>* inbound;
        if (e.to == 0) inbound = &nodes.output_edges;
        else {
            auto it = nodes.nodes.find(e.to);
            if (it == nodes.nodes.end()) throw Panic(-1, "Attempt to delete edge, but it was never created! (reference.rs:131)");
            inbound = &it->second.inbound;
        }
        if (e.to_slot < inbound->size()) (*inbound)[e.to_slot].reset();
        nodes.clear_memo();
    }

    // reference.rs:46-86.  rows: jagged input rows.  buff: row-major [n_slots x n_times].
    void fill_buffer(float* buff, uint32_t n_slots, uint64_t n_times, uint64_t idx,
                     const std::vector<std::vector<float>>& rows) {
        nodes.set_ext_mode(ext_mode);
        nodes.set_sparkle_delay(sparkle_delay);
        nodes.set_sparkle_min(sparkle_min);
        nodes.clear_memo();
        if (idx != head) {                                                      // :52-58 seek: every slot := idx zeros
            for (auto& slot : inputs) { slot.base = idx; slot.data.clear(); }
            epochs.clear();
            if (n_slot_vectors > inputs.size()) epochs.emplace_back(n_slot_vectors, idx);
        }
        uint64_t buff_len = (uint64_t)n_slots * n_times;                        // ndarray .len() = element count (:60)
        if (n_slot_vectors < buff_len) {                                        // :60-65 new slots hold idx zeros
            epochs.emplace_back(buff_len, idx);
            n_slot_vectors = buff_len;
        }
        // The reference panics (process abort) in the middle of the zip below; a library cannot, so the two
        // asserts are evaluated for every row first and the call is refused with the state left as it was.
        for (size_t r = 0; r < rows.size() && r < n_slot_vectors; r++) {
            materialise_upto(r);
            if (inputs[r].len() != idx) throw Panic(-3, "input slot length != idx (reference.rs:69 assert_eq)");
            if (rows[r].size() > n_times) throw Panic(-2, "cannot send inputs ahead of outputs (reference.rs:71 assert)");
        }
        for (size_t r = 0; r < rows.size() && r < n_slot_vectors; r++) {        // :66-74 zip(rows, slots)
            InputSlot& dest = inputs[r];
            dest.data.insert(dest.data.end(), rows[r].begin(), rows[r].end());
            float pad_val = dest.last_or_zero();                                // :72
            dest.data.resize(idx + n_times - dest.base, pad_val);               // :73
        }
        auto get_input = [this](uint64_t time2, uint32_t slot2) -> float {      // :90-96
            return slot2 < inputs.size() ? inputs[slot2].at(time2) : 0.0f;
        };
        for (uint32_t slot = 0; slot < n_slots; slot++)                         // :78-82
            for (uint64_t time = idx; time < idx + n_times; time++)
                buff[(uint64_t)slot * n_times + (time - idx)] = nodes.get_output(time, slot, get_input);
        head = idx + n_times;                                                   // :84
    }

private:
    void materialise_upto(size_t r) {
        while (inputs.size() <= r) {
            size_t s = inputs.size();
            uint64_t base = 0;
            for (auto& ep : epochs) if (s < ep.first) { base = ep.second; break; }
            InputSlot sl; sl.base = base;
            inputs.push_back(std::move(sl));
        }
    }
};

}  // namespace oracle
