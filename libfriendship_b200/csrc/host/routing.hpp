// routing.hpp — C++ restatement of the reference's routing layer, the step BEFORE the hot path:
//   EffectId / EffectMeta / EffectDesc / Effect      reference src/routing/effect.rs
//   AdjList                                            reference src/routing/adjlist.rs
//   RouteGraph (validated DAG)                         reference src/routing/routegraph.rs
//   ResMan (where effect definitions are looked up)    reference src/resman.rs
// Same names, argument meaning and error behaviour; Rc<Effect> becomes shared_ptr<const Effect>.
//
// One deliberate deviation: the reference's cycle check cannot fire as written — `is_edge_reachable`
// (routegraph.rs:218-237) has no base case that compares with `target`, so it always returns false and
// `Error::WouldCycle` is never raised.  Its documented intent ("if no path exists from the edge to itself, then it is
// safe to add the edge", routegraph.rs:196-201) is implemented here: an edge is refused with WouldCycle when its
// target node, entered at the edge's to_slot, reaches the edge's source slot through internally connected slots.
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <filesystem>
#include <fstream>
#include <map>
#include <memory>
#include <optional>
#include <set>
#include <sstream>
#include <unordered_set>
#include <string>
#include <vector>

#include "json.hpp"
#include "sha256.hpp"

namespace frb {
namespace host {

using Sha = std::array<uint8_t, 32>;

// reference routegraph.rs:46-62 (+ effect.rs:18-22)
enum class RgError { None = 0, WouldCycle, NodeInUse, NodeExists, SlotAlreadyConnected, NoSuchNode, NoSuchSlot, NoMatchingEffect };

// reference routegraph.rs:38-44; handle 0 == toplevel (nullable_int.rs)
struct Edge {
    uint32_t from = 0, to = 0, from_slot = 0, to_slot = 0;
    bool operator<(const Edge& o) const { return std::tie(from, to, from_slot, to_slot) < std::tie(o.from, o.to, o.from_slot, o.to_slot); }
    bool operator==(const Edge& o) const { return from == o.from && to == o.to && from_slot == o.from_slot && to_slot == o.to_slot; }
    Json to_json() const {
        return Json::object()
            .add("from", Json::object().add("node_handle", Json::integer(from)))
            .add("to", Json::object().add("node_handle", Json::integer(to)))
            .add("weight", Json::object().add("from_slot", Json::integer(from_slot)).add("to_slot", Json::integer(to_slot)));
    }
    static Edge from_json(const Json& j) {
        Edge e;
        e.from = (uint32_t)j.at("from").at("node_handle").as_u64();
        e.to = (uint32_t)j.at("to").at("node_handle").as_u64();
        e.from_slot = (uint32_t)j.at("weight").at("from_slot").as_u64();
        e.to_slot = (uint32_t)j.at("weight").at("to_slot").as_u64();
        return e;
    }
};

// reference effect.rs:86-112
enum class Primitive { Delay, F32Constant, Sum2, Multiply, Divide, Modulo, Minimum };

struct Url {
    std::string text, scheme, path;
    static Url parse(const std::string& s) {
        Url u;
        u.text = s;
        size_t c = s.find(':');
        if (c == std::string::npos) return u;
        u.scheme = s.substr(0, c);
        size_t p = c + 1;
        if (s.compare(p, 2, "//") == 0) {              // skip the authority
            p += 2;
            size_t slash = s.find('/', p);
            p = slash == std::string::npos ? s.size() : slash;
        }
        size_t end = s.find_first_of("?#", p);
        u.path = s.substr(p, end == std::string::npos ? std::string::npos : end - p);
        return u;
    }
};

// reference effect.rs:357-377
inline std::optional<Primitive> primitive_from_url(const Url& u) {
    if (u.scheme != "primitive") return std::nullopt;
    if (u.path == "/Delay") return Primitive::Delay;
    if (u.path == "/F32Constant") return Primitive::F32Constant;
    if (u.path == "/Sum2") return Primitive::Sum2;
    if (u.path == "/Multiply") return Primitive::Multiply;
    if (u.path == "/Divide") return Primitive::Divide;
    if (u.path == "/Modulo") return Primitive::Modulo;
    if (u.path == "/Minimum") return Primitive::Minimum;
    return std::nullopt;
}

// reference effect.rs:26-39
struct EffectId {
    std::string name;
    std::optional<Sha> sha256;
    std::vector<std::string> urls;

    bool is_primitive() const {                                   // effect.rs:244-248
        return urls.size() == 1 && Url::parse(urls[0]).scheme == "primitive";
    }
    std::optional<Primitive> primitive() const {                  // effect.rs:249-255 + from_url
        if (!is_primitive()) return std::nullopt;
        return primitive_from_url(Url::parse(urls[0]));
    }
    Json to_json() const {
        Json sha = Json::null();
        if (sha256) { sha = Json::array(); for (uint8_t b : *sha256) sha.a.push_back(Json::integer(b)); }
        Json u = Json::array();
        for (auto& s : urls) u.a.push_back(Json::string(s));
        return Json::object().add("name", Json::string(name)).add("sha256", sha).add("urls", u);
    }
    static EffectId from_json(const Json& j) {
        EffectId id;
        id.name = j.at("name").as_string();
        const Json& sha = j.at("sha256");
        if (sha.type != Json::Null) {
            auto& arr = sha.as_array();
            if (arr.size() != 32) throw std::runtime_error("sha256 must have 32 bytes");
            Sha s;
            for (int i = 0; i < 32; i++) s[i] = (uint8_t)arr[i].as_u64();
            id.sha256 = s;
        }
        for (auto& u : j.at("urls").as_array()) id.urls.push_back(u.as_string());
        return id;
    }
};

// reference effect.rs:66-72
struct EffectIO {
    std::string name;
    uint8_t channel = 0;
    Json to_json() const { return Json::object().add("name", Json::string(name)).add("channel", Json::integer(channel)); }
    static EffectIO from_json(const Json& j) { return EffectIO{j.at("name").as_string(), (uint8_t)j.at("channel").as_u64()}; }
};

// reference effect.rs:59-64, 297-336
struct EffectMeta {
    EffectId id;
    std::vector<EffectIO> inputs_, outputs_;

    // number of inputs; primitives have fixed signatures (effect.rs:297-314)
    uint64_t n_inputs() const {
        auto p = id.primitive();
        if (!p) return inputs_.size();
        return *p == Primitive::F32Constant ? 0 : 2;
    }
    // number of outputs; F32Constant exposes every u32 but the last (effect.rs:315-321, 390-393)
    uint64_t n_outputs() const {
        auto p = id.primitive();
        if (!p) return outputs_.size();
        return *p == Primitive::F32Constant ? 0xFFFFFFFFull : 1;
    }
    bool is_valid_input(uint32_t slot) const { return slot < n_inputs(); }      // effect.rs:328-330
    bool is_valid_output(uint32_t slot) const { return slot < n_outputs(); }    // effect.rs:331-333
    std::vector<EffectIO> inputs() const {
        auto p = id.primitive();
        if (!p) return inputs_;
        switch (*p) {
            case Primitive::Delay: return {{"source", 0}, {"frames", 0}};
            case Primitive::F32Constant: return {};
            case Primitive::Sum2: case Primitive::Multiply: case Primitive::Minimum: return {{"source", 0}, {"source2", 0}};
            default: return {{"source", 0}, {"divisor", 0}};
        }
    }
    Json to_json() const {
        Json in = Json::array(), out = Json::array();
        for (auto& i : inputs_) in.a.push_back(i.to_json());
        for (auto& o : outputs_) out.a.push_back(o.to_json());
        return Json::object().add("id", id.to_json()).add("inputs", in).add("outputs", out);
    }
    static EffectMeta from_json(const Json& j) {
        EffectMeta m;
        m.id = EffectId::from_json(j.at("id"));
        for (auto& i : j.at("inputs").as_array()) m.inputs_.push_back(EffectIO::from_json(i));
        for (auto& o : j.at("outputs").as_array()) m.outputs_.push_back(EffectIO::from_json(o));
        return m;
    }
};

// reference adjlist.rs:11-16
struct AdjList {
    std::vector<std::pair<uint32_t, EffectId>> nodes;
    std::vector<Edge> edges;
    Json to_json() const {
        Json n = Json::array(), e = Json::array();
        for (auto& kv : nodes) {
            Json pair = Json::array();
            pair.a.push_back(Json::object().add("node_handle", Json::integer(kv.first)));
            pair.a.push_back(kv.second.to_json());
            n.a.push_back(pair);
        }
        for (auto& ed : edges) e.a.push_back(ed.to_json());
        return Json::object().add("nodes", n).add("edges", e);
    }
    static AdjList from_json(const Json& j) {
        AdjList a;
        for (auto& pair : j.at("nodes").as_array()) {
            auto& pr = pair.as_array();
            if (pr.size() != 2) throw std::runtime_error("adjlist node must be [handle, id]");
            a.nodes.emplace_back((uint32_t)pr[0].at("node_handle").as_u64(), EffectId::from_json(pr[1]));
        }
        for (auto& e : j.at("edges").as_array()) a.edges.push_back(Edge::from_json(e));
        return a;
    }
};

// reference effect.rs:43-48
struct EffectDesc {
    EffectMeta meta;
    AdjList adjlist;
    Json to_json() const { return Json::object().add("meta", meta.to_json()).add("adjlist", adjlist.to_json()); }
    static EffectDesc from_json(const Json& j) { return EffectDesc{EffectMeta::from_json(j.at("meta")), AdjList::from_json(j.at("adjlist"))}; }
    // effect.rs:272-281: hash of the re-serialised description when the id carries none
    void update_id() {
        if (!meta.id.sha256) meta.id.sha256 = Sha256::digest(to_json().dump());
    }
};

class ResMan;
class RouteGraph;

// reference effect.rs:50-57, 79-83
struct Effect {
    EffectMeta meta;
    std::optional<Primitive> primitive;            // EffectData::Primitive
    std::shared_ptr<const RouteGraph> graph;       // EffectData::RouteGraph

    bool are_slots_connected(uint32_t from_slot, uint32_t to_slot) const;   // effect.rs:115-121
    static std::shared_ptr<const Effect> from_id(const EffectId& id, const ResMan& resman, RgError* err);   // effect.rs:135-220
};
using NodeData = std::shared_ptr<const Effect>;

// reference resman.rs
class ResMan {
public:
    void add_dir(const std::string& dir) { dirs_.push_back(dir); }                       // resman.rs:34-36
    // resman.rs:39-97: cached path for the hash first, then every regular file of every directory (flat scan);
    // with a hash in the id, only files whose SHA-256 matches
    std::vector<std::string> find_effect(const EffectId& id) const {
        std::vector<std::string> cand;
        if (id.sha256) {
            auto it = cache_.find(*id.sha256);
            if (it != cache_.end()) cand.push_back(it->second);
        }
        for (auto& d : dirs_) {
            std::error_code ec;
            std::vector<std::string> files;
            for (auto it = std::filesystem::directory_iterator(d, ec); !ec && it != std::filesystem::directory_iterator(); it.increment(ec))
                if (it->is_regular_file(ec)) files.push_back(it->path().string());
            std::sort(files.begin(), files.end());     // read_dir order is unspecified; sorted here for determinism
            cand.insert(cand.end(), files.begin(), files.end());
        }
        std::vector<std::string> out;
        for (auto& f : cand) {
            if (id.sha256) {
                std::string bytes;
                if (!read_file(f, &bytes)) continue;
                Sha h = Sha256::digest(bytes);
                cache_[h] = f;
                if (h != *id.sha256) continue;
            }
            out.push_back(f);
        }
        return out;
    }
    static bool read_file(const std::string& path, std::string* out) {
        std::ifstream in(path, std::ios::binary);
        if (!in) return false;
        std::ostringstream ss;
        ss << in.rdbuf();
        *out = ss.str();
        return true;
    }

private:
    std::vector<std::string> dirs_;
    mutable std::map<Sha, std::string> cache_;      // resman.rs:23-27 (RefCell<ResCache>)
};

// reference routegraph.rs:64-79
class RouteGraph {
public:
    struct Node {
        std::set<Edge> outbound, inbound;
        NodeData data;       // null for the toplevel
    };
    RouteGraph() { nodes_[0]; }                                                             // routegraph.rs:82-89

    const std::map<uint32_t, Node>& nodes() const { return nodes_; }
    NodeData get_data(uint32_t handle) const {                                              // :143-145
        auto it = nodes_.find(handle);
        return it == nodes_.end() ? nullptr : it->second.data;
    }
    std::vector<Edge> edges() const {                                                       // iter_edges :127-129
        std::vector<Edge> out;
        for (auto& kv : nodes_) out.insert(out.end(), kv.second.outbound.begin(), kv.second.outbound.end());
        return out;
    }
    const std::set<Edge>& outbound_edges() const { return nodes_.at(0).inbound; }           // :131-134 (edges into outputs)
    const std::set<Edge>& inbound_edges() const { return nodes_.at(0).outbound; }           // :136-138 (edges from inputs)

    // deterministic dependency-first order (reference :105-126 iterates a HashMap)
    std::vector<uint32_t> nodes_dep_first() const {
        std::vector<uint32_t> order;
        std::set<uint32_t> visited;
        for (auto& kv : nodes_) dep_first(kv.first, visited, order);
        return order;
    }

    RgError add_node(uint32_t handle, NodeData data) {                                      // :153-162
        if (nodes_.count(handle)) return RgError::NodeExists;
        nodes_[handle].data = std::move(data);
        return RgError::None;
    }
    RgError add_edge(const Edge& e) {                                                       // :165-208
        auto to = nodes_.find(e.to);
        if (to == nodes_.end()) return RgError::NoSuchNode;
        for (auto& in : to->second.inbound) if (in.to_slot == e.to_slot) return RgError::SlotAlreadyConnected;
        if (to->second.data && !to->second.data->meta.is_valid_input(e.to_slot)) return RgError::NoSuchSlot;
        auto from = nodes_.find(e.from);
        if (from == nodes_.end()) return RgError::NoSuchNode;
        if (from->second.data && !from->second.data->meta.is_valid_output(e.from_slot)) return RgError::NoSuchSlot;
        if (would_cycle(e)) return RgError::WouldCycle;
        nodes_[e.from].outbound.insert(e);
        nodes_[e.to].inbound.insert(e);
        return RgError::None;
    }
    RgError del_node(uint32_t handle) {                                                     // :263-277
        auto it = nodes_.find(handle);
        if (it == nodes_.end()) return RgError::None;
        if (!it->second.outbound.empty() || !it->second.inbound.empty()) return RgError::NodeInUse;
        nodes_.erase(it);
        return RgError::None;
    }
    void del_edge(const Edge& e) {                                                          // :278-285
        auto f = nodes_.find(e.from);
        if (f != nodes_.end()) f->second.outbound.erase(e);
        auto t = nodes_.find(e.to);
        if (t != nodes_.end()) t->second.inbound.erase(e);
    }
    // Is there a path from toplevel input `in_slot` to toplevel output `out_slot`?  (:245-262)
    bool are_slots_connected(uint32_t in_slot, uint32_t out_slot) const {
        for (auto& ef : nodes_.at(0).outbound) {
            if (ef.from_slot != in_slot) continue;
            if (ef.to == 0) { if (ef.to_slot == out_slot) return true; continue; }          // direct pass-through edge
            if (reaches(ef, 0, out_slot)) return true;
        }
        return false;
    }
    AdjList to_adjlist() const {                                                            // :287-303
        AdjList a;
        for (auto& kv : nodes_) if (kv.second.data) a.nodes.emplace_back(kv.first, kv.second.data->meta.id);
        a.edges = edges();
        return a;
    }
    static std::shared_ptr<RouteGraph> from_adjlist(const AdjList& adj, const ResMan& res, RgError* err) {   // :305-326
        auto g = std::make_shared<RouteGraph>();
        for (auto& kv : adj.nodes) {
            NodeData d = Effect::from_id(kv.second, res, err);
            if (!d) return nullptr;
            g->nodes_[kv.first].data = d;
        }
        for (auto& e : adj.edges) {
            RgError r = g->add_edge(e);
            if (r != RgError::None) { *err = r; return nullptr; }
        }
        return g;
    }

private:
    void dep_first(uint32_t h, std::set<uint32_t>& visited, std::vector<uint32_t>& order) const {
        if (h == 0 || visited.count(h)) return;
        visited.insert(h);
        auto it = nodes_.find(h);
        if (it != nodes_.end()) for (auto& e : it->second.inbound) dep_first(e.from, visited, order);
        order.push_back(h);
    }
    // From edge `from` (already at node from.to, slot from.to_slot): can a chain of internally connected slots reach
    // node `target`, arriving at an edge that leaves through... (for target == toplevel: an edge into output `slot`;
    // for a real node: an edge into `target` whose to_slot is internally connected to output `slot` of `target`).
    // Iterative (a chain of tens of thousands of nodes must not be a stack depth), and a state is the (node, input slot)
    // an edge arrives at: two edges into the same slot continue the same way.
    bool reaches(const Edge& start, uint32_t target, uint32_t slot) const {
        std::unordered_set<uint64_t> seen;
        std::vector<Edge> todo{start};
        while (!todo.empty()) {
            const Edge from = todo.back();
            todo.pop_back();
            if (from.to == 0) continue;
            if (!seen.insert(((uint64_t)from.to << 32) | from.to_slot).second) continue;
            auto it = nodes_.find(from.to);
            if (it == nodes_.end()) continue;
            const Node& n = it->second;
            if (from.to == target && target != 0) {
                if (!n.data || n.data->are_slots_connected(from.to_slot, slot)) return true;
            }
            for (auto& cand : n.outbound) {
                if (n.data && !n.data->are_slots_connected(from.to_slot, cand.from_slot)) continue;   // :239-243
                if (target == 0 && cand.to == 0 && cand.to_slot == slot) return true;
                todo.push_back(cand);
            }
        }
        return false;
    }
    bool would_cycle(const Edge& e) const {
        if (e.to == 0 || e.from == 0) return false;      // the graph's own inputs/outputs never close a loop
        if (e.from == e.to) {
            auto it = nodes_.find(e.to);
            return !it->second.data || it->second.data->are_slots_connected(e.to_slot, e.from_slot);
        }
        return reaches(e, e.from, e.from_slot);
    }
    std::map<uint32_t, Node> nodes_;
};

inline bool Effect::are_slots_connected(uint32_t from_slot, uint32_t to_slot) const {
    if (graph) return graph->are_slots_connected(from_slot, to_slot);
    return true;    // primitives: every input feeds every output (effect.rs:118-120)
}

inline std::shared_ptr<const Effect> Effect::from_id(const EffectId& id, const ResMan& resman, RgError* err) {
    auto prim = id.primitive();
    if (prim && !id.sha256) {                                                              // effect.rs:137-151
        auto e = std::make_shared<Effect>();
        e->meta.id = id;
        e->primitive = prim;
        return e;
    }
    // Files whose sub-nodes are being resolved right now, innermost last.  An effect file that names itself (or a file
    // that is loading it) by name only would recurse without end — the reference overflows its stack on such a file;
    // a library inside someone else's process skips the candidate instead, and the id ends up NoMatchingEffect.
    static thread_local std::vector<std::string> loading;
    struct Loading {
        explicit Loading(const std::string& p) { loading.push_back(p); }
        ~Loading() { loading.pop_back(); }
    };
    for (auto& path : resman.find_effect(id)) {                                             // effect.rs:158-216
        if (loading.size() >= 256 || std::find(loading.begin(), loading.end(), path) != loading.end()) continue;
        std::string text;
        if (!ResMan::read_file(path, &text)) continue;
        EffectDesc desc;
        try { desc = EffectDesc::from_json(Json::parse(text)); } catch (const std::exception&) { continue; }   // :213-215 warn, next
        if (desc.meta.id.name != id.name) continue;                                         // :163, :209-211
        desc.update_id();
        RgError sub = RgError::None;
        Loading in_progress(path);
        auto graph = RouteGraph::from_adjlist(desc.adjlist, resman, &sub);
        if (!graph) continue;                                                               // :207 warn, next
        // all declared outputs driven, exactly (:168-175)
        std::vector<uint32_t> real_out;
        for (auto& e : graph->outbound_edges()) real_out.push_back(e.to_slot);
        std::sort(real_out.begin(), real_out.end());
        bool outputs_driven = real_out.size() == desc.meta.outputs_.size();
        for (size_t i = 0; outputs_driven && i < real_out.size(); i++) outputs_driven = real_out[i] == i;
        // every input edge is declared (:179-188)
        bool inputs_valid = true;
        for (auto& e : graph->inbound_edges()) inputs_valid = inputs_valid && e.from_slot < desc.meta.inputs_.size();
        // every sub-node has exactly its declared inputs driven (:189-195)
        bool subnodes_driven = true;
        for (auto& kv : graph->nodes()) {
            if (!kv.second.data) continue;
            std::vector<uint32_t> driven;
            for (auto& e : kv.second.inbound) driven.push_back(e.to_slot);
            std::sort(driven.begin(), driven.end());
            bool ok = driven.size() == kv.second.data->meta.n_inputs();
            for (size_t i = 0; ok && i < driven.size(); i++) ok = driven[i] == i;
            subnodes_driven = subnodes_driven && ok;
        }
        if (inputs_valid && outputs_driven && subnodes_driven) {
            auto e = std::make_shared<Effect>();
            e->meta = desc.meta;
            e->graph = graph;
            return e;
        }
    }
    if (err) *err = RgError::NoMatchingEffect;                                              // :218-219
    return nullptr;
}

}  // namespace host
}  // namespace frb
