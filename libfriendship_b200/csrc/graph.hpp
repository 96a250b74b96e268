// graph.hpp — host-side mirror of the routed effect tree, kept in sync through the GraphWatcher callbacks.
// Same observable bookkeeping as the reference's renderer-side mirror (reference src/render/reference.rs:14-44,
// :98-153): nodes by handle, inbound edges indexed by to_slot, output edges indexed by toplevel to_slot,
// nested effects deep-copied at add time, deleted edges leave a hole.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <optional>
#include <string>
#include <vector>

#include "../../include/friendship_b200.h"

namespace frb {

struct Error {
    int code;
    std::string msg;
};

struct OscBankDef;      // osc.hpp
struct DirectFormDef;   // scan.hpp
struct FbDelayDef;

struct Graph;

struct GraphNode {
    uint32_t kind = FRB_KIND_F32CONSTANT;
    uint64_t key = 0;
    std::shared_ptr<const Graph> body;               // FRB_KIND_EFFECT: private deep copy (reference.rs:98-113)
    std::vector<std::optional<frb_edge>> inbound;    // by to_slot (reference.rs:35)
};

struct Graph {
    // std::map (ordered by handle) rather than a hash map: traversal order is deterministic.
    std::map<uint32_t, GraphNode> nodes;
    std::vector<std::optional<frb_edge>> output_edges;   // reference.rs:17

    // reference.rs:141-153.  Returns false when the target node does not exist (reference: unwrap panic).
    bool add_edge(const frb_edge& e) {
        std::vector<std::optional<frb_edge>>* inbound;
        if (e.to == 0) {
            inbound = &output_edges;
        } else {
            auto it = nodes.find(e.to);
            if (it == nodes.end()) return false;
            inbound = &it->second.inbound;
        }
        if (inbound->size() <= e.to_slot) inbound->resize((size_t)e.to_slot + 1);
        (*inbound)[e.to_slot] = e;
        return true;
    }
    // reference.rs:127-136: the slot becomes None; the vector never shrinks.
    bool del_edge(const frb_edge& e) {
        std::vector<std::optional<frb_edge>>* inbound;
        if (e.to == 0) {
            inbound = &output_edges;
        } else {
            auto it = nodes.find(e.to);
            if (it == nodes.end()) return false;
            inbound = &it->second.inbound;
        }
        if (e.to_slot < inbound->size()) (*inbound)[e.to_slot].reset();
        return true;
    }
    std::shared_ptr<Graph> deep_copy() const {
        auto g = std::make_shared<Graph>();
        g->output_edges = output_edges;
        for (auto& kv : nodes) {
            GraphNode n = kv.second;
            if (n.body) n.body = n.body->deep_copy();
            g->nodes.emplace(kv.first, std::move(n));
        }
        return g;
    }
};

}  // namespace frb
