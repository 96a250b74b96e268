// flatten.hpp — see flatten.cc
#pragma once
#include <functional>

#include "graph.hpp"
#include "schedule.hpp"

namespace frb {

struct FlattenEnv {
    // number of lanes (outputs) of an extension definition, or -1 if (kind, key) is not defined
    std::function<int64_t(uint32_t kind, uint64_t key)> ext_lanes;
    // largest per-lane delay of a feedback-delay definition
    std::function<uint64_t(uint64_t key)> ext_max_delay;
    bool sparkle_delay = false;   // FRB_FLAG_SPARKLE_DELAY: a negative / NaN constant amount makes the Delay output 0
    uint32_t max_regs = 48;   // registers (8 samples each) per thread the interpreter kernel can hold in shared memory
};

// Throws frb::Error.
Schedule flatten(const Graph& top, uint32_t n_slots, const FlattenEnv& env);

}  // namespace frb
