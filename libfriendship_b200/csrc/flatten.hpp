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
    // Voice sharding over the devices of one renderer (multi.cu): with shard_world > 1 this schedule evaluates the graph
    // RESTRICTED to the oscillator-bank lanes this rank owns (lane mod shard_world == shard_rank) — every other bank lane
    // is the zero signal, and zeros fold through Sum2 / Multiply / Divide / Delay — so the sum of the ranks' outputs is the
    // whole graph's output wherever that output is linear in the bank lanes (lane_use_of_outputs checks it).
    uint32_t shard_rank = 0, shard_world = 1;
    // External-input slots at or beyond this number read as the zero signal.  A graph may name any u32 slot (the
    // reference's RouteGraph does not check toplevel slots and RefRenderer returns 0 for a slot that was never fed,
    // reference.rs:90-96); the renderer sets the cap above every slot that exists (and re-flattens if a call ever feeds
    // a slot beyond it), so the device table can cover every slot a program names and kernels index it unchecked.
    uint32_t input_slot_cap = 1u << 16;
};

// Throws frb::Error.
Schedule flatten(const Graph& top, uint32_t n_slots, const FlattenEnv& env);

// How the outputs of `s` (flattened with shard_world == 1) depend on the oscillator-bank lanes:
//   LANES_LINEAR  every output slot is zero or a homogeneous linear function of the lanes — sums, gains and divisions
//                 by lane-independent signals, delays, zero-state linear filters — so rendering the lanes on different
//                 devices and adding the results is exact up to f32 summation order;
//   LANES_UNUSED  no output depends on a lane (one device renders everything);
//   LANES_OTHER   anything else (a lane through Minimum / Modulo / a product of lanes, or lanes mixed with
//                 lane-independent signals in one output).
enum LaneUse { LANES_LINEAR = 0, LANES_UNUSED = 1, LANES_OTHER = 2 };
LaneUse lane_use_of_outputs(const Schedule& s);

}  // namespace frb
