// schedule.hpp — the flattened, topologically ordered, device-resident schedule of primitive nodes.
//
// The reference evaluates the routed effect tree by recursive pull, one sample at a time
// (reference src/render/reference.rs:158-266).  Here the tree is flattened once per graph edit into
//   (1) a value DAG in deterministic dependency-first order (the successor of
//       RouteGraph::iter_nodes_dep_first, reference src/routing/routegraph.rs:105-126),
//   (2) stages: every Delay whose source is a computed signal cuts the DAG, the source is materialised
//       into a ring buffer in HBM and the Delay becomes an indexed read at t - d,
//   (3) one register program per stage, interpreted by a single fused kernel over a block of time.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace frb {

// ---- value DAG ----
enum ValueOp : uint8_t {
    V_ZERO = 0,     // unconnected input / missing edge: 0.0f (reference.rs:164-173)
    V_CONST = 1,    // F32Constant, imm = bits (reference.rs:217-220)
    V_INPUT = 2,    // external input slot imm (reference.rs:181-183, :90-96)
    V_DELAY = 3,    // a = source, b = frames (reference.rs:197-216)
    V_SUM2 = 4,     // a + b
    V_MUL = 5,      // a * b
    V_DIV = 6,      // a / b
    V_MOD = 7,      // true modulo (reference.rs:249-262)
    V_MIN = 8,      // minNum
    V_EXT = 9,      // output imm of extension instance a
    V_TAP = 10,     // stored signal a (an external input or an extension lane) read at t - shift, 0 for t < shift;
                    // shift = (b << 32) | imm.  Produced by re-evaluating a cheap Delay source at the shifted time
                    // instead of materialising it (flatten.cc)
    V_GATE = 11,    // (t >= threshold) ? a : 0 ; threshold = (b << 32) | imm : the "0 before t = d" of a Delay
};

struct Value {
    uint8_t op;
    uint32_t a, b, imm;
};

enum ExtKind : uint8_t { EXT_OSCBANK = 0, EXT_DIRECTFORM = 1, EXT_FBDELAY = 2 };

struct ExtInstance {
    uint8_t kind;
    uint64_t key;                    // definition key
    uint32_t n_lanes;                // outputs (and inputs, for the filters)
    std::vector<uint32_t> inputs;    // value id per lane (filters only)
    uint32_t stage = 0;
    uint32_t first_out_buf = 0;      // buffer id of lane 0 (lanes are consecutive buffers)
    uint32_t first_in_buf = 0;       // buffer id of input lane 0 when inputs are one contiguous EXT range; else per-lane table
    std::vector<uint32_t> in_bufs;   // buffer id per input lane
};

// ---- stage programs ----
// One instruction = 4 u32 words: { op | flags<<8 | dst<<16, a, b, aux }.
enum InstrOp : uint8_t {
    I_END = 0,
    I_ADD = 1, I_MUL = 2, I_DIV = 3, I_MOD = 4, I_MIN = 5,   // dst = a op b
    I_LDIN = 6,      // dst = input[aux](t)
    I_LDBUF = 7,     // dst = buffer[aux](t)
    I_DLY_IN = 8,    // dst = input[aux](t - floor(a))   with the Delay clamps
    I_DLY_BUF = 9,   // dst = buffer[aux](t - floor(a))
    I_DLY_TI = 10,   // dst = (t >= floor(a)) ? b : 0    source is time-invariant (constant expression)
    I_STBUF = 11,    // buffer[aux](t) = a
    I_STOUT = 12,    // out[aux](t - t0) = a
    I_MOV = 13,      // dst = a
    I_TAP_IN = 14,   // dst = input[aux](t - shift), shift = (b << 32) | a, 0 for t < shift
    I_TAP_BUF = 15,  // dst = buffer[aux](t - shift)
    I_GATE = 16,     // dst = (t >= threshold) ? a : 0, threshold = (aux << 32) | b
};
constexpr uint32_t IF_A_IMM = 1u;   // a is an immediate f32 bit pattern, not a register
constexpr uint32_t IF_B_IMM = 2u;

struct Instr {
    uint32_t w0, a, b, aux;
    static Instr make(uint8_t op, uint32_t flags, uint32_t dst, uint32_t a, uint32_t b, uint32_t aux) {
        return Instr{(uint32_t)op | (flags << 8) | (dst << 16), a, b, aux};
    }
};

constexpr uint64_t LOOKBACK_FULL = ~0ull;   // signal-driven delay: the whole history from t = 0 is addressable

struct BufferInfo {
    uint32_t value;          // value id materialised here (or ~0u for ext outputs: see ext/lane)
    uint64_t lookback;       // samples before the block start that must stay addressable; LOOKBACK_FULL = all
    uint32_t ext = ~0u;      // producing extension instance, if any
    uint32_t lane = 0;
};

// A Sum2 chain over consecutive lanes of one extension instance, ((x0 + x1) + x2) + ... : evaluated by a dedicated
// streaming kernel (same left-to-right order, so bit-exact) into its own ring before the stage's program runs.
struct FoldJob {
    uint32_t first_buf, count, out_buf;
};

struct Stage {
    std::vector<uint32_t> ext;      // extension instances launched at the start of this stage (creation order)
    std::vector<FoldJob> folds;     // run after the extension instances, before the program
    std::vector<Instr> program;     // register programs of the stage's strands, each ended by two I_ENDs
    std::vector<uint32_t> strand_offsets;   // first instruction of every strand, then program.size(); strands are
                                            // independent sub-programs run by separate thread blocks (grid.y)
    uint32_t n_regs = 0;
};

struct Schedule {
    std::vector<Value> values;            // dependency-first order
    std::vector<ExtInstance> ext;
    std::vector<uint32_t> outputs;        // value id per output slot
    std::vector<uint32_t> value_stage;    // stage per value
    std::vector<int32_t> value_buffer;    // buffer id per value or -1
    std::vector<BufferInfo> buffers;
    std::vector<Stage> stages;
    uint64_t n_input_slots = 0;           // 1 + highest external input slot read (slot 0xFFFFFFFF is legal: u64)
    uint64_t max_lookback = 0;            // finite part
    bool from_zero = false;               // a recurrence or a signal-driven delay: non-contiguous fills restart at t = 0
    bool full_history = false;            // at least one buffer with LOOKBACK_FULL

    // Flat u32 dump for parity tests of routing order / buffer indexing / delay-line offsets:
    //  [0] magic 'FRBS' [1] n_values [2] n_outputs [3] n_buffers [4] n_stages [5] n_ext [6] n_input_slots [7] flags
    //  values:  n_values  x {op, a, b, imm, stage, buffer(+1, 0 = none)}
    //  outputs: n_outputs x {value}
    //  buffers: n_buffers x {value, lookback_lo, lookback_hi, ext(+1), lane}
    //  stages:  n_stages  x {n_ext, n_instr, n_regs, ext ids..., instr words...}
    std::vector<uint32_t> dump() const;
};

}  // namespace frb
