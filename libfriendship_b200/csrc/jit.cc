// jit.cc — stage JIT: one flattened stage program -> one fused sm_100a kernel through NVRTC.
//
// The reference's fast renderer is itself a JIT (`SparkleRenderer`, LLVM MCJIT: one function per effect,
// src/render/sparkle.rs:169-243, finalised lazily at the next render, :271-288).  This is its B200 counterpart for the
// fused elementwise + Delay path: where the interpreter (interp.cu) pays ~25 issued instructions per node per 8
// samples, the generated kernel keeps every intermediate in registers and is bound by HBM.  The generated code calls
// the very same device helpers as the interpreter (interp_device.inc is embedded verbatim), compiled with
// --fmad=false, so the two agree bit for bit (tested), and both agree with the CPU oracle.
//
// NVRTC and the driver API are loaded with dlopen at first use: the library still loads on a machine without them,
// and a stage that cannot be compiled simply stays on the interpreter (still a GPU path; there is no CPU fallback).
#include "jit.hpp"

#include <cuda.h>
#include <dlfcn.h>
#include <nvrtc.h>

#include <algorithm>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <set>
#include <sstream>
#include <thread>

namespace frb {

extern const char* const kInterpDeviceSource;   // generated at build time from interp_device.inc

namespace {

struct Api {
    bool ok = false;
    std::string why;
    // NVRTC
    nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*);
    nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*);
    nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*);
    nvrtcResult (*GetCUBIN)(nvrtcProgram, char*);
    nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*);
    nvrtcResult (*GetProgramLog)(nvrtcProgram, char*);
    nvrtcResult (*DestroyProgram)(nvrtcProgram*);
    // driver
    CUresult (*ModuleLoadData)(CUmodule*, const void*);
    CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*);
    CUresult (*ModuleUnload)(CUmodule);
    CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void**, void**);
    CUresult (*OccupancyMaxActiveBlocks)(int*, CUfunction, int, size_t);
};

template <typename T>
bool sym(void* lib, const char* name, T* out) {
    *out = reinterpret_cast<T>(dlsym(lib, name));
    return *out != nullptr;
}

Api& api(bool need_driver) {
    static Api a;
    static std::once_flag once_rtc, once_drv;
    static bool rtc_ok = false, drv_ok = false;
    std::call_once(once_rtc, [] {
        const char* names[] = {"libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so"};
        void* lib = nullptr;
        for (const char* n : names) if ((lib = dlopen(n, RTLD_NOW | RTLD_LOCAL))) break;
        if (!lib) { a.why = "libnvrtc.so.12 not found"; return; }
        rtc_ok = sym(lib, "nvrtcCreateProgram", &a.CreateProgram) && sym(lib, "nvrtcCompileProgram", &a.CompileProgram) &&
                 sym(lib, "nvrtcGetCUBINSize", &a.GetCUBINSize) && sym(lib, "nvrtcGetCUBIN", &a.GetCUBIN) &&
                 sym(lib, "nvrtcGetProgramLogSize", &a.GetProgramLogSize) && sym(lib, "nvrtcGetProgramLog", &a.GetProgramLog) &&
                 sym(lib, "nvrtcDestroyProgram", &a.DestroyProgram);
        if (!rtc_ok) a.why = "NVRTC symbols missing";
    });
    if (need_driver) {
        std::call_once(once_drv, [] {
            void* lib = dlopen("libcuda.so.1", RTLD_NOW | RTLD_LOCAL);
            if (!lib) { a.why = "libcuda.so.1 not found"; return; }
            drv_ok = sym(lib, "cuModuleLoadData", &a.ModuleLoadData) && sym(lib, "cuModuleGetFunction", &a.ModuleGetFunction) &&
                     sym(lib, "cuModuleUnload", &a.ModuleUnload) && sym(lib, "cuLaunchKernel", &a.LaunchKernel) &&
                     sym(lib, "cuOccupancyMaxActiveBlocksPerMultiprocessor", &a.OccupancyMaxActiveBlocks);
            if (!drv_ok) a.why = "driver API symbols missing";
        });
    }
    a.ok = rtc_ok && (!need_driver || drv_ok);
    return a;
}

std::string reg(uint32_t r, int w) { return "r" + std::to_string(r) + "_" + std::to_string(w); }

// ------------------------------------------------------------------------------------------------------------------
// Loops, not straight lines.  A strand is a straight-line register program; what the reference's vocabulary builds out
// of it is mostly REPETITION: a Sum2 chain over 2,000 partial terms is the same three instructions 2,000 times, the
// 64-voice mix of the cfg4 graph the same five 64 times.  NVRTC's time grows faster than linearly in the straight-line
// code it is handed (500 instructions 16 s, 2,000 more than 5 minutes), so the generator folds every run of like
// instruction groups (a tandem repeat of the instruction SHAPES: same operation on the same registers) into one loop
// whose trip count and per-iteration operands (slot / buffer indices, immediates, shifts, thresholds) come from a
// device-resident operand table.  Runs of like loops fold again (a chain per voice, 64 voices: two levels).  The
// generated code therefore depends on the program's structure only — not on its constants or its repeat counts — so
// two strands of the same structure share one body, and a graph edit that adds a partial or changes a constant finds
// its cubin in the cache.
struct Node {
    bool loop = false;
    uint32_t instr = 0;                        // leaf: index into Stage::program
    uint32_t sym = 0;                          // structure symbol: leaves by shape, loops by their body's symbols
    std::vector<std::vector<Node>> iters;      // loop: every iteration (equal symbol sequences); code comes from iters[0]
};

struct Interner {
    std::map<std::string, uint32_t> ids;
    uint32_t get(const std::string& k) {
        auto it = ids.find(k);
        if (it != ids.end()) return it->second;
        const uint32_t id = (uint32_t)ids.size();
        ids.emplace(k, id);
        return id;
    }
};

bool reads_a_reg(uint32_t op, uint32_t flags) { return !(flags & IF_A_IMM) && op != I_LDIN && op != I_LDBUF && op != I_TAP_IN && op != I_TAP_BUF; }
bool reads_b_reg(uint32_t op, uint32_t flags) { return !(flags & IF_B_IMM) && (op <= I_MIN || op == I_DLY_TI); }

std::string leaf_shape(const Instr& in) {
    const uint32_t op = in.w0 & 0xFFu, flags = (in.w0 >> 8) & 0xFFu;
    char buf[64];
    snprintf(buf, sizeof buf, "%x.%x.%x", in.w0, reads_a_reg(op, flags) ? in.a : 0xffffu, reads_b_reg(op, flags) ? in.b : 0xffffu);
    return buf;
}

constexpr size_t kMaxPeriod = 96;              // longest loop body looked for, in symbols
constexpr size_t kMinSaved = 6;                // a loop must remove at least this many symbols

// One pass of tandem-repeat folding over a symbol sequence: greedy from the left, at every position the SHORTEST period
// (up to max_period) that removes at least kMinSaved symbols.  `work` bounds
// the comparisons (adversarial programs stay straight-line).
std::vector<Node> fold_pass(std::vector<Node>& seq, Interner& syms, size_t max_period, bool* changed, size_t* work) {
    std::vector<Node> out;
    const size_t n = seq.size();
    size_t i = 0;
    while (i < n) {
        size_t best_p = 0, best_r = 0;
        if (*work < (200u << 20)) {
            for (size_t P = 1; P <= max_period && i + 2 * P <= n; P++) {
                if (seq[i].sym != seq[i + P].sym) continue;
                size_t m = 0;
                while (i + P + m < n && seq[i + m].sym == seq[i + P + m].sym) m++;
                *work += m + 1;
                const size_t R = 1 + m / P;
                if (R >= 2 && (R - 1) * P >= kMinSaved) { best_p = P; best_r = R; break; }
            }
        }
        if (!best_p) { out.push_back(std::move(seq[i])); i++; continue; }
        Node L;
        L.loop = true;
        std::string key = "L";
        for (size_t j = 0; j < best_p; j++) key += ":" + std::to_string(seq[i + j].sym);
        L.sym = syms.get(key);
        L.iters.resize(best_r);
        for (size_t r = 0; r < best_r; r++)
            for (size_t j = 0; j < best_p; j++) L.iters[r].push_back(std::move(seq[i + r * best_p + j]));
        out.push_back(std::move(L));
        i += best_p * best_r;
        *changed = true;
    }
    return out;
}

std::vector<Node> fold_strand(const Stage& st, size_t sd, Interner& syms) {
    std::vector<Node> seq;
    for (uint32_t i = st.strand_offsets[sd]; i < st.strand_offsets[sd + 1]; i++) {
        const Instr& in = st.program[i];
        if ((in.w0 & 0xFFu) == I_END) break;
        Node nd;
        nd.instr = i;
        nd.sym = syms.get(leaf_shape(in));
        seq.push_back(std::move(nd));
    }
    // short periods everywhere first (a long period found early would swallow the short runs inside it unrolled),
    // then longer ones over what the earlier passes left: loops count as one symbol whatever their trip count
    size_t work = 0;
    for (size_t cap : {(size_t)4, (size_t)12, (size_t)32, kMaxPeriod}) {
        for (int again = 0; again < 3; again++) {
            bool changed = false;
            seq = fold_pass(seq, syms, cap, &changed, &work);
            if (!changed) break;
        }
    }
    return seq;
}

// operand-table words of one instruction, in the order the generated code reads them: [a][b][aux]
void leaf_words(const Instr& in, std::vector<uint32_t>* w, int* n_a, int* n_b, int* n_x) {
    const uint32_t op = in.w0 & 0xFFu, flags = (in.w0 >> 8) & 0xFFu;
    const bool tap = op == I_TAP_IN || op == I_TAP_BUF;
    const bool wa = (flags & IF_A_IMM) || tap;
    const bool wb = (flags & IF_B_IMM) || tap || op == I_GATE;
    const bool wx = op == I_LDIN || op == I_LDBUF || op == I_STBUF || op == I_STOUT || op == I_DLY_IN || op == I_DLY_BUF || tap || op == I_GATE;
    if (w) { if (wa) w->push_back(in.a); if (wb) w->push_back(in.b); if (wx) w->push_back(in.aux); }
    if (n_a) { *n_a = wa; *n_b = wb; *n_x = wx; }
}

void emit_words(const Stage& st, const std::vector<Node>& seq, std::vector<uint32_t>* w) {
    for (const Node& nd : seq) {
        if (!nd.loop) { leaf_words(st.program[nd.instr], w, nullptr, nullptr, nullptr); continue; }
        w->push_back((uint32_t)nd.iters.size());
        for (const auto& it : nd.iters) emit_words(st, it, w);
    }
}

size_t count_leaves(const std::vector<Node>& seq) {
    size_t n = 0;
    for (const Node& nd : seq) n += nd.loop ? count_leaves(nd.iters[0]) : 1;
    return n;
}
bool has_loop(const std::vector<Node>& seq) {
    for (const Node& nd : seq) if (nd.loop) return true;
    return false;
}
// how often the compiler is asked to unroll an innermost loop: enough iterations side by side that their loads are in
// flight together (a loop iteration is a dependent chain operand word -> descriptor -> samples), small enough that the
// unrolled body stays a few dozen instructions
unsigned unroll_of(const Node& loop) {
    if (has_loop(loop.iters[0])) return 1;
    const size_t n = std::max<size_t>(count_leaves(loop.iters[0]), 1);
    return (unsigned)std::max<size_t>(1, std::min<size_t>(8, 32 / n));
}
size_t code_size_node(const Stage& st, const Node& nd, int W);
bool is_pure_load(const Instr& in) {
    const uint32_t op = in.w0 & 0xFFu;
    return op == I_LDIN || op == I_LDBUF || op == I_TAP_IN || op == I_TAP_BUF;
}
// Unroll factor of the loads-first form (Emitter::seq) for an innermost loop of W-column code, or 1 when the loop does not
// qualify: it must load something, and store nothing (a store may alias a later iteration's load: nothing moves across it).
unsigned hoist_unroll(const Stage& st, const Node& loop, int W) {
    if (has_loop(loop.iters[0])) return 1;
    size_t loads = 0;
    for (const Node& b : loop.iters[0]) {
        const Instr& in = st.program[b.instr];
        const uint32_t op = in.w0 & 0xFFu;
        if (op == I_STBUF || op == I_STOUT) return 1;
        loads += is_pure_load(in);
    }
    static const size_t quads = [] { const char* e = getenv("FRB_JIT_QUADS"); return e ? (size_t)std::max(4, atoi(e)) : (size_t)16; }();   // measurement knob
    if (!loads || loads * (size_t)W > quads) return 1;
    size_t U = std::min<size_t>(8, quads / (loads * (size_t)W));           // ~16 quads (256 B) in flight per thread
    U = std::min<size_t>(U, std::max<size_t>(1, 64 / loop.iters[0].size()));   // ... and an unrolled body of a few dozen instructions
    return (unsigned)std::max<size_t>(U, 2);
}
// statements the compiler sees (unrolled bodies counted as often as they are unrolled): what NVRTC's time depends on
size_t code_size(const Stage& st, const std::vector<Node>& seq, int W) {
    size_t n = 0;
    for (const Node& nd : seq) {
        if (!nd.loop) { n++; continue; }
        const unsigned U = hoist_unroll(st, nd, W);
        const size_t body = code_size(st, nd.iters[0], W);
        n += (U > 1 ? (U + 1) * body : unroll_of(nd) * body) + 1;         // loads-first form: U copies + the remainder loop
    }
    return n;
}

// Registers a node reads and writes (a loop: everything its body does, none of it killing — conservative)
void node_regs(const Stage& st, const Node& nd, std::set<uint32_t>* use, std::set<uint32_t>* def) {
    if (nd.loop) { for (const Node& b : nd.iters[0]) node_regs(st, b, use, def); return; }
    const Instr& in = st.program[nd.instr];
    const uint32_t op = in.w0 & 0xFFu, flags = (in.w0 >> 8) & 0xFFu;
    if (reads_a_reg(op, flags)) use->insert(in.a);
    if (reads_b_reg(op, flags)) use->insert(in.b);
    if (op != I_STBUF && op != I_STOUT) def->insert(in.w0 >> 16);
}

// A long APERIODIC program (what is left after folding is still hundreds of statements) is cut into chunks that become
// __noinline__ device functions: the compiler sees a few dozen statements at a time, so NVRTC's time is linear in the
// program instead of superlinear (400 straight-line statements: 8 s; in chunks of 64: 2.5 s).  Registers that live across
// a cut travel through a per-thread array in local memory: a chunk loads the ones it reads before writing and stores the
// ones it wrote that somebody later reads.
constexpr size_t kInlineMax = 160;      // statements a shape may have as one straight body
constexpr size_t kChunk = 64;           // statements per chunk function

struct Chunk {
    size_t first, count;                // top-level nodes [first, first + count)
    std::set<uint32_t> regs, live_in, store;
};

std::vector<Chunk> make_chunks(const Stage& st, const std::vector<Node>& code, int W) {
    std::vector<Chunk> cs;
    size_t i = 0;
    while (i < code.size()) {
        Chunk c{i, 0, {}, {}, {}};
        size_t sz = 0;
        while (i < code.size()) {
            const size_t n = code[i].loop ? code_size_node(st, code[i], W) : 1;
            if (c.count && sz + n > kChunk) break;
            sz += n; c.count++; i++;
        }
        cs.push_back(std::move(c));
    }
    // per chunk: registers read before any (killing) write in the chunk, registers written
    std::vector<std::set<uint32_t>> use(cs.size()), kill(cs.size()), def(cs.size());
    for (size_t k = 0; k < cs.size(); k++) {
        for (size_t j = cs[k].first; j < cs[k].first + cs[k].count; j++) {
            std::set<uint32_t> u, d;
            node_regs(st, code[j], &u, &d);
            for (uint32_t r : u) { cs[k].regs.insert(r); if (!kill[k].count(r)) use[k].insert(r); }
            for (uint32_t r : d) { cs[k].regs.insert(r); def[k].insert(r); if (!code[j].loop) kill[k].insert(r); }
        }
    }
    std::set<uint32_t> live;
    for (size_t k = cs.size(); k-- > 0;) {
        for (uint32_t r : def[k]) if (live.count(r)) cs[k].store.insert(r);
        for (uint32_t r : kill[k]) live.erase(r);
        for (uint32_t r : use[k]) live.insert(r);
        cs[k].live_in = use[k];
    }
    return cs;
}

size_t code_size_node(const Stage& st, const Node& nd, int W) {
    std::vector<Node> one;
    Node c;
    c.loop = nd.loop; c.instr = nd.instr; c.sym = nd.sym;
    if (nd.loop) c.iters.push_back(nd.iters[0]);
    one.push_back(std::move(c));
    return code_size(st, one, W);
}

struct Emitter {
    const Stage& st;
    std::ostringstream& o;
    int W;
    int next_id = 0;
    int next_desc = 0;

    // hoist: "" = emit the instruction where it stands.  Otherwise the name of a temporary for a PURE LOAD (an instruction
    // whose address depends on the time and the operand table only): phase 0 declares the temporary and issues the load,
    // phase 1 emits everything else in program order, the load's place taken by a copy from the temporary.
    void leaf(const Instr& in, const std::string& cur, size_t* off, const std::string& ind, const std::string& hoist = "", int phase = 1) {
        const uint32_t op = in.w0 & 0xFFu, flags = (in.w0 >> 8) & 0xFFu, dst = in.w0 >> 16;
        int na, nb, nx;
        leaf_words(in, nullptr, &na, &nb, &nx);
        const bool pure_load = is_pure_load(in);
        if (!hoist.empty() && phase == 0 && !pure_load) { *off += (size_t)(na + nb + nx); return; }
        if (!hoist.empty() && phase == 1 && pure_load) {
            *off += (size_t)(na + nb + nx);
            for (int w = 0; w < W; w++) o << ind << reg(dst, w) << " = " << hoist << "_" << w << ";\n";
            return;
        }
        auto word = [&]() { return cur + "[" + std::to_string((*off)++) + "]"; };
        std::string wa, wb, wx;
        if (na) wa = word();
        if (nb) wb = word();
        if (nx) wx = word();
        // the descriptor of the slot / ring this instruction addresses: loaded once, not once per column
        const bool uses_in = op == I_LDIN || op == I_TAP_IN || op == I_DLY_IN;
        const bool uses_buf = op == I_LDBUF || op == I_STBUF || op == I_TAP_BUF || op == I_DLY_BUF;
        std::string desc;
        if (uses_in || uses_buf) {
            desc = "d" + std::to_string(next_desc++);
            o << ind << (uses_in ? "const InputDesc " : "const BufferDesc ") << desc << " = "
              << (uses_in ? "input_desc(p, " + wx + ")" : "p.buffers[" + wx + "]") << ";\n";
        }
        for (int w = 0; w < W; w++) {
            const std::string t = "t_" + std::to_string(w);
            const char* guard = w >= 2 ? "if (two) " : "";
            const std::string a = (flags & IF_A_IMM) ? "f4splat(" + wa + ")" : reg(in.a, w);
            const std::string b = (flags & IF_B_IMM) ? "f4splat(" + wb + ")" : reg(in.b, w);
            const std::string sh = "(((unsigned long long)" + wb + " << 32) | " + wa + ")";
            o << ind;
            const std::string dreg = (!hoist.empty() && phase == 0) ? "const float4 " + hoist + "_" + std::to_string(w) : reg(dst, w);
            switch (op) {
                case I_ADD: o << reg(dst, w) << " = f4add(" << a << ", " << b << ");"; break;
                case I_MUL: o << reg(dst, w) << " = f4mul(" << a << ", " << b << ");"; break;
                case I_DIV: o << reg(dst, w) << " = f4div(" << a << ", " << b << ");"; break;
                case I_MOD: o << reg(dst, w) << " = f4mod(" << a << ", " << b << ");"; break;
                case I_MIN: o << reg(dst, w) << " = f4min(" << a << ", " << b << ", p.sparkle_min);"; break;
                case I_MOV: o << reg(dst, w) << " = " << a << ";"; break;
                case I_LDIN: o << dreg << " = f4ld_in(" << desc << ", " << t << ");"; break;
                case I_LDBUF: o << dreg << " = f4ld_buf(" << desc << ", " << t << ");"; break;
                case I_STBUF: o << guard << "f4st_buf(" << desc << ", " << t << ", " << a << ");"; break;
                case I_STOUT: o << guard << "f4st_out(p, " << wx << ", " << t << ", " << a << ");"; break;
                case I_TAP_IN: o << dreg << " = f4tap_in(" << desc << ", " << t << ", " << sh << ");"; break;
                case I_TAP_BUF: o << dreg << " = f4tap_buf(" << desc << ", " << t << ", " << sh << ");"; break;
                case I_GATE: o << reg(dst, w) << " = f4gate(" << a << ", " << t << ", (((unsigned long long)" << wx << " << 32) | " << wb << "));"; break;
                case I_DLY_IN: o << reg(dst, w) << " = f4delay<0>(" << desc << ", no_buf, " << a << ", z4, " << t << ", p.sparkle_delay);"; break;
                case I_DLY_BUF: o << reg(dst, w) << " = f4delay<1>(no_in, " << desc << ", " << a << ", z4, " << t << ", p.sparkle_delay);"; break;
                case I_DLY_TI: o << reg(dst, w) << " = f4delay<2>(no_in, no_buf, " << a << ", " << b << ", " << t << ", p.sparkle_delay);"; break;
                default: o << "/* unknown op " << op << " */"; break;
            }
            o << "\n";
        }
    }

    void seq(const std::vector<Node>& nodes, const std::string& cur, const std::string& ind) { seq(nodes, 0, nodes.size(), cur, ind); }
    // `cur` is the operand cursor (a `const unsigned*` variable); on return it has been advanced past everything seq read
    void seq(const std::vector<Node>& nodes, size_t first, size_t count, const std::string& cur, const std::string& ind) {
        size_t off = 0;
        for (size_t ni = first; ni < first + count; ni++) {
            const Node& nd = nodes[ni];
            if (!nd.loop) { leaf(st.program[nd.instr], cur, &off, ind); continue; }
            const std::string id = std::to_string(next_id++);
            o << ind << "const unsigned n_" << id << " = " << cur << "[" << off << "];\n";
            o << ind << "const unsigned* q_" << id << " = " << cur << " + " << (off + 1) << ";\n";
            const unsigned U = hoist_unroll(st, nd, W);
            if (U > 1) {
                // Innermost loop that only loads and computes (the terms of a Sum2 chain): an iteration is a dependent
                // chain operand word -> descriptor -> samples -> arithmetic, and the compiler will not move a load across
                // the branches of the iterations before it.  So U iterations at a time, ALL their loads first (into
                // temporaries), then their arithmetic in program order: U x loads x W quads in flight per thread.
                const std::vector<Node>& body = nd.iters[0];
                size_t wpp = 0;
                for (const Node& b : body) { int a, bb, x; leaf_words(st.program[b.instr], nullptr, &a, &bb, &x); wpp += (size_t)(a + bb + x); }
                o << ind << "unsigned i_" << id << " = 0;\n";
                o << ind << "for (; i_" << id << " + " << U << " <= n_" << id << "; i_" << id << " += " << U << ") {\n";
                for (int phase = 0; phase < 2; phase++) {
                    for (unsigned k = 0; k < U; k++) {
                        size_t o2 = (size_t)k * wpp;
                        for (size_t j = 0; j < body.size(); j++)
                            leaf(st.program[body[j].instr], "q_" + id, &o2, ind + "  ", "h" + id + "_" + std::to_string(k) + "_" + std::to_string(j), phase);
                    }
                }
                o << ind << "  q_" << id << " += " << (size_t)U * wpp << ";\n";
                o << ind << "}\n";
                o << ind << "for (; i_" << id << " < n_" << id << "; i_" << id << "++) {\n";
            } else {
                o << ind << "#pragma unroll " << unroll_of(nd) << "\n";
                o << ind << "for (unsigned i_" << id << " = 0; i_" << id << " < n_" << id << "; i_" << id << "++) {\n";
            }
            seq(nd.iters[0], "q_" + id, ind + "  ");
            o << ind << "}\n";
            o << ind << cur << " = q_" << id << ";\n";
            off = 0;
        }
        if (off) o << ind << cur << " += " << off << ";\n";
    }
};

}  // namespace

// The generated kernel of a stage: CUDA source (a function of the program's structure), the operand table it reads
// (per strand: shape, row offset; then the rows) and the statements the compiler sees.
JitProgram jit_generate(const Stage& st) {
    const size_t n_strands = st.strand_offsets.empty() ? 0 : st.strand_offsets.size() - 1;
    Interner syms;
    struct Shape { std::vector<uint32_t> key; std::vector<Node> code; uint32_t nreg = 0; };
    std::vector<Shape> shapes;
    JitProgram out;
    out.table.assign(2 * std::max<size_t>(n_strands, 1), 0);
    for (size_t sd = 0; sd < n_strands; sd++) {
        std::vector<Node> seq = fold_strand(st, sd, syms);
        std::vector<uint32_t> key;
        for (const Node& nd : seq) key.push_back(nd.sym);
        size_t k = 0;
        for (; k < shapes.size(); k++) if (shapes[k].key == key) break;
        out.table[2 * sd] = (uint32_t)k;
        out.table[2 * sd + 1] = (uint32_t)out.table.size();
        emit_words(st, seq, &out.table);
        out.table.push_back(0);                                   // a row is never empty
        if (k == shapes.size()) {
            Shape sh;
            sh.key = std::move(key);
            for (uint32_t i = st.strand_offsets[sd]; i < st.strand_offsets[sd + 1]; i++) {
                const uint32_t op = st.program[i].w0 & 0xFFu;
                if (op == I_END) break;
                if (op != I_STBUF && op != I_STOUT) sh.nreg = std::max(sh.nreg, (st.program[i].w0 >> 16) + 1);
            }
            sh.code = std::move(seq);
            shapes.push_back(std::move(sh));
        }
    }

    std::ostringstream o, funcs;
    funcs << kInterpDeviceSource << "\n";
    o << "extern \"C\" __global__ void __launch_bounds__(128) frb_stage(const InterpParams p, const unsigned* __restrict__ tab) {\n";
    o << "  const InputDesc no_in = {nullptr, 0ull, 0ull};\n  const BufferDesc no_buf = {nullptr, 0ull};\n";
    o << "  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);\n";
    o << "  (void)no_in; (void)no_buf; (void)z4;\n";
    o << "  const unsigned strand = blockIdx.y;\n";
    o << "  const unsigned* const row = tab + tab[2 * strand + 1];\n";
    o << "  switch (tab[2 * strand]) {\n";
    for (size_t k = 0; k < shapes.size(); k++) {
        const Shape& sh = shapes[k];
        const bool chunked = code_size(st, sh.code, 2) > kInlineMax;
        // A stage is a stream and what bounds it is the bytes it keeps in flight: 12 resident CTAs x 128 threads x two
        // 16-byte loads per input are 6 MB on the whole GPU, ~5 TB/s at the loaded DRAM latency (measured: 5.0).  A small
        // program therefore walks TWO groups of 8 samples per iteration (one grid stride apart, so every load instruction
        // still covers 512 contiguous bytes), instruction by instruction, which puts the second group's loads in front of
        // the first group's stores.  Loads never alias this stage's stores (a stage reads what earlier stages wrote).
        const int W = (sh.nreg <= 6 && !chunked) ? 4 : 2;            // float4 columns per thread and iteration
        out.code_instrs += code_size(st, sh.code, W);
        if (k == 0 || (W == 4 ? 2u : 1u) < out.groups_per_thread) out.groups_per_thread = W == 4 ? 2u : 1u;
        o << "  case " << k << ": {\n";
        o << "  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;\n";
        o << "  for (unsigned long long g = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; g < p.n_groups; g += " << (W / 2) << "ull * stride) {\n";
        o << "    const unsigned long long t_0 = p.t_begin + 8ull * g, t_1 = t_0 + 4ull;\n";
        if (W == 4) {
            o << "    const bool two = g + stride < p.n_groups;\n";       // the last iteration may have one group only:
            o << "    const unsigned long long t_2 = p.t_begin + 8ull * (two ? g + stride : g), t_3 = t_2 + 4ull;\n";   // it loads the first twice and stores once
        }
        if (chunked) {
            const std::vector<Chunk> cs = make_chunks(st, sh.code, W);
            o << "    float4 R[" << std::max<uint32_t>(sh.nreg, 1) * 2 << "];\n";
            o << "    const unsigned* q = row;\n";
            for (size_t c = 0; c < cs.size(); c++) {
                const std::string fn = "frb_c" + std::to_string(k) + "_" + std::to_string(c);
                o << "    " << fn << "(R, p, q, t_0, t_1);\n";
                funcs << "static __device__ __noinline__ void " << fn << "(float4* __restrict__ R, const InterpParams& p, const unsigned*& qref, "
                      << "const unsigned long long t_0, const unsigned long long t_1) {\n";
                funcs << "  const InputDesc no_in = {nullptr, 0ull, 0ull};\n  const BufferDesc no_buf = {nullptr, 0ull};\n";
                funcs << "  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);\n  (void)no_in; (void)no_buf; (void)z4;\n";
                funcs << "  const unsigned* q = qref;\n";
                for (uint32_t r : cs[c].regs) {
                    funcs << "  float4 " << reg(r, 0) << ", " << reg(r, 1) << ";\n";
                    if (cs[c].live_in.count(r)) funcs << "  " << reg(r, 0) << " = R[" << 2 * r << "]; " << reg(r, 1) << " = R[" << 2 * r + 1 << "];\n";
                }
                Emitter em{st, funcs, W};
                em.next_id = (int)(1000 * c);
                em.seq(sh.code, cs[c].first, cs[c].count, "q", "  ");
                for (uint32_t r : cs[c].store) funcs << "  R[" << 2 * r << "] = " << reg(r, 0) << "; R[" << 2 * r + 1 << "] = " << reg(r, 1) << ";\n";
                funcs << "  qref = q;\n}\n";
            }
            o << "  }\n  } break;\n";
            continue;
        }
        for (uint32_t r = 0; r < sh.nreg; r++) {
            o << "    float4 " << reg(r, 0);
            for (int w = 1; w < W; w++) o << ", " << reg(r, w);
            o << ";\n";
        }
        o << "    const unsigned* q = row;\n";
        Emitter em{st, o, W};
        em.seq(sh.code, "q", "    ");
        o << "  }\n  } break;\n";
    }
    o << "  default: break;\n  }\n}\n";
    out.source = funcs.str() + o.str();
    return out;
}

std::string jit_generate_source(const Stage& st) { return jit_generate(st).source; }
size_t jit_code_instructions(const Stage& st) { return jit_generate(st).code_instrs; }

// Compiled cubins by source text, process-wide: the source is a function of the program's structure only, so a graph
// edit that changes constants or repeat counts — and every renderer that builds the same graph — compiles nothing.
namespace {
std::mutex g_cache_mu;
std::map<std::string, std::shared_ptr<const std::string>> g_cubin_cache;
size_t g_cache_bytes = 0;
std::mutex g_jobs_mu;
std::condition_variable g_jobs_cv;
int g_jobs_in_flight = 0;
}  // namespace

std::shared_ptr<const std::string> jit_cache_lookup(const std::string& source) {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    auto it = g_cubin_cache.find(source);
    return it == g_cubin_cache.end() ? nullptr : it->second;
}

static void cache_store(const std::string& source, const std::string& cubin) {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    if (g_cache_bytes > (256u << 20)) { g_cubin_cache.clear(); g_cache_bytes = 0; }
    if (g_cubin_cache.emplace(source, std::make_shared<const std::string>(cubin)).second) g_cache_bytes += source.size() + cubin.size();
}

bool jit_compile_to_cubin(const std::string& source, std::string* cubin, std::string* log) {
    if (auto hit = jit_cache_lookup(source)) { *cubin = *hit; return true; }
    Api& a = api(false);
    if (!a.ok) { if (log) *log = a.why; return false; }
    nvrtcProgram prog;
    if (a.CreateProgram(&prog, source.c_str(), "frb_stage.cu", 0, nullptr, nullptr) != NVRTC_SUCCESS) {
        if (log) *log = "nvrtcCreateProgram failed";
        return false;
    }
    // same arithmetic contract as the nvcc-built interpreter: no FMA contraction, IEEE division, no flush-to-zero
    const char* opts[] = {"--gpu-architecture=sm_100a", "--fmad=false", "--prec-div=true", "--prec-sqrt=true", "--ftz=false",
                          "-std=c++17", "-lineinfo"};
    std::vector<const char*> optv(opts, opts + sizeof(opts) / sizeof(opts[0]));
    if (getenv("FRB_JIT_NO_LINEINFO")) optv.pop_back();
    std::vector<std::string> extra;
    if (const char* e = getenv("FRB_JIT_NVRTC_OPTS")) {            // measurement knob: extra NVRTC options, space separated
        std::istringstream is(e);
        for (std::string w; is >> w;) extra.push_back(w);
        for (const std::string& w : extra) optv.push_back(w.c_str());
    }
    nvrtcResult rc = a.CompileProgram(prog, (int)optv.size(), optv.data());
    size_t ls = 0;
    a.GetProgramLogSize(prog, &ls);
    if (log && ls > 1) { log->resize(ls); a.GetProgramLog(prog, &(*log)[0]); }
    bool ok = rc == NVRTC_SUCCESS;
    if (ok) {
        size_t n = 0;
        ok = a.GetCUBINSize(prog, &n) == NVRTC_SUCCESS && n > 0;
        if (ok) { cubin->resize(n); ok = a.GetCUBIN(prog, &(*cubin)[0]) == NVRTC_SUCCESS; }
    }
    a.DestroyProgram(&prog);
    if (ok) cache_store(source, *cubin);
    if (const char* dir = getenv("FRB_JIT_DUMP")) {                // tuning aid: the last compiled stage's source and cubin
        if (FILE* f = fopen((std::string(dir) + "/frb_stage.cu").c_str(), "w")) { fwrite(source.data(), 1, source.size(), f); fclose(f); }
        if (ok) if (FILE* f = fopen((std::string(dir) + "/frb_stage.cubin").c_str(), "wb")) { fwrite(cubin->data(), 1, cubin->size(), f); fclose(f); }
    }
    return ok;
}

// A compile beside the render loop.  The thread owns a reference to the job, so a renderer that no longer wants the
// result (graph edit, n_slots change) just drops its own: nothing waits for NVRTC, and the cubin still lands in the cache.
std::shared_ptr<JitJob> jit_compile_async(std::string source) {
    auto job = std::make_shared<JitJob>();
    {
        std::lock_guard<std::mutex> lk(g_jobs_mu);
        g_jobs_in_flight++;
    }
    std::thread([job, src = std::move(source)]() {
        std::string cubin, log;
        const bool ok = jit_compile_to_cubin(src, &cubin, &log);
        job->cubin = ok ? std::move(cubin) : std::string();
        job->log = std::move(log);
        job->done.store(ok ? 1 : -1, std::memory_order_release);
        {
            std::lock_guard<std::mutex> lk(g_jobs_mu);
            g_jobs_in_flight--;
        }
        g_jobs_cv.notify_all();
    }).detach();
    return job;
}

// Called when the last renderer of a process goes away (and by tests): no compile thread outlives the library's users.
void jit_wait_idle() {
    std::unique_lock<std::mutex> lk(g_jobs_mu);
    g_jobs_cv.wait(lk, [] { return g_jobs_in_flight == 0; });
}

struct JitKernel {
    CUmodule mod = nullptr;
    CUfunction fn = nullptr;
    unsigned* d_tab = nullptr;      // operand table (device)
    unsigned groups_per_thread = 1;
    int ctas_per_sm = 16;           // resident 128-thread CTAs per SM (register-limited): sizes the grid to ONE wave
};

JitKernel* jit_build(const Stage& st, std::string* err) {
    const JitProgram prog = jit_generate(st);
    std::string cubin, log;
    if (!jit_compile_to_cubin(prog.source, &cubin, &log)) { if (err) *err = "NVRTC: " + log; return nullptr; }
    return jit_load(cubin, prog.table, prog.groups_per_thread, err);
}

JitKernel* jit_load(const std::string& cubin, const std::vector<uint32_t>& table, unsigned groups_per_thread, std::string* err) {
    Api& a = api(true);
    if (!a.ok) { if (err) *err = a.why; return nullptr; }
    auto* k = new JitKernel();
    k->groups_per_thread = groups_per_thread ? groups_per_thread : 1;
    if (a.ModuleLoadData(&k->mod, cubin.data()) != CUDA_SUCCESS || a.ModuleGetFunction(&k->fn, k->mod, "frb_stage") != CUDA_SUCCESS) {
        if (k->mod) a.ModuleUnload(k->mod);
        delete k;
        if (err) *err = "cuModuleLoadData failed";
        return nullptr;
    }
    int occ = 0;
    // 16 CTAs per SM whatever the kernel's occupancy (10 at 48 registers): sizing the grid to exactly one resident wave
    // measured slower (0.397 vs 0.381 ms on the 64-slot elementwise stage; profiles/k2_grid_bounds_ab_r2.txt)
    (void)occ;
    if (const char* e = getenv("FRB_JIT_CTAS_PER_SM")) k->ctas_per_sm = std::max(1, atoi(e));   // measurement knob
    const size_t bytes = std::max<size_t>(table.size(), 1) * sizeof(uint32_t);
    if (cudaMalloc(&k->d_tab, bytes) != cudaSuccess ||
        (!table.empty() && cudaMemcpy(k->d_tab, table.data(), table.size() * sizeof(uint32_t), cudaMemcpyHostToDevice) != cudaSuccess)) {
        if (err) *err = "operand table upload failed";
        jit_free(k);
        return nullptr;
    }
    return k;
}

void jit_free(JitKernel* k) {
    if (!k) return;
    Api& a = api(true);
    if (k->d_tab) cudaFree(k->d_tab);
    if (a.ok && k->mod) a.ModuleUnload(k->mod);
    delete k;
}

bool jit_launch(JitKernel* k, const InterpParams& p, int sm_count, cudaStream_t stream) {
    Api& a = api(true);
    if (!a.ok || !k) return false;
    if (p.n_groups == 0) return true;
    const unsigned long long per_block = 128ull * k->groups_per_thread;          // every thread gets its two groups when the body walks two
    unsigned long long blocks = (p.n_groups + per_block - 1) / per_block;
    // rounded down: never a last wave of a few CTAs
    unsigned long long cap = (unsigned long long)sm_count * (unsigned)k->ctas_per_sm / p.n_strands;
    if (cap < 1) cap = 1;
    if (blocks > cap) blocks = cap;
    InterpParams pp = p;
    const unsigned* tab = k->d_tab;
    void* args[] = {&pp, &tab};
    return a.LaunchKernel(k->fn, (unsigned)blocks, p.n_strands, 1, 128, 1, 1, 0, (CUstream)stream, args, nullptr) == CUDA_SUCCESS;
}

}  // namespace frb
