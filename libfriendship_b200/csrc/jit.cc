// jit.cc — stage JIT: one flattened stage program -> one fused, straight-line sm_100a kernel through NVRTC.
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
#include <cstdio>
#include <cstring>
#include <mutex>
#include <sstream>

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
                     sym(lib, "cuModuleUnload", &a.ModuleUnload) && sym(lib, "cuLaunchKernel", &a.LaunchKernel);
            if (!drv_ok) a.why = "driver API symbols missing";
        });
    }
    a.ok = rtc_ok && (!need_driver || drv_ok);
    return a;
}

std::string reg(uint32_t r, int w) { return "r" + std::to_string(r) + "_" + std::to_string(w); }

}  // namespace

// ------------------------------------------------------------------------------------------------------------------
// Strands with the same SHAPE (same operations on the same registers; only slot / buffer indices, immediates, shifts
// and thresholds differ) share one code body; what differs comes from a per-strand row of a constant table.  A stage
// that applies the same chain to 64 slots compiles one body, not 64.
// what makes two strands share a code body; *n_ops = instructions before the strand's I_END
static std::string strand_shape_key(const Stage& st, size_t sd, size_t* n_ops) {
    std::string key;
    size_t n = 0;
    for (uint32_t i = st.strand_offsets[sd]; i < st.strand_offsets[sd + 1]; i++) {
        const Instr& in = st.program[i];
        const uint32_t op = in.w0 & 0xFFu, flags = (in.w0 >> 8) & 0xFFu;
        if (op == I_END) break;
        const bool a_reg = !(flags & IF_A_IMM) && op != I_LDIN && op != I_LDBUF && op != I_TAP_IN && op != I_TAP_BUF;
        const bool b_reg = !(flags & IF_B_IMM) && (op <= I_MIN || op == I_DLY_TI);
        char buf[64];
        snprintf(buf, sizeof buf, "%x.%x.%x.%x;", in.w0, a_reg ? in.a : 0xffffu, b_reg ? in.b : 0xffffu, 0u);
        key += buf;
        n++;
    }
    if (n_ops) *n_ops = n;
    return key;
}

// Instructions the generated kernel holds as straight-line code: one body per distinct shape.  NVRTC's time grows faster
// than linearly in it (sm_100a, CUDA 12.9, a Sum2 chain: 50 -> 0.5 s, 100 -> 0.8 s, 200 -> 1.9 s, 500 -> 16 s,
// 2,000 -> more than 5 minutes), so the renderer bounds what it compiles by this number (renderer.cu, poll_stage_jit).
size_t jit_code_instructions(const Stage& st) {
    const size_t n_strands = st.strand_offsets.empty() ? 0 : st.strand_offsets.size() - 1;
    std::vector<std::string> keys;
    size_t total = 0;
    for (size_t sd = 0; sd < n_strands; sd++) {
        size_t n = 0;
        std::string key = strand_shape_key(st, sd, &n);
        bool seen = false;
        for (const std::string& k : keys) if (k == key) { seen = true; break; }
        if (!seen) { keys.push_back(std::move(key)); total += n; }
    }
    return total;
}

std::string jit_generate_source(const Stage& st) {
    const size_t n_strands = st.strand_offsets.empty() ? 0 : st.strand_offsets.size() - 1;
    struct Shape { std::string key; std::vector<uint32_t> strands; };
    std::vector<Shape> shapes;
    std::vector<uint32_t> shape_of(n_strands), idx_of(n_strands);
    for (size_t sd = 0; sd < n_strands; sd++) {
        const std::string key = strand_shape_key(st, sd, nullptr);
        size_t k = 0;
        for (; k < shapes.size(); k++) if (shapes[k].key == key) break;
        if (k == shapes.size()) shapes.push_back(Shape{key, {}});
        shape_of[sd] = (uint32_t)k;
        idx_of[sd] = (uint32_t)shapes[k].strands.size();
        shapes[k].strands.push_back((uint32_t)sd);
    }

    std::ostringstream o, tables;
    o << kInterpDeviceSource << "\n";
    std::ostringstream body;
    body << "extern \"C\" __global__ void __launch_bounds__(128) frb_stage(const InterpParams p) {\n";
    body << "  const InputDesc no_in = {nullptr, 0ull, 0ull};\n  const BufferDesc no_buf = {nullptr, 0ull};\n";
    body << "  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);\n";
    body << "  (void)no_in; (void)no_buf; (void)z4;\n";
    body << "  const unsigned strand = blockIdx.y;\n";
    body << "  switch (frb_shape_of[strand]) {\n";
    for (size_t k = 0; k < shapes.size(); k++) {
        // the varying words of every strand of this shape, in instruction order
        std::vector<std::vector<uint32_t>> rows(shapes[k].strands.size());
        const uint32_t rep = shapes[k].strands[0];
        body << "  case " << k << ": {\n";
        body << "  const unsigned* q = frb_tab" << k << "[frb_idx_of[strand]];\n  (void)q;\n";
        uint32_t nreg = 0;
        for (uint32_t i = st.strand_offsets[rep]; i < st.strand_offsets[rep + 1]; i++) {
            const uint32_t op = st.program[i].w0 & 0xFFu;
            if (op == I_END) break;
            if (op != I_STBUF && op != I_STOUT) nreg = std::max(nreg, (st.program[i].w0 >> 16) + 1);
        }
        // A stage is a stream and what bounds it is the bytes it keeps in flight: 12 resident CTAs x 128 threads x two
        // 16-byte loads per input are 6 MB on the whole GPU, ~5 TB/s at the loaded DRAM latency (measured: 5.0).  A small
        // program therefore walks TWO groups of 8 samples per iteration (one grid stride apart, so every load instruction
        // still covers 512 contiguous bytes), instruction by instruction, which puts the second group's loads in front of
        // the first group's stores.  Loads never alias this stage's stores (a stage reads what earlier stages wrote).
        const int W = nreg <= 6 ? 4 : 2;            // float4 columns per thread and iteration
        body << "  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;\n";
        body << "  for (unsigned long long g = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; g < p.n_groups; g += " << (W / 2) << "ull * stride) {\n";
        body << "    const unsigned long long t_0 = p.t_begin + 8ull * g, t_1 = t_0 + 4ull;\n";
        if (W == 4) {
            body << "    const bool two = g + stride < p.n_groups;\n";       // the last iteration may have one group only:
            body << "    const unsigned long long t_2 = p.t_begin + 8ull * (two ? g + stride : g), t_3 = t_2 + 4ull;\n";   // it loads the first twice and stores once
        }
        for (uint32_t r = 0; r < nreg; r++) {
            body << "    float4 " << reg(r, 0);
            for (int w = 1; w < W; w++) body << ", " << reg(r, w);
            body << ";\n";
        }
        uint32_t n_words = 0;
        const uint32_t len = st.strand_offsets[rep + 1] - st.strand_offsets[rep];
        for (uint32_t j = 0; j < len; j++) {
            const Instr& in = st.program[st.strand_offsets[rep] + j];
            const uint32_t op = in.w0 & 0xFFu, flags = (in.w0 >> 8) & 0xFFu, dst = in.w0 >> 16;
            if (op == I_END) break;
            // table words for this instruction: [a if immediate / shift-lo][b if immediate / shift-hi][aux]
            auto word = [&](int which) {   // 0: a, 1: b, 2: aux
                for (size_t m = 0; m < rows.size(); m++) {
                    const Instr& im = st.program[st.strand_offsets[shapes[k].strands[m]] + j];
                    rows[m].push_back(which == 0 ? im.a : which == 1 ? im.b : im.aux);
                }
                return "q[" + std::to_string(n_words++) + "]";
            };
            const bool tap = op == I_TAP_IN || op == I_TAP_BUF;
            std::string wa, wb, wx;
            if ((flags & IF_A_IMM) || tap) wa = word(0);
            if ((flags & IF_B_IMM) || tap || op == I_GATE) wb = word(1);
            if (op == I_LDIN || op == I_LDBUF || op == I_STBUF || op == I_STOUT || op == I_DLY_IN || op == I_DLY_BUF || tap || op == I_GATE) wx = word(2);
            for (int w = 0; w < W; w++) {
                const std::string t = "t_" + std::to_string(w);
                const char* guard = w >= 2 ? "if (two) " : "";
                const std::string a = (flags & IF_A_IMM) ? "f4splat(" + wa + ")" : reg(in.a, w);
                const std::string b = (flags & IF_B_IMM) ? "f4splat(" + wb + ")" : reg(in.b, w);
                const std::string sh = "(((unsigned long long)" + wb + " << 32) | " + wa + ")";
                body << "    ";
                switch (op) {
                    case I_ADD: body << reg(dst, w) << " = f4add(" << a << ", " << b << ");"; break;
                    case I_MUL: body << reg(dst, w) << " = f4mul(" << a << ", " << b << ");"; break;
                    case I_DIV: body << reg(dst, w) << " = f4div(" << a << ", " << b << ");"; break;
                    case I_MOD: body << reg(dst, w) << " = f4mod(" << a << ", " << b << ");"; break;
                    case I_MIN: body << reg(dst, w) << " = f4min(" << a << ", " << b << ");"; break;
                    case I_MOV: body << reg(dst, w) << " = " << a << ";"; break;
                    case I_LDIN: body << reg(dst, w) << " = f4ld_in(p.inputs[" << wx << "], " << t << ");"; break;
                    case I_LDBUF: body << reg(dst, w) << " = f4ld_buf(p.buffers[" << wx << "], " << t << ");"; break;
                    case I_STBUF: body << guard << "f4st_buf(p.buffers[" << wx << "], " << t << ", " << a << ");"; break;
                    case I_STOUT: body << guard << "f4st_out(p, " << wx << ", " << t << ", " << a << ");"; break;
                    case I_TAP_IN: body << reg(dst, w) << " = f4tap_in(p.inputs[" << wx << "], " << t << ", " << sh << ");"; break;
                    case I_TAP_BUF: body << reg(dst, w) << " = f4tap_buf(p.buffers[" << wx << "], " << t << ", " << sh << ");"; break;
                    case I_GATE: body << reg(dst, w) << " = f4gate(" << a << ", " << t << ", (((unsigned long long)" << wx << " << 32) | " << wb << "));"; break;
                    case I_DLY_IN: body << reg(dst, w) << " = f4delay<0>(p.inputs[" << wx << "], no_buf, " << a << ", z4, " << t << ", p.sparkle_delay);"; break;
                    case I_DLY_BUF: body << reg(dst, w) << " = f4delay<1>(no_in, p.buffers[" << wx << "], " << a << ", z4, " << t << ", p.sparkle_delay);"; break;
                    case I_DLY_TI: body << reg(dst, w) << " = f4delay<2>(no_in, no_buf, " << a << ", " << b << ", " << t << ", p.sparkle_delay);"; break;
                    default: body << "/* unknown op " << op << " */"; break;
                }
                body << "\n";
            }
        }
        body << "  }\n  } break;\n";
        tables << "__device__ const unsigned frb_tab" << k << "[" << rows.size() << "][" << std::max<uint32_t>(n_words, 1) << "] = {";
        for (size_t m = 0; m < rows.size(); m++) {
            tables << (m ? ",{" : "{");
            if (rows[m].empty()) tables << "0u";
            for (size_t j = 0; j < rows[m].size(); j++) tables << (j ? "," : "") << rows[m][j] << "u";
            tables << "}";
        }
        tables << "};\n";
    }
    body << "  default: break;\n  }\n}\n";
    tables << "__device__ const unsigned frb_shape_of[" << std::max<size_t>(n_strands, 1) << "] = {";
    for (size_t sd = 0; sd < n_strands; sd++) tables << (sd ? "," : "") << shape_of[sd] << "u";
    if (!n_strands) tables << "0u";
    tables << "};\n__device__ const unsigned frb_idx_of[" << std::max<size_t>(n_strands, 1) << "] = {";
    for (size_t sd = 0; sd < n_strands; sd++) tables << (sd ? "," : "") << idx_of[sd] << "u";
    if (!n_strands) tables << "0u";
    tables << "};\n";
    o << tables.str() << body.str();
    return o.str();
}

bool jit_compile_to_cubin(const std::string& source, std::string* cubin, std::string* log) {
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
    nvrtcResult rc = a.CompileProgram(prog, (int)(sizeof(opts) / sizeof(opts[0])), opts);
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
    return ok;
}

struct JitKernel {
    CUmodule mod = nullptr;
    CUfunction fn = nullptr;
};

JitKernel* jit_build(const Stage& st, std::string* err) {
    std::string cubin, log;
    if (!jit_compile_to_cubin(jit_generate_source(st), &cubin, &log)) { if (err) *err = "NVRTC: " + log; return nullptr; }
    return jit_load(cubin, err);
}

JitKernel* jit_load(const std::string& cubin, std::string* err) {
    Api& a = api(true);
    if (!a.ok) { if (err) *err = a.why; return nullptr; }
    auto* k = new JitKernel();
    if (a.ModuleLoadData(&k->mod, cubin.data()) != CUDA_SUCCESS || a.ModuleGetFunction(&k->fn, k->mod, "frb_stage") != CUDA_SUCCESS) {
        if (k->mod) a.ModuleUnload(k->mod);
        delete k;
        if (err) *err = "cuModuleLoadData failed";
        return nullptr;
    }
    return k;
}

void jit_free(JitKernel* k) {
    if (!k) return;
    Api& a = api(true);
    if (a.ok && k->mod) a.ModuleUnload(k->mod);
    delete k;
}

bool jit_launch(JitKernel* k, const InterpParams& p, int sm_count, cudaStream_t stream) {
    Api& a = api(true);
    if (!a.ok || !k) return false;
    if (p.n_groups == 0) return true;
    unsigned long long blocks = (p.n_groups + 127) / 128;
    unsigned long long cap = (unsigned long long)sm_count * 16 / p.n_strands;   // rounded down: never a second wave of a few CTAs
    if (cap < 1) cap = 1;
    if (blocks > cap) blocks = cap;
    InterpParams pp = p;
    void* args[] = {&pp};
    return a.LaunchKernel(k->fn, (unsigned)blocks, p.n_strands, 1, 128, 1, 1, 0, (CUstream)stream, args, nullptr) == CUDA_SUCCESS;
}

}  // namespace frb
