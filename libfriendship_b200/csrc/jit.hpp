// jit.hpp — see jit.cc
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <memory>
#include <string>
#include <vector>

#include "interp.cuh"
#include "schedule.hpp"

namespace frb {

struct JitKernel;

// What the stage JIT makes of one stage program (thread = 8 consecutive samples, like the interpreter):
//   source      CUDA source — a function of the program's STRUCTURE only (runs of like instruction groups are loops)
//   table       the operand table the kernel reads: loop trip counts, slot / buffer indices, immediates, shifts
//   code_instrs statements the compiler sees (unrolled loop bodies counted as often as unrolled): what NVRTC's time depends on
struct JitProgram {
    std::string source;
    std::vector<uint32_t> table;
    size_t code_instrs = 0;
    unsigned groups_per_thread = 1;   // 8-sample groups a thread walks per iteration (2 for small programs): sizes the grid
};
JitProgram jit_generate(const Stage& st);
std::string jit_generate_source(const Stage& st);
size_t jit_code_instructions(const Stage& st);
// NVRTC: source -> sm_100a cubin, through a process-wide cache keyed by the source text.  Needs no GPU.
// Returns false (and the compiler log) on failure.
bool jit_compile_to_cubin(const std::string& source, std::string* cubin, std::string* log);
std::shared_ptr<const std::string> jit_cache_lookup(const std::string& source);
// the same on a thread of its own; done: 0 running, 1 ok (cubin), -1 failed (log)
struct JitJob {
    std::atomic<int> done{0};
    std::string cubin, log;
};
std::shared_ptr<JitJob> jit_compile_async(std::string source);
void jit_wait_idle();
// compile + load into the current context; nullptr on failure
JitKernel* jit_build(const Stage& st, std::string* err);
// load an already compiled cubin (jit_compile_to_cubin) and its operand table into the current context
JitKernel* jit_load(const std::string& cubin, const std::vector<uint32_t>& table, unsigned groups_per_thread, std::string* err);
void jit_free(JitKernel* k);
bool jit_launch(JitKernel* k, const InterpParams& p, int sm_count, cudaStream_t stream);

}  // namespace frb
