// jit.hpp — see jit.cc
#pragma once
#include <cuda_runtime.h>

#include <string>

#include "interp.cuh"
#include "schedule.hpp"

namespace frb {

struct JitKernel;

// CUDA source of the fused kernel for one stage program (thread = 8 consecutive samples, like the interpreter).
std::string jit_generate_source(const Stage& st);
// instructions held as straight-line code in that source (one body per distinct strand shape): what NVRTC's time depends on
size_t jit_code_instructions(const Stage& st);
// NVRTC: source -> sm_100a cubin.  Needs no GPU.  Returns false (and the compiler log) on failure.
bool jit_compile_to_cubin(const std::string& source, std::string* cubin, std::string* log);
// compile + load into the current context; nullptr on failure
JitKernel* jit_build(const Stage& st, std::string* err);
// load an already compiled cubin (jit_compile_to_cubin) into the current context
JitKernel* jit_load(const std::string& cubin, std::string* err);
void jit_free(JitKernel* k);
bool jit_launch(JitKernel* k, const InterpParams& p, int sm_count, cudaStream_t stream);

}  // namespace frb
