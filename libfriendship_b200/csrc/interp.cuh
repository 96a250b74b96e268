// interp.cuh — launch interface of the fused elementwise/Delay stage kernel (interp.cu)
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "schedule.hpp"

namespace frb {

constexpr unsigned INTERP_THREADS = 128;
constexpr int INTERP_VW = 2;           // float4 per thread per register: 8 consecutive samples

// External-input history of one slot, device resident: values for absolute times [base_time, end_time),
// zeros elsewhere (reference src/render/reference.rs:22-25, :90-96).  base_time is a multiple of 4.
struct InputDesc {
    const float* data;
    unsigned long long base_time;
    unsigned long long end_time;
};

// A materialised signal: ring buffer in HBM addressed by absolute time, data[t & mask]; capacity is a power of two.
struct BufferDesc {
    float* data;
    unsigned long long mask;
};

struct InterpParams {
    const uint32_t* program;        // device: Instr words
    unsigned n_instr;               // including the trailing I_END pair
    unsigned n_regs;                // set by launch_interp: float4 register columns per thread
    unsigned prog_in_smem;          // set by launch_interp: the program is staged in shared memory
    const InputDesc* inputs;        // device table, indexed by external input slot
    const BufferDesc* buffers;      // device table, indexed by buffer id
    float* out;                     // device: [n_slots x out_stride], column 0 == time t0
    unsigned long long out_stride;
    unsigned long long t_begin;     // first absolute time evaluated (multiple of 8)
    unsigned long long n_groups;    // number of 8-sample groups evaluated
    unsigned long long t0, t1;      // output window [t0, t1)
    int out_vec_ok;                 // rows and t0 are 16-byte aligned: 128-bit output stores allowed
    int sparkle_delay;              // FRB_FLAG_SPARKLE_DELAY
};

cudaError_t launch_fold(const BufferDesc* d_bufdesc, unsigned first, unsigned count, unsigned out_buf,
                        unsigned long long lo, unsigned long long hi, int sm_count, cudaStream_t stream);
cudaError_t launch_interp(const InterpParams& p, unsigned n_regs, int sm_count, cudaStream_t stream);

}  // namespace frb
