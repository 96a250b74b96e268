// interp.cuh — launch interface of the fused elementwise/Delay stage kernel (interp.cu)
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "schedule.hpp"

namespace frb {

constexpr unsigned INTERP_THREADS = 128;
constexpr int INTERP_VW = 2;           // float4 per thread per register: 8 consecutive samples

#include "interp_device.inc"

cudaError_t launch_fold(const BufferDesc* d_bufdesc, unsigned first, unsigned count, unsigned out_buf,
                        unsigned long long lo, unsigned long long hi, int sm_count, cudaStream_t stream);
cudaError_t interp_init_device();
cudaError_t launch_interp(const InterpParams& p, unsigned n_regs, int sm_count, cudaStream_t stream);

}  // namespace frb
