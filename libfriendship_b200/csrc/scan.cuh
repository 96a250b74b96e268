// scan.cuh — K4 Direct-Form recurrences (extension nodes FRB_KIND_DIRECTFORM / FRB_KIND_FBDELAY); see scan.cu
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <memory>
#include <string>

#include "../../include/friendship_b200.h"
#include "interp.cuh"

namespace frb {

struct DirectFormDev;
struct FbDelayDev;

std::shared_ptr<DirectFormDev> directform_create(const frb_directform_desc* d, cudaStream_t stream, std::string* err);
std::shared_ptr<FbDelayDev> fbdelay_create(const frb_fbdelay_desc* d, cudaStream_t stream, std::string* err);
uint32_t directform_lanes(const DirectFormDev& f);
uint32_t fbdelay_lanes(const FbDelayDev& f);
uint64_t fbdelay_max_delay(const FbDelayDev& f);

// y over [lo, hi) per lane; x read from ring in_bufs[lane], y written to ring out (first_out_buf + lane).
// State (x[n-1], x[n-2], y[n-1], y[n-2]) is read back from the rings, so consecutive calls continue exactly.
cudaError_t launch_directform(const DirectFormDev& f, const BufferDesc* d_bufdesc, const uint32_t* d_in_bufs,
                              uint32_t first_out_buf, uint64_t lo, uint64_t hi, int sm_count, cudaStream_t stream,
                              uint64_t* n_launches);
cudaError_t launch_fbdelay(const FbDelayDev& f, const BufferDesc* d_bufdesc, const uint32_t* d_in_bufs,
                           uint32_t first_out_buf, uint64_t lo, uint64_t hi, int sm_count, cudaStream_t stream,
                           uint64_t* n_launches);

}  // namespace frb
