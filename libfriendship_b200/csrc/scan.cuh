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
struct ChainStateDev;

std::shared_ptr<DirectFormDev> directform_create(const frb_directform_desc* d, cudaStream_t stream, std::string* err);
std::shared_ptr<FbDelayDev> fbdelay_create(const frb_fbdelay_desc* d, cudaStream_t stream, std::string* err);
uint32_t directform_lanes(const DirectFormDev& f);
uint32_t fbdelay_lanes(const FbDelayDev& f);
uint64_t fbdelay_max_delay(const FbDelayDev& f);
uint64_t fbdelay_min_delay(const FbDelayDev& f);

// y over [lo, hi) per lane; x read from ring in_bufs[lane], y written to ring out (first_out_buf + lane).
// State (x[n-1], x[n-2], y[n-1], y[n-2]) is read back from the rings, so consecutive calls continue exactly.
cudaError_t launch_directform(const DirectFormDev& f, const BufferDesc* d_bufdesc, const uint32_t* d_in_bufs,
                              uint32_t first_out_buf, uint64_t lo, uint64_t hi, int sm_count, cudaStream_t stream,
                              uint64_t* n_launches);
cudaError_t launch_fbdelay(const FbDelayDev& f, const BufferDesc* d_bufdesc, const uint32_t* d_in_bufs,
                           uint32_t first_out_buf, uint64_t lo, uint64_t hi, int sm_count, cudaStream_t stream,
                           uint64_t* n_launches);

// Fused chain: DirectForm lane l -> FbDelay lane l in one kernel (8 B instead of 16 B per lane-sample); the biquad's
// output is never stored, its carry lives in `st`.  x from ring in_bufs[lane] (the biquad's inputs), z to ring
// first_out_buf + lane (the comb's outputs).  lo must be 0 or the previous call's hi.
bool chain_fusable(const DirectFormDev& df, const FbDelayDev& fb);
std::shared_ptr<ChainStateDev> chain_state_create(uint32_t n_lanes);
// `exciter` != nullptr: the biquad's input of lane l is voice d_exc_voice[l] of that one-partial oscillator bank, evaluated
// inside the kernel (osc_one.cuh; same bits as the bank's own kernel): d_in_bufs is not read and the chain's only HBM
// traffic is its output.
struct OscOneSrc;
cudaError_t launch_dfcomb(const DirectFormDev& df, const FbDelayDev& fb, ChainStateDev& st, const BufferDesc* d_bufdesc,
                          const uint32_t* d_in_bufs, uint32_t first_out_buf, uint64_t lo, uint64_t hi, cudaStream_t stream,
                          uint64_t* n_launches, const OscOneSrc* exciter = nullptr, const uint32_t* d_exc_voice = nullptr);

}  // namespace frb
