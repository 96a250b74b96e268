// osc.cuh — K1 oscillator bank (extension node FRB_KIND_OSCBANK); see osc.cu
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "../../include/friendship_b200.h"
#include "interp.cuh"

namespace frb {

struct OscBankDev;   // device-resident, preprocessed bank (defined in osc.cu)

struct OscBankInfo {
    uint32_t n_voices;
    uint64_t n_partials;
};

// Uploads and preprocesses a bank (class/rank, fill and fp64 setup kernels; host work is O(n_voices)).  The arrays of `d`
// may be pageable or pinned host memory.  Returns nullptr and sets *err on failure.
// `recycle`: the bank this one replaces under the same key, if any.  The new bank takes over all its device allocations
// (records, scratch, partial-range planes), so that re-defining a bank before every render allocates nothing.  On success
// `recycle` is left hollow; on failure it gets its allocations back (check osc_usable: a failed reallocation loses them).
std::shared_ptr<OscBankDev> osc_create(const frb_oscbank_desc* d, cudaStream_t stream, std::string* err,
                                       const std::shared_ptr<OscBankDev>& recycle = nullptr,
                                       uint32_t shard_rank = 0, uint32_t shard_world = 1,
                                       bool allow_tensor = true);   // false: FRB_FLAG_NO_TENSOR_OSC
OscBankInfo osc_info(const OscBankDev& b);
bool osc_usable(const OscBankDev& b);   // false for a bank whose allocations were lost to a failed re-definition

cudaError_t osc_init_device();

// Banks whose voices have at most one partial each (and no partial-range split) can also be evaluated inside another
// kernel (osc_one.cuh): fills *out with the bank's device arrays and returns true for such a bank.
struct OscOneSrc;
bool osc_one_source(const OscBankDev& b, OscOneSrc* out);

// Renders every voice of the bank over absolute times [lo, hi) into the voices' ring buffers
// bufdesc[first_buf + v].  `anchor` = samples between exact re-anchors of each partial.
cudaError_t launch_osc(const OscBankDev& b, const BufferDesc* d_bufdesc, uint32_t first_buf, uint64_t lo, uint64_t hi,
                       uint32_t anchor, int sm_count, cudaStream_t stream, uint64_t* n_launches,
                       uint64_t* n_tensor_launches = nullptr);   // of *n_launches: the matrix-product kernels (osc_gemm.cuh, osc_tc.cuh)

}  // namespace frb
