// scan.cu — placeholder until K4 lands: definitions are recorded, launches are refused.
#include "scan.cuh"
#include <algorithm>
#include <vector>

namespace frb {
struct DirectFormDev { uint32_t n_lanes; };
struct FbDelayDev { uint32_t n_lanes; uint64_t max_delay; };
std::shared_ptr<DirectFormDev> directform_create(const frb_directform_desc* d, cudaStream_t, std::string*) {
    auto f = std::make_shared<DirectFormDev>(); f->n_lanes = d->n_lanes; return f;
}
std::shared_ptr<FbDelayDev> fbdelay_create(const frb_fbdelay_desc* d, cudaStream_t, std::string* err) {
    auto f = std::make_shared<FbDelayDev>(); f->n_lanes = d->n_lanes; f->max_delay = 0;
    for (uint32_t i = 0; i < d->n_lanes; i++) {
        if (d->delay[i] < 1) { if (err) *err = "fbdelay: delay must be >= 1"; return nullptr; }
        f->max_delay = std::max<uint64_t>(f->max_delay, d->delay[i]);
    }
    return f;
}
uint32_t directform_lanes(const DirectFormDev& f) { return f.n_lanes; }
uint32_t fbdelay_lanes(const FbDelayDev& f) { return f.n_lanes; }
uint64_t fbdelay_max_delay(const FbDelayDev& f) { return f.max_delay; }
cudaError_t launch_directform(const DirectFormDev&, const BufferDesc*, const uint32_t*, uint32_t, uint64_t, uint64_t, int, cudaStream_t, uint64_t*) { return cudaErrorNotSupported; }
cudaError_t launch_fbdelay(const FbDelayDev&, const BufferDesc*, const uint32_t*, uint32_t, uint64_t, uint64_t, int, cudaStream_t, uint64_t*) { return cudaErrorNotSupported; }
}  // namespace frb
