// osc.cu — placeholder until K1 lands (next commit): definitions are recorded, launches are refused.
#include "osc.cuh"

namespace frb {
struct OscBankDev { uint32_t n_voices; uint64_t n_partials; };
std::shared_ptr<OscBankDev> osc_create(const frb_oscbank_desc* d, cudaStream_t, std::string*) {
    auto b = std::make_shared<OscBankDev>();
    b->n_voices = d->n_voices; b->n_partials = d->n_partials;
    return b;
}
OscBankInfo osc_info(const OscBankDev& b) { return OscBankInfo{b.n_voices, b.n_partials}; }
cudaError_t launch_osc(const OscBankDev&, const BufferDesc*, uint32_t, uint64_t, uint64_t, uint32_t, int, cudaStream_t, uint64_t*) {
    return cudaErrorNotSupported;
}
}  // namespace frb
