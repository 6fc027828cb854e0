// K1 (cdf_diff likelihood mode) — placeholder until the kernel lands; reports SIC_E_UNSUPPORTED loudly.
#include "common.cuh"
namespace sic {
int bottleneck_cdfdiff_fwd(const float *, const float *, uint64_t *, const float *, const float *, const float *, int, int, int,
                           int, int, float *, float *, float *, void *, size_t, cudaStream_t) {
    set_error("SIC_LIK_STUDENTT_CDFDIFF forward is not implemented yet");
    return SIC_E_UNSUPPORTED;
}
int bottleneck_cdfdiff_bwd(const float *, const float *, const float *, const float *, const float *, const float *, const float *,
                           int, int, int, int, int, float *, float *, float *, float *, void *, size_t, cudaStream_t) {
    set_error("SIC_LIK_STUDENTT_CDFDIFF backward is not implemented yet");
    return SIC_E_UNSUPPORTED;
}
}  // namespace sic
