// K1, cdf_diff likelihood mode: host launchers for the MODE_CDF_* instances of the shared kernel templates
// (bottleneck_kernels.cuh holds the quadrature and its derivation).  north_star's named kernel; the reference only
// *intends* this likelihood (eval_selfcontained_entropy.py:56-59 calls a CDF torch does not have).
#include "bottleneck_kernels.cuh"

namespace sic {
namespace {
inline bool al16(const void *p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
}  // namespace

SIC_REGISTER_KERNEL("bottleneck_fwd_kernel<3,1,3,0>", bottleneck_fwd_kernel<MODE_CDF_BCAST, true, SIC_QUANT_NOISE_PHILOX, false>);
SIC_REGISTER_KERNEL("bottleneck_bwd_kernel<3,1>", bottleneck_bwd_kernel<MODE_CDF_BCAST, true>);

int bottleneck_cdfdiff_fwd(const float *y, const float *noise, uint64_t *philox, const float *mu, const float *sigma,
                           const float *nu, int B, int C, int HW, int quant_mode, int param_layout, float *y_tilde, float *nll,
                           float *bits, void *workspace, size_t workspace_bytes, cudaStream_t st) {
    if (param_layout != SIC_PARAM_BROADCAST && param_layout != SIC_PARAM_SPATIAL) {
        set_error("sic_bottleneck_fwd: cdf_diff needs broadcast or spatial sigma/nu");
        return SIC_E_UNSUPPORTED;
    }
    SIC_CHECK_ARG(nu != nullptr, "sic_bottleneck_fwd: nu required for Student-t");
    const bool spatial = param_layout == SIC_PARAM_SPATIAL;
    Shape sh = make_shape(B, C, HW);
    bool vec = (HW % 4 == 0) && al16(y) && al16(noise) && al16(y_tilde) && al16(nll) && (!spatial || (al16(sigma) && al16(nu) && al16(mu)));
    unsigned int *ticket = reinterpret_cast<unsigned int *>(workspace);
    float *psum = reinterpret_cast<float *>(static_cast<char *>(workspace) + kWsHeader);
    dim3 grid((unsigned)((sh.units + kWarpsPerCta - 1) / kWarpsPerCta)), block(kThreads);
    const bool has_mu = mu != nullptr;
#define LAUNCH(M, V, Q, U) \
    bottleneck_fwd_kernel<M, V, Q, U><<<grid, block, 0, st>>>(y, noise, philox, mu, sigma, nu, sh, param_layout, y_tilde, nll, bits, psum, ticket)
#define LAUNCH_Q(M, V, U)                                                            \
    switch (quant_mode) {                                                            \
        case SIC_QUANT_NONE: LAUNCH(M, V, SIC_QUANT_NONE, U); break;                 \
        case SIC_QUANT_ROUND: LAUNCH(M, V, SIC_QUANT_ROUND, U); break;               \
        case SIC_QUANT_NOISE_TENSOR: LAUNCH(M, V, SIC_QUANT_NOISE_TENSOR, U); break; \
        default: LAUNCH(M, V, SIC_QUANT_NOISE_PHILOX, U); break;                     \
    }
#define LAUNCH_V(M, U)                \
    if (vec) { LAUNCH_Q(M, true, U) } \
    else { LAUNCH_Q(M, false, U) }
    if (spatial) { if (has_mu) { LAUNCH_V(MODE_CDF_SPATIAL, true) } else { LAUNCH_V(MODE_CDF_SPATIAL, false) } }
    else { if (has_mu) { LAUNCH_V(MODE_CDF_BCAST, true) } else { LAUNCH_V(MODE_CDF_BCAST, false) } }
#undef LAUNCH_V
#undef LAUNCH_Q
#undef LAUNCH
    SIC_CHECK_LAUNCH("sic_bottleneck_fwd (cdf_diff)");
    return 0;
}

int bottleneck_cdfdiff_bwd(const float *y_tilde, const float *mu, const float *sigma, const float *nu, const float *g_nll,
                           const float *g_bits, const float *g_ytilde, int B, int C, int HW, int quant_mode, int param_layout,
                           float *dy, float *dmu, float *dsigma, float *dnu, void *workspace, size_t workspace_bytes,
                           cudaStream_t st) {
    if (param_layout != SIC_PARAM_BROADCAST && param_layout != SIC_PARAM_SPATIAL) {
        set_error("sic_bottleneck_bwd: cdf_diff needs broadcast or spatial sigma/nu");
        return SIC_E_UNSUPPORTED;
    }
    SIC_CHECK_ARG(nu != nullptr, "sic_bottleneck_bwd: nu required for Student-t");
    const bool spatial = param_layout == SIC_PARAM_SPATIAL;
    Shape sh = make_shape(B, C, HW);
    bool vec = (HW % 4 == 0) && al16(y_tilde) && al16(g_nll) && al16(g_ytilde) && al16(dy) &&
               (!spatial || (al16(sigma) && al16(nu) && al16(dsigma) && al16(dnu) && al16(mu) && al16(dmu)));
    unsigned int *ticket = reinterpret_cast<unsigned int *>(workspace);
    float *pa = reinterpret_cast<float *>(static_cast<char *>(workspace) + kWsHeader);
    float *pb = pa + sh.units, *pc = pb + sh.units;
    dim3 grid((unsigned)((sh.units + kWarpsPerCta - 1) / kWarpsPerCta)), block(kThreads);
#define LAUNCH(M, V)                                                                                                              \
    bottleneck_bwd_kernel<M, V><<<grid, block, 0, st>>>(y_tilde, mu, sigma, nu, g_nll, g_bits, g_ytilde, sh, quant_mode, param_layout, \
                                                        dy, dmu, dsigma, dnu, pa, pb, pc, ticket)
    if (spatial) { if (vec) LAUNCH(MODE_CDF_SPATIAL, true); else LAUNCH(MODE_CDF_SPATIAL, false); }
    else { if (vec) LAUNCH(MODE_CDF_BCAST, true); else LAUNCH(MODE_CDF_BCAST, false); }
#undef LAUNCH
    SIC_CHECK_LAUNCH("sic_bottleneck_bwd (cdf_diff)");
    return 0;
}

}  // namespace sic
