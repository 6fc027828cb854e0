// K1: fused quantise + likelihood + rate (forward) and its analytic backward.
//
// Replaces, per call, the eager chains of  model.py:27-35 (quantize),  distributions.py:20-31 (Student-t density),
// distributions.py:39-46 (Gaussian z prior) and the .sum() reductions of model.py:77  (paths relative to
// /root/reference/code/modelv2).
//
// Work decomposition: a ROW is one (b,c) plane of HW elements; rows are cut into SEGMENTS of `seg` elements and one
// warp owns one (row, segment) unit.  In the broadcast/channel layouts sigma/nu are row constants, so the lgamma /
// log prefactor is evaluated once per unit (redundantly across the 32 lanes: one issue slot either way) and the
// per-element work is  mul, mul, mul, log1p, fma.  Loads/stores are 128-bit, coalesced, L1-bypassing; four vectors
// are in flight per lane.  Each unit leaves one partial sum; the last CTA to retire (ticket counter) folds the partials
// per patch in a fixed order with a float64 accumulator => bit-reproducible rate, single launch.
//
// HBM-bound by design: 12 B/element (broadcast) against ~25 issue slots per element-lane.
#include "bottleneck_kernels.cuh"

namespace sic {
namespace {

inline bool aligned16(const void *p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int resolve_mode(int lik_mode, int param_layout, int *mode) {
    if (lik_mode == SIC_LIK_STUDENTT_DENSITY && param_layout == SIC_PARAM_BROADCAST) { *mode = MODE_T_BCAST; return 0; }
    if (lik_mode == SIC_LIK_STUDENTT_DENSITY && param_layout == SIC_PARAM_SPATIAL) { *mode = MODE_T_SPATIAL; return 0; }
    if (lik_mode == SIC_LIK_GAUSSIAN && param_layout == SIC_PARAM_CHANNEL) { *mode = MODE_GAUSS; return 0; }
    return SIC_E_UNSUPPORTED;
}

}  // namespace

// cdf_diff mode lives in bottleneck_cdf.cu
int bottleneck_cdfdiff_fwd(const float *y, const float *noise, uint64_t *philox, const float *mu, const float *sigma,
                           const float *nu, int B, int C, int HW, int quant_mode, int param_layout, float *y_tilde, float *nll,
                           float *bits, void *workspace, size_t workspace_bytes, cudaStream_t st);
int bottleneck_cdfdiff_bwd(const float *y_tilde, const float *mu, const float *sigma, const float *nu, const float *g_nll,
                           const float *g_bits, const float *g_ytilde, int B, int C, int HW, int quant_mode, int param_layout,
                           float *dy, float *dmu, float *dsigma, float *dnu, void *workspace, size_t workspace_bytes,
                           cudaStream_t st);

SIC_REGISTER_KERNEL("bottleneck_fwd_kernel<0,1,3,0>", bottleneck_fwd_kernel<MODE_T_BCAST, true, SIC_QUANT_NOISE_PHILOX, false>);
SIC_REGISTER_KERNEL("bottleneck_fwd_kernel<0,1,1,0>", bottleneck_fwd_kernel<MODE_T_BCAST, true, SIC_QUANT_ROUND, false>);
SIC_REGISTER_KERNEL("bottleneck_fwd_kernel<1,1,3,0>", bottleneck_fwd_kernel<MODE_T_SPATIAL, true, SIC_QUANT_NOISE_PHILOX, false>);
SIC_REGISTER_KERNEL("bottleneck_fwd_kernel<2,1,3,0>", bottleneck_fwd_kernel<MODE_GAUSS, true, SIC_QUANT_NOISE_PHILOX, false>);
SIC_REGISTER_KERNEL("bottleneck_bwd_kernel<0,1>", bottleneck_bwd_kernel<MODE_T_BCAST, true>);
SIC_REGISTER_KERNEL("bottleneck_bwd_kernel<1,1>", bottleneck_bwd_kernel<MODE_T_SPATIAL, true>);
}  // namespace sic

using namespace sic;

extern "C" size_t sic_bottleneck_workspace_bytes(int B, int C, int HW) {
    if (B <= 0 || C <= 0 || HW <= 0) return kWsHeader;
    long units = (long)B * C * ((HW + kMinSeg - 1) / kMinSeg);
    return kWsHeader + (size_t)units * 3 * sizeof(float);
}

extern "C" int sic_bottleneck_fwd(const float *y, const float *noise, uint64_t *philox, const float *mu,
                                  const float *sigma, const float *nu, int B, int C, int HW, int quant_mode, int lik_mode,
                                  int param_layout, float *y_tilde, float *nll, float *bits, void *workspace,
                                  size_t workspace_bytes, void *stream) {
    SIC_CHECK_ARG(B > 0 && C > 0 && HW > 0, "sic_bottleneck_fwd: empty shape B=%d C=%d HW=%d", B, C, HW);
    SIC_CHECK_ARG(y && bits && sigma && workspace, "sic_bottleneck_fwd: null required pointer");
    SIC_CHECK_ARG(y_tilde || quant_mode == SIC_QUANT_NONE, "sic_bottleneck_fwd: y_tilde may be NULL only with SIC_QUANT_NONE");
    SIC_CHECK_ARG(quant_mode >= SIC_QUANT_NONE && quant_mode <= SIC_QUANT_NOISE_PHILOX, "sic_bottleneck_fwd: bad quant_mode %d", quant_mode);
    SIC_CHECK_ARG(quant_mode != SIC_QUANT_NOISE_TENSOR || noise, "sic_bottleneck_fwd: noise tensor required");
    SIC_CHECK_ARG(quant_mode != SIC_QUANT_NOISE_PHILOX || philox, "sic_bottleneck_fwd: philox state required");
    SIC_CHECK_ARG((long)B * C * HW < (1L << 40), "sic_bottleneck_fwd: tensor too large");
    if (workspace_bytes < sic_bottleneck_workspace_bytes(B, C, HW)) {
        set_error("sic_bottleneck_fwd: workspace %zu < %zu bytes", workspace_bytes, sic_bottleneck_workspace_bytes(B, C, HW));
        return SIC_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (lik_mode == SIC_LIK_STUDENTT_CDFDIFF)
        return bottleneck_cdfdiff_fwd(y, noise, philox, mu, sigma, nu, B, C, HW, quant_mode, param_layout, y_tilde, nll, bits,
                                      workspace, workspace_bytes, st);
    int mode;
    if (resolve_mode(lik_mode, param_layout, &mode) != 0) {
        set_error("sic_bottleneck_fwd: unsupported lik_mode %d with param_layout %d", lik_mode, param_layout);
        return SIC_E_UNSUPPORTED;
    }
    SIC_CHECK_ARG(mode == MODE_GAUSS || nu, "sic_bottleneck_fwd: nu required for Student-t");
    int mu_layout = mode == MODE_GAUSS ? SIC_PARAM_CHANNEL : param_layout;
    Shape sh = make_shape(B, C, HW);
    bool vec = (HW % 4 == 0) && aligned16(y) && aligned16(noise) && aligned16(y_tilde) && aligned16(nll) &&
               (mode != MODE_T_SPATIAL || (aligned16(sigma) && aligned16(nu))) && (param_layout != SIC_PARAM_SPATIAL || aligned16(mu));
    unsigned int *ticket = reinterpret_cast<unsigned int *>(workspace);
    float *psum = reinterpret_cast<float *>(static_cast<char *>(workspace) + kWsHeader);
    dim3 grid((unsigned)((sh.units + kWarpsPerCta - 1) / kWarpsPerCta)), block(kThreads);
    const bool has_mu = mu != nullptr;
#define LAUNCH(M, V, Q, U)                                                                                              \
    bottleneck_fwd_kernel<M, V, Q, U><<<grid, block, 0, st>>>(y, noise, philox, mu, sigma, nu, sh, mu_layout, y_tilde, nll, \
                                                              bits, psum, ticket)
#define LAUNCH_Q(M, V, U)                                                        \
    switch (quant_mode) {                                                        \
        case SIC_QUANT_NONE: LAUNCH(M, V, SIC_QUANT_NONE, U); break;             \
        case SIC_QUANT_ROUND: LAUNCH(M, V, SIC_QUANT_ROUND, U); break;           \
        case SIC_QUANT_NOISE_TENSOR: LAUNCH(M, V, SIC_QUANT_NOISE_TENSOR, U); break; \
        default: LAUNCH(M, V, SIC_QUANT_NOISE_PHILOX, U); break;                 \
    }
#define LAUNCH_V(M, U)              \
    if (vec) { LAUNCH_Q(M, true, U) } \
    else { LAUNCH_Q(M, false, U) }
    if (mode == MODE_T_BCAST) { if (has_mu) { LAUNCH_V(MODE_T_BCAST, true) } else { LAUNCH_V(MODE_T_BCAST, false) } }
    else if (mode == MODE_T_SPATIAL) { if (has_mu) { LAUNCH_V(MODE_T_SPATIAL, true) } else { LAUNCH_V(MODE_T_SPATIAL, false) } }
    else { if (has_mu) { LAUNCH_V(MODE_GAUSS, true) } else { LAUNCH_V(MODE_GAUSS, false) } }
#undef LAUNCH_V
#undef LAUNCH_Q
#undef LAUNCH
    SIC_CHECK_LAUNCH("sic_bottleneck_fwd");
    return 0;
}

extern "C" int sic_bottleneck_bwd(const float *y_tilde, const float *mu, const float *sigma, const float *nu,
                                  const float *g_nll, const float *g_bits, const float *g_ytilde, int B, int C, int HW,
                                  int quant_mode, int lik_mode, int param_layout, float *dy, float *dmu, float *dsigma,
                                  float *dnu, void *workspace, size_t workspace_bytes, void *stream) {
    SIC_CHECK_ARG(B > 0 && C > 0 && HW > 0, "sic_bottleneck_bwd: empty shape B=%d C=%d HW=%d", B, C, HW);
    SIC_CHECK_ARG(y_tilde && sigma && workspace, "sic_bottleneck_bwd: null required pointer");
    SIC_CHECK_ARG(dmu == nullptr || mu != nullptr, "sic_bottleneck_bwd: dmu requested without mu");
    if (workspace_bytes < sic_bottleneck_workspace_bytes(B, C, HW)) {
        set_error("sic_bottleneck_bwd: workspace %zu < %zu bytes", workspace_bytes, sic_bottleneck_workspace_bytes(B, C, HW));
        return SIC_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (lik_mode == SIC_LIK_STUDENTT_CDFDIFF)
        return bottleneck_cdfdiff_bwd(y_tilde, mu, sigma, nu, g_nll, g_bits, g_ytilde, B, C, HW, quant_mode, param_layout, dy,
                                      dmu, dsigma, dnu, workspace, workspace_bytes, st);
    int mode;
    if (resolve_mode(lik_mode, param_layout, &mode) != 0) {
        set_error("sic_bottleneck_bwd: unsupported lik_mode %d with param_layout %d", lik_mode, param_layout);
        return SIC_E_UNSUPPORTED;
    }
    SIC_CHECK_ARG(mode == MODE_GAUSS || nu, "sic_bottleneck_bwd: nu required for Student-t");
    int mu_layout = mode == MODE_GAUSS ? SIC_PARAM_CHANNEL : param_layout;
    Shape sh = make_shape(B, C, HW);
    bool vec = (HW % 4 == 0) && aligned16(y_tilde) && aligned16(g_nll) && aligned16(g_ytilde) && aligned16(dy) &&
               (mode != MODE_T_SPATIAL || (aligned16(sigma) && aligned16(nu) && aligned16(dsigma) && aligned16(dnu))) &&
               (param_layout != SIC_PARAM_SPATIAL || (aligned16(mu) && aligned16(dmu)));
    unsigned int *ticket = reinterpret_cast<unsigned int *>(workspace);
    float *pa = reinterpret_cast<float *>(static_cast<char *>(workspace) + kWsHeader);
    float *pb = pa + sh.units, *pc = pb + sh.units;
    dim3 grid((unsigned)((sh.units + kWarpsPerCta - 1) / kWarpsPerCta)), block(kThreads);
#define LAUNCH(M, V)                                                                                                          \
    bottleneck_bwd_kernel<M, V><<<grid, block, 0, st>>>(y_tilde, mu, sigma, nu, g_nll, g_bits, g_ytilde, sh, quant_mode, mu_layout, \
                                                        dy, dmu, dsigma, dnu, pa, pb, pc, ticket)
    if (mode == MODE_T_BCAST) { if (vec) LAUNCH(MODE_T_BCAST, true); else LAUNCH(MODE_T_BCAST, false); }
    else if (mode == MODE_T_SPATIAL) { if (vec) LAUNCH(MODE_T_SPATIAL, true); else LAUNCH(MODE_T_SPATIAL, false); }
    else { if (vec) LAUNCH(MODE_GAUSS, true); else LAUNCH(MODE_GAUSS, false); }
#undef LAUNCH
    SIC_CHECK_LAUNCH("sic_bottleneck_bwd");
    return 0;
}
